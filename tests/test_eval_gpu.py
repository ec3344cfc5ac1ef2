"""The fused N-sample evaluation tail (uz_eval_sample_stats / uz_ncc_dice_from_sums / uz_ged_pairwise behind
b200.train.EvalStep) against (a) the reference-shaped path through the module API + drop-in utils, (b) the numpy oracle
of the reference's utils.py, and -- with two GPUs -- the sample-sharded evaluation against the single-rank one.
Bar: GED bit exact (integer IoU counts, Python-order sums); NCC 1e-4 absolute (north star); Dice exact (integer counts)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as mo
from oracle import synth
from tests.keygrammar import dropin_phiseg

pytestmark = pytest.mark.gpu


def _fake_noise():
    """deterministic stand-in for torch.randn_like: a function of the flat index, so equal shapes see equal noise"""
    def fake(t, **kw):
        n = t.numel()
        return torch.sin(torch.arange(n, device=t.device, dtype=torch.float32) * 12.9898).mul(1.7).reshape(t.shape)
    return fake


def _net(filters=(16, 32, 32, 32, 32, 32, 32), seed=1):
    net = dropin_phiseg(list(filters))
    net.load_state_dict(synth.synth_state_dict(net.state_dict(), seed=seed))
    return net.cuda().eval()


def _reference_metrics(probs, pred, masks, n_classes, annot):
    """numpy restatement of utils.py on what the module API returned: GED, NCC, Dice per class (train_model.py:207-222)"""
    ged = mo.generalised_energy_distance(pred, masks, n_classes - 1, range(1, n_classes))
    ncc = float(mo.variance_ncc_dist(probs, mo.convert_batch_to_onehot(masks[:, None], n_classes))[0])
    mean_pred = probs.mean(0).argmax(0)
    dice = []
    for lbl in range(n_classes):
        bp, bg = mean_pred == lbl, masks[annot] == lbl
        if bp.sum() == 0 and bg.sum() == 0:
            dice.append(1.0)
        elif bp.sum() == 0 or bg.sum() == 0:
            dice.append(0.0)
        else:
            dice.append(2.0 * float((bp & bg).sum()) / float(bp.sum() + bg.sum()))
    return ged, ncc, dice


@pytest.mark.parametrize('images', [1, 3])
def test_fused_evaluation_matches_module_api_and_oracle(images):
    from b200 import train
    net = _net()
    n, C = 12, 2
    patch, labels, _ = synth.lidc_like_batch(images, seed=11)
    img = patch[:, 0].contiguous().pin_memory()                  # [I, H, W]
    lab = labels.contiguous().pin_memory()                       # [I, H, W, M]
    orig = torch.randn_like
    torch.randn_like = _fake_noise()
    try:
        ev = train.EvalStep(net, n, C, images_per_step=images, use_graph=False)
        got = ev.run_host(img, lab) if images > 1 else None
        if images == 1:
            ev.run_host(img[0], lab[0])
            got = ev.out_host.clone()
        # the same forward through the module API (full-resolution logits), same fake noise (same shapes)
        with torch.no_grad():
            masks = lab.cuda().permute(0, 3, 1, 2).contiguous()
            s_list = net.forward(img.cuda()[:, None], masks[:, 0:1].float(), training=False, replicate=n)
            probs = net.accumulate_output(s_list, use_softmax=True)       # [n*I, C, H, W], index = sample * I + image
    finally:
        torch.randn_like = orig
    probs = probs.view(n, images, C, *probs.shape[-2:]).cpu().numpy()
    for i in range(images):
        p = probs[:, i]
        pred = p.argmax(1)
        ged, ncc, dice = _reference_metrics(p, pred, lab[i].permute(2, 0, 1).numpy(), C, 0)
        assert float(got[i, 0]) == ged, (float(got[i, 0]), ged)                # bit exact
        assert abs(float(got[i, 1]) - ncc) < 1e-4, (float(got[i, 1]), ncc)     # north star: 1e-4 absolute
        assert [float(v) for v in got[i, 2:]] == dice
        assert 0.0 < (pred != 0).mean() < 1.0, 'degenerate masks would skip the popcount branch'


def test_fused_equals_unfused_evalstep():
    from b200 import train
    net = _net(seed=2)
    patch, labels, _ = synth.lidc_like_batch(1, seed=5)
    img, lab = patch[0, 0].contiguous().pin_memory(), labels[0].contiguous().pin_memory()
    orig = torch.randn_like
    torch.randn_like = _fake_noise()
    try:
        a = train.EvalStep(net, 16, 2, use_graph=False, fused=True).run_host(img, lab)
        b = train.EvalStep(net, 16, 2, use_graph=False, fused=False).run_host(img, lab)
        c = train.EvalStep(net, 16, 2, use_graph=True, fused=True).run_host(img, lab)
    finally:
        torch.randn_like = orig
    assert a[0] == b[0] == c[0]
    assert abs(a[1] - b[1]) < 1e-5 and a[1] == c[1]
    # static_weights: weight packing / BatchNorm folds once instead of per call -- same numbers; after the parameters
    # change, refresh_weights() brings the copies up to date
    torch.randn_like = _fake_noise()
    try:
        st = train.EvalStep(net, 16, 2, use_graph=True, fused=True, static_weights=True)
        d = st.run_host(img, lab)
        assert d == c
        with torch.no_grad():
            for p in net.parameters():
                p.mul_(1.01)
        st.refresh_weights()
        torch.randn_like = _fake_noise()
        e = st.run_host(img, lab)
        torch.randn_like = _fake_noise()
        f = train.EvalStep(net, 16, 2, use_graph=True, fused=True).run_host(img, lab)
    finally:
        torch.randn_like = orig
    assert e == f and e != d


def _rank_main(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from b200 import dp, train
    dp.init_from_env('nccl')
    net = _net()
    n, C, I = 10, 2, 2
    patch, labels, _ = synth.lidc_like_batch(I, seed=11)
    img, lab = patch[:, 0].contiguous().pin_memory(), labels.contiguous().pin_memory()
    counts = dp.shard_counts(n, world)
    lo = sum(counts[:rank])
    orig = torch.randn_like

    def sharded(t, **kw):
        # the noise the single-rank run gives to samples [lo, lo + n_local): full tensor [n*I, ...], rows sample-major
        full = (n * I,) + tuple(t.shape[1:])
        k = int(np.prod(full))
        v = torch.sin(torch.arange(k, device=t.device, dtype=torch.float32) * 12.9898).mul(1.7).reshape(full)
        return v[lo * I:(lo + counts[rank]) * I].contiguous()

    torch.randn_like = sharded
    try:
        ev = train.EvalStep(net, n, C, shard=(rank, world), images_per_step=I, use_graph=False, gather_results=True)
        got = ev.run_host(img, lab)
    finally:
        torch.randn_like = orig
    if rank == 0:
        torch.randn_like = _fake_noise()
        try:
            ref = train.EvalStep(net, n, C, images_per_step=I, use_graph=False).run_host(img, lab)
        finally:
            torch.randn_like = orig
        q.put((got.numpy(), ref.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (gpurun --gpus 2)')
def test_sharded_evaluation_is_bit_identical_to_single_rank():
    """2 ranks, the N samples split 5 + 5, masks all-gathered bit-packed, per-pixel sums all-reduced: GED must be bit
    identical to the single-rank evaluation of the same samples, NCC equal to fp32 summation order (1e-6), Dice equal."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, 29611, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, ref = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(got[:, 0], ref[:, 0]), (got, ref)
    assert np.allclose(got[:, 1], ref[:, 1], atol=1e-6, rtol=0)
    assert np.array_equal(got[:, 2:], ref[:, 2:])
