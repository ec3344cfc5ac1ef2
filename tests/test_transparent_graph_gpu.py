"""Transparent CUDA-graph capture behind the module API (UNETZOO_TRANSPARENT_GRAPH / net.transparent_graph): the training
loop of the UNMODIFIED caller (reference train_model.py:100-134: forward -> loss -> zero_grad -> backward -> stock
torch.optim.Adam.step -> scheduler.step(loss)) replayed statement by statement, once eagerly and once with the capture
switched on.  With deterministic kernels and identical noise the two runs must agree bit for bit: the captured graphs ARE
the eager step."""
import time

import numpy as np
import pytest
import torch

from oracle import synth
from tests.keygrammar import dropin_phiseg

pytestmark = pytest.mark.gpu
FILTERS = [16, 32, 32, 32, 32, 32, 32]
B = 4


def _fake_noise(t, **kw):
    n = t.numel()
    return torch.sin(torch.arange(n, device=t.device, dtype=torch.float32) * 12.9898).mul(1.7).reshape(t.shape)


def _caller_loop(transparent, iterations, filters=FILTERS, batch=B, timing=False, fused_adam=False):
    from b200 import kern
    net = dropin_phiseg(filters)
    net.load_state_dict(synth.synth_state_dict(net.state_dict(), seed=1))
    net = net.cuda()
    net.transparent_graph = transparent
    # ---- UNetModel.__init__ (train_model.py:49-51)
    if fused_adam:
        from b200.optim import FusedAdam           # what launch.py --graph installs as torch.optim.Adam
        optimizer = FusedAdam(net.parameters(), lr=1e-3, weight_decay=1e-5)
    else:
        optimizer = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-5)
    scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, 'min', min_lr=1e-4, patience=100)
    net.train()
    losses, kls = [], []
    orig = torch.randn_like
    torch.randn_like = _fake_noise
    prev = kern.set_deterministic(not timing)
    from b200 import ops
    saved_share = ops._WGRAD_SM_PERCENT
    if not timing:
        # same split-K plan captured and eager (a captured step normally plans its overlapped weight gradients for a quarter
        # of the SMs: other partial sums, other fp32 rounding) -> the two runs can be compared bit for bit
        ops._WGRAD_SM_PERCENT = 100
        ops._aux['planned_for'] = None
    t0 = None
    data = [synth.lidc_like_batch(batch, seed=100 + k) for k in range(4)]       # data.train.next_batch: host numpy arrays
    data = [(x.numpy(), s.numpy()[:, 0]) for x, _, s in data]
    try:
        for iteration in range(1, iterations + 1):
            if timing and iteration == 6:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            x_b, s_b = data[iteration % 4]
            # ---- train_model.py:103-122
            patch = torch.tensor(x_b, dtype=torch.float32).to('cuda')
            mask = torch.tensor(s_b, dtype=torch.float32).to('cuda')
            mask = torch.unsqueeze(mask, 1)
            net.forward(patch, mask, training=True)
            loss = net.loss(mask)
            reconstruction_loss, kl_loss = net.reconstruction_loss, net.kl_divergence_loss
            assert kl_loss is loss and reconstruction_loss is loss            # quirk Q1
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            scheduler.step(loss)                                             # float(loss): the per-iteration sync (:134)
            if not timing:
                losses.append(float(loss))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) if t0 is not None else None
    finally:
        torch.randn_like = orig
        kern.set_deterministic(prev)
        ops._WGRAD_SM_PERCENT = saved_share
        ops._aux['planned_for'] = None
    params = {n: p.detach().clone() for n, p in net.named_parameters()}
    stats = {n: b.detach().clone() for n, b in net.named_buffers()}
    return losses, params, stats, dt


def test_transparent_capture_equals_eager_loop():
    l_e, p_e, s_e, _ = _caller_loop(False, 7)
    l_t, p_t, s_t, _ = _caller_loop(True, 7)          # iterations 1-2 eager warm-up, capture at 3, replays from 3 on
    print('\nlosses eager       %s\nlosses transparent %s' % (l_e, l_t))
    assert l_e == l_t
    for n in p_e:
        assert torch.equal(p_e[n], p_t[n]), n
    for n in s_e:
        assert torch.equal(s_e[n], s_t[n]), n


def test_transparent_capture_survives_validation_and_is_faster():
    """eval-mode forwards in between (validate()) do not disturb the captured training step; and the point of it all:
    the loop is several times faster than the eager one at the benchmark configuration."""
    from tests.test_parity_conditioned_gpu import FILTERS as BIG
    net = dropin_phiseg(FILTERS)
    net.load_state_dict(synth.synth_state_dict(net.state_dict(), seed=1))
    net = net.cuda().train()
    net.transparent_graph = True
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-5)
    x_b, _, s_b = synth.lidc_like_batch(B, seed=100)
    patch, mask = x_b.cuda(), s_b.cuda()
    for it in range(5):
        net.forward(patch, mask, training=True)
        loss = net.loss(mask)
        opt.zero_grad()
        loss.backward()
        opt.step()
        if it == 3:
            net.eval()
            with torch.no_grad():
                s = net.forward(patch.repeat(2, 1, 1, 1), mask.repeat(2, 1, 1, 1), training=False)
                probs = net.accumulate_output(s, use_softmax=True)
                assert torch.isfinite(probs).all()
            net.train()
    assert np.isfinite(float(loss))
    _, _, _, t_eager = _caller_loop(False, 25, filters=BIG, batch=12, timing=True)
    _, _, _, t_graph = _caller_loop(True, 25, filters=BIG, batch=12, timing=True)
    _, _, _, t_fused = _caller_loop(True, 25, filters=BIG, batch=12, timing=True, fused_adam=True)
    ips_e, ips_g, ips_f = 20 * 12 / t_eager, 20 * 12 / t_graph, 20 * 12 / t_fused
    print('\nunmodified-caller loop, PHiSeg-7/5 B=12: eager + stock Adam %.0f images/s; transparent graph + stock Adam %.0f; '
          'transparent graph + FusedAdam (launch.py --graph) %.0f' % (ips_e, ips_g, ips_f))
    assert ips_g > 1.5 * ips_e and ips_f > 3.0 * ips_e
