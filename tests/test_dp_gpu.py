"""Data-parallel training on two GPUs (NCCL): the gradients after the bucketed, overlapped all-reduce must equal the mean
of the gradients each rank computes on its own batch, including the weight gradients that the kernels write
straight into the flat buckets.  Needs two devices: `gpurun --gpus 2 -- python -m pytest tests/test_dp_gpu.py -m gpu`."""
import os

import numpy as np
import pytest
import torch

from oracle import synth
from tests.keygrammar import dropin_phiseg

pytestmark = pytest.mark.gpu
FILTERS = [16, 32, 32, 32, 32, 32, 32]
B = 4


def _grads(net, rank_seed, ar=None):
    from oracle.ref_run import injected_noise
    patch, _, mask = synth.lidc_like_batch(B, seed=30 + rank_seed)
    eps = synth.noise_list(synth.phiseg_noise_shapes(B), seed=40 + rank_seed)
    if ar is not None:
        ar.zero_grad()
    else:
        for p in net.parameters():
            p.grad = None
    with injected_noise(eps):
        net.forward(patch.cuda(), mask.cuda(), training=True)
        loss = net.loss(mask.cuda())
    loss.backward()
    if ar is not None:
        ar.finish()
    torch.cuda.synchronize()
    return {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}


def _rank_main(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from b200 import dp, kern
    dp.init_from_env('nccl')
    kern.set_deterministic(True)               # bit-reproducible kernels: the only difference left is the all-reduce
    net = dropin_phiseg(FILTERS)
    sd = synth.synth_state_dict(net.state_dict(), seed=1)
    net.load_state_dict(sd)
    net = net.cuda().train()
    bn = {n: b.detach().clone() for n, b in net.named_buffers()}

    def reset_bn():
        for n, b in net.named_buffers():
            b.copy_(bn[n])

    ar = dp.GradientAllReduce(net.parameters(), bucket_bytes=1 << 20, tail_bytes=64 << 10)
    _grads(net, rank, ar)                      # discovery step (plain all-reduce), records the production order
    ar.freeze_buckets()
    reset_bn()
    got = _grads(net, rank, ar)                # bucketed + overlapped, conv weight gradients produced inside the buckets
    in_place = sum(1 for b in ar.buckets for p, v in zip(b.params, b.views) if p.grad is not None and
                   p.grad.data_ptr() == v.data_ptr())
    nb = len(ar.buckets)
    ar.remove()
    if rank == 0:
        per_rank = []
        for r in range(world):
            reset_bn()
            per_rank.append(_grads(net, r))
        q.put(({k: v.cpu().numpy() for k, v in got.items()},
               [{k: v.cpu().numpy() for k, v in g.items()} for g in per_rank], nb, in_place))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (gpurun --gpus 2)')
def test_bucketed_allreduce_equals_mean_of_rank_gradients():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, 29641, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, per_rank, nb, in_place = q.get(timeout=600)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert nb >= 2 and in_place == len(got)
    worst = 0.0
    for n, g in got.items():
        mean = (per_rank[0][n] + per_rank[1][n]) / 2
        den = float(np.abs(mean).max())
        if den == 0:
            assert float(np.abs(g).max()) == 0
            continue
        worst = max(worst, float(np.abs(g - mean).max()) / den)
    assert worst < 1e-5, worst                 # fp32 rounding of (a + b) / 2 only
