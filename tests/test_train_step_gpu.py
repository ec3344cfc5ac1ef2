"""b200.train.TrainStep: the captured training step (reference train_model.py:101-122) with the parameter update
pipelined into backward (per-bucket Adam + weight re-pack on a side stream) against the same step with the update after
backward.  Deterministic statistics mode, so the two schedules must agree BIT FOR BIT: every parameter is updated
exactly once per step from its complete gradient in both."""
import numpy as np
import pytest
import torch

from tests.gpu_util import kern

pytestmark = pytest.mark.gpu

FILTERS = [16, 32, 32, 32, 32, 32, 32]
B = 4


def _make(overlap, use_graph=True):
    from b200 import build, synth, train
    net = build.phiseg(FILTERS)
    net.load_state_dict(synth.synth_state_dict(net.state_dict(), seed=1))
    net = net.cuda()
    st = train.TrainStep(net, train.make_adam(net), B, (1, 128, 128), use_graph=use_graph, overlap_optimizer=overlap)
    return net, st


def _run(st, batches, seed=1234):
    torch.cuda.manual_seed(seed)
    losses = []
    for patch, mask in batches:
        st.patch.copy_(patch)
        st.mask.copy_(mask)
        st.step_device()
        losses.append(float(st.loss))
    torch.cuda.synchronize()
    return losses


def _batches(n):
    from b200 import synth
    out = []
    for i in range(n):
        patch, _, mask = synth.lidc_like_batch(B, seed=10 + i)
        out.append((patch.cuda(), mask.cuda()))
    return out


def test_pipelined_optimizer_matches_update_after_backward():
    k = kern()
    prev = k.set_deterministic(True)
    try:
        net_a, st_a = _make(True)
        net_b, st_b = _make(False)
        assert st_a.dp is not None and st_b.dp is None
        batches = _batches(3)
        st_a.patch.copy_(batches[0][0]); st_a.mask.copy_(batches[0][1])
        st_b.patch.copy_(batches[0][0]); st_b.mask.copy_(batches[0][1])
        st_a.prepare(warmup=2)
        st_b.prepare(warmup=2)
        assert st_a.dp.owns_optimizer and len(st_a.dp.buckets) >= 2
        # prepare() restores the model / optimizer state: both start from the initial weights again
        for (n, p), (_, q) in zip(net_a.named_parameters(), net_b.named_parameters()):
            assert torch.equal(p, q), n
        la = _run(st_a, batches)
        lb = _run(st_b, batches)
        assert la == lb, (la, lb)
        assert all(np.isfinite(la)) and len(set(la)) == 3
        for (n, p), (_, q) in zip(net_a.named_parameters(), net_b.named_parameters()):
            assert torch.equal(p, q), n
        for (n, p), (_, q) in zip(net_a.named_buffers(), net_b.named_buffers()):
            assert torch.equal(p, q), n
        sa, sb = st_a.opt.state, st_b.opt.state
        for p, q in zip(net_a.parameters(), net_b.parameters()):
            if p in sa or q in sb:
                for key in ('exp_avg', 'exp_avg_sq', 'step'):
                    assert torch.equal(sa[p][key], sb[q][key]), key
        # fewer kernels on the step's critical path: no pack launch at the start, no Adam launch at the end
        assert st_a.launches_per_step >= st_b.launches_per_step          # same work, split into per-bucket launches

        # parameters changed from outside: the load_state_dict hook re-packs the tensor-core copies
        from b200 import synth
        init = synth.synth_state_dict(net_a.state_dict(), seed=2)
        net_a.load_state_dict(init)
        net_b.load_state_dict(init)
        st_a.opt.reset_state()
        st_b.opt.reset_state()
        la = _run(st_a, batches[:1], seed=7)
        lb = _run(st_b, batches[:1], seed=7)
        assert la == lb
        for (n, p), (_, q) in zip(net_a.named_parameters(), net_b.named_parameters()):
            assert torch.equal(p, q), n
    finally:
        k.set_deterministic(prev)
