"""Host logic of the data-parallel layer on CPU: two gloo ranks.  No CUDA kernels are involved: the
gradient bucketing / overlap hooks and the sample sharding are backend independent."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.gpu_util import PKG  # noqa: F401


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, PKG)
    from b200 import dp
    r, w, _ = dp.init_from_env('gloo')
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    unused = torch.nn.Parameter(torch.ones(3))                 # never receives a gradient (upsampling_path.4.* case)
    params = list(net.parameters()) + [unused]
    ar = dp.GradientAllReduce(params, bucket_bytes=256)        # tiny buckets -> several all-reduces
    results = []
    for step in range(3):
        torch.manual_seed(100 * step + rank)                   # every rank draws its own batch
        x = torch.randn(5, 8)
        ar.zero_grad()
        loss = net(x).pow(2).mean()
        loss.backward()
        ar.finish()
        # reference: average of the per-rank gradients computed independently
        ref = []
        for rr in range(world):
            torch.manual_seed(100 * step + rr)
            xr = torch.randn(5, 8)
            net2 = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
            net2.load_state_dict(net.state_dict())
            net2(xr).pow(2).mean().backward()
            ref.append([p.grad.clone() for p in net2.parameters()])
        avg = [sum(g[i] for g in ref) / world for i in range(len(ref[0]))]
        ok = all(torch.allclose(p.grad, a, rtol=1e-5, atol=1e-7) for p, a in zip(net.parameters(), avg))
        results.append(ok and unused.grad is None)
        if step == 0:
            ar.freeze_buckets()                                # discovery step done -> overlapped buckets from now on
    assert len(ar.buckets) > 1
    # sample sharding: N = 5 over 2 ranks -> 3 + 2, gathered in rank order
    counts = dp.shard_counts(5, world)
    local = torch.full((counts[rank], 2), float(rank))
    gathered = dp.gather_samples(local, counts)
    results.append(tuple(gathered[:, 0].tolist()) == (0.0, 0.0, 0.0, 1.0, 1.0))
    q.put((rank, results))
    dist.destroy_process_group()


def test_gradient_allreduce_and_sample_sharding_two_ranks():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29731
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, results in out:
        assert all(results), (rank, results)


def test_shard_counts_match_survey():
    sys.path.insert(0, PKG)
    from b200 import dp
    assert dp.shard_counts(100, 8) == [13, 13, 13, 13, 12, 12, 12, 12]
    assert dp.shard_counts(100, 1) == [100]
    assert sum(dp.shard_counts(16, 3)) == 16


def test_buckets_are_cut_from_the_end_of_the_production_order():
    """single process, no process group: the bucket layout only depends on the order the gradients were produced in
    (recorded by the hooks) and on tail_bytes / bucket_bytes -- tail first from the END, doubling up to the cap, production
    order preserved inside and across buckets; parameters without a gradient stay outside"""
    sys.path.insert(0, PKG)
    from b200 import dp
    sizes = [40, 10, 30, 20, 50, 5, 60, 15]                    # floats; forward order
    params = [torch.nn.Parameter(torch.randn(n)) for n in sizes]
    unused = torch.nn.Parameter(torch.ones(7))
    ar = dp.GradientAllReduce(params + [unused], bucket_bytes=100 * 4, tail_bytes=20 * 4)
    ar.zero_grad()
    loss = sum((k + 1) * p.sum() for k, p in enumerate(params))
    loss.backward()                                            # one backward: the hooks record the production order
    ar.finish()
    produced = list(ar._order)
    assert len(produced) == len(params)
    ar.freeze_buckets()
    flat_order = [id(p) for b in ar.buckets for p in b.params]
    assert flat_order == produced                              # production order, nothing lost, nothing duplicated
    assert all(id(unused) != pid for pid in flat_order) and unused.grad is None
    by_id = {id(p): p.numel() for p in params}
    bucket_floats = [sum(by_id[id(p)] for p in b.params) for b in ar.buckets]
    # counted from the end: the last bucket holds at most tail_bytes (or one oversized parameter), the limits double
    limit = 20
    for nfl, b in zip(reversed(bucket_floats), reversed(ar.buckets)):
        assert nfl <= limit or len(b.params) == 1, (bucket_floats, limit)
        limit = min(2 * limit, 100)
    # a second step through the frozen buckets: gradients land in the flat buffers (views), values unchanged
    ar.zero_grad()
    sum((k + 1) * p.sum() for k, p in enumerate(params)).backward()
    ar.finish()
    for k, p in enumerate(params):
        assert torch.equal(p.grad, torch.full_like(p, float(k + 1)))
        assert p.grad.data_ptr() == dp.grad_view_for(p).data_ptr()
    ar.remove()
