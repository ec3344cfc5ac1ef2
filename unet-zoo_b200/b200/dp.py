"""One-process-per-GPU data parallelism for the drop-in models (the reference has none).

Training: batch data-parallel.  Parameters are grouped into a few flat fp32 buckets in reverse registration
(~ reverse autograd) order; a post-accumulate hook counts arrivals and, when a bucket is complete, gathers its gradients
with one multi-tensor copy, re-points ``.grad`` at views of the flat buffer and issues ONE NCCL all-reduce (average).  torch.distributed's NCCL process group runs the collective on its own stream
after the producing kernels and ``finish()`` makes the compute stream wait for it, so communication overlaps the rest
of backward.  Parameters that never receive a gradient (``{posterior,prior}.upsampling_path.4.*``)
keep ``grad is None`` exactly like in the reference, so stock Adam skips them.

Evaluation: the N samples of an image are sharded over ranks; the per-rank class probabilities are exchanged with a
single all-gather and the GED / NCC kernels run on the gathered set.

BatchNorm uses per-rank statistics (each replica == the reference at its local batch size); running statistics stay
per rank (rank 0's are the ones saved).  The gloo backend is supported for the CPU tests of the host logic.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def shard_counts(n, world):
    """N=100 over 8 ranks -> 13,13,13,13,12,12,12,12."""
    base, rem = divmod(n, world)
    return [base + (1 if r < rem else 0) for r in range(world)]


def decorrelate_rank_seeds(rank, world, group=None):
    """Sample-sharded evaluation needs DIFFERENT noise on every rank (identical seeds would evaluate the same samples
    world times).  If two ranks report the same CUDA generator seed, every rank re-seeds with seed + 7919 * rank."""
    if world <= 1 or not dist.is_initialized() or not torch.cuda.is_available():
        return False
    seed = torch.cuda.initial_seed()
    t = torch.tensor([seed & 0x7FFFFFFFFFFF], dtype=torch.int64, device='cuda')
    allv = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allv, t, group=group)
    seeds = [int(v.item()) for v in allv]
    if len(set(seeds)) == world:
        return False
    torch.cuda.manual_seed(seed + 7919 * rank)
    return True


class _Bucket:
    __slots__ = ('flat', 'params', 'pending', 'work', 'views')

    def __init__(self, flat, params):
        self.flat, self.params, self.pending, self.work = flat, params, 0, None


class GradientAllReduce:
    """Bucketed, overlapped gradient averaging.  Usage per step:
         dp.zero_grad(); loss.backward(); dp.finish(); optimizer.step()
    The first steps (before ``freeze_buckets``) use a plain post-backward all-reduce and discover which parameters
    actually receive gradients."""

    def __init__(self, params, group=None, bucket_bytes=32 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.bucket_bytes = bucket_bytes
        self.buckets = None
        self._hooks = []
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    # -- discovery phase ------------------------------------------------------------------------------------------
    def _finish_unbucketed(self):
        grads = [p.grad for p in self.params if p.grad is not None]
        if not grads or self.world == 1:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.div_(self.world)
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    def freeze_buckets(self):
        """Call after at least one backward: builds the flat buckets over the parameters that have gradients."""
        live = [p for p in self.params if p.grad is not None]
        live.reverse()                                     # last-registered parameters finish backward first
        groups, cur, cur_bytes = [], [], 0
        for p in live:
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= self.bucket_bytes:
                groups.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            groups.append(cur)
        self.buckets = []
        for plist in groups:
            flat = torch.zeros(sum(p.numel() for p in plist), dtype=torch.float32, device=plist[0].device)
            b = _Bucket(flat, plist)
            b.views = []
            off = 0
            for p in plist:
                b.views.append(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(b)))
            self.buckets.append(b)
        self.zero_grad()

    def _make_hook(self, bucket):
        def hook(param):
            bucket.pending -= 1
            if bucket.pending == 0:
                # gather the bucket's freshly produced gradients into its flat buffer with ONE multi-tensor copy and
                # re-point .grad at the views: the all-reduce then averages the gradients in place
                if bucket.flat.is_cuda:
                    from . import ops
                    ops.sync_aux_streams()          # weight gradients are produced on the auxiliary stream
                torch._foreach_copy_(bucket.views, [p.grad for p in bucket.params])
                for p, v in zip(bucket.params, bucket.views):
                    p.grad = v
                if self.world > 1:
                    bucket.work = dist.all_reduce(bucket.flat, op=dist.ReduceOp.AVG if bucket.flat.is_cuda
                                                  else dist.ReduceOp.SUM, group=self.group, async_op=True)
        return hook

    # -- per step -------------------------------------------------------------------------------------------------
    def zero_grad(self):
        for p in self.params:
            p.grad = None
        if self.buckets is not None:
            for b in self.buckets:
                b.pending = len(b.params)
                b.work = None

    def finish(self):
        if self.buckets is None:
            self._finish_unbucketed()
            return
        for b in self.buckets:
            if b.work is not None:
                b.work.wait()                             # compute stream waits for the NCCL stream (no host sync on CUDA)
                if not b.flat.is_cuda:
                    b.flat.div_(self.world)               # gloo has no AVG
                b.work = None

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def gather_samples(local, counts, group=None):
    """all-gather per-rank sample tensors [n_r, ...] (n_r = counts[rank]) into [sum(counts), ...] on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    nmax = max(counts)
    pad = local
    if local.shape[0] < nmax:
        pad = torch.cat([local, local.new_zeros((nmax - local.shape[0],) + tuple(local.shape[1:]))])
    out = local.new_empty((world * nmax,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    parts = [out[r * nmax:r * nmax + counts[r]] for r in range(world)]
    return torch.cat(parts)
