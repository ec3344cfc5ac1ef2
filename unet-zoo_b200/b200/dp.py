"""One-process-per-GPU data parallelism for the drop-in models (the reference has none).

Training: batch data-parallel.  Parameters are grouped into a few flat fp32 buckets in the order their gradients are
produced (recorded during the warm-up steps), cut from the end of backward (``tail_bytes`` for the last arrivals, doubling
up to ``bucket_bytes``; by default both are 24 MB: few collectives beat a short last one); the tensor-core weight
gradients are written straight into the buckets (``grad_view_for``), the few remaining gradients are gathered with one
multi-tensor copy, ``.grad`` is re-pointed at views of the flat buffer and ONE NCCL all-reduce (average) per bucket is
issued from a side stream.  ``finish()`` makes the compute stream wait, so communication overlaps the rest of backward.  Parameters that never receive a gradient (``{posterior,prior}.upsampling_path.4.*``)
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
    __slots__ = ('flat', 'params', 'pending', 'work', 'views', 'streams', 'index')

    def __init__(self, flat, params):
        self.flat, self.params, self.pending, self.work = flat, params, 0, None
        self.streams = {}          # compute streams the gradients of this bucket were accumulated on (this step)
        self.index = 0


# parameter -> fp32 view into its flat bucket.  b200.ops hands the view to the weight-gradient kernel as its output, so
# the gradient is PRODUCED inside the bucket and ``.grad`` becomes that view without a gather copy.
_grad_views = {}


def grad_view_for(param):
    return _grad_views.get(id(param))


class GradientAllReduce:
    """Bucketed, overlapped gradient averaging.  Usage per step:
         dp.zero_grad(); loss.backward(); dp.finish(); optimizer.step()
    The first steps (before ``freeze_buckets``) use a plain post-backward all-reduce and record the ORDER in which the
    gradients are produced; ``freeze_buckets`` then builds the flat buckets in that order, counted from the END of
    backward (``tail_bytes`` for the parameters whose gradients arrive last, doubling up to ``bucket_bytes``).  A completed bucket is handed to NCCL from a side stream that waits for the producing
    streams (the compute stream is never stalled by communication set-up); ``finish()`` joins.

    ``optimizer`` (a b200.optim.FusedAdam) moves the parameter update into the bucket pipeline as well: as soon as a
    bucket is complete (and averaged), its parameters are stepped (``optimizer.step_params``) and -- with ``packer``, the
    model's kern.WeightPacker -- their bf16 tensor-core copies re-packed, on the side stream, next to the rest of backward.
    Only the last (small) bucket's update remains on the critical path; the per-forward weight packing disappears.  The
    caller then skips ``optimizer.step()`` (``owns_optimizer``).  Works with a single rank (no process group) too."""

    def __init__(self, params, group=None, bucket_bytes=24 << 20, tail_bytes=24 << 20, optimizer=None, packer=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.optimizer = optimizer
        self.packer = packer
        self.bucket_bytes = int(os.environ.get('UNETZOO_DP_BUCKET_MB', 0)) << 20 or bucket_bytes
        self.tail_bytes = int(float(os.environ.get('UNETZOO_DP_TAIL_MB', 0)) * (1 << 20)) or tail_bytes
        self.buckets = None
        self._hooks = []
        self._order = []
        self._comm_stream = None
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        for p in self.params:                              # discovery: the order gradients become available in
            self._hooks.append(p.register_post_accumulate_grad_hook(self._record_order))

    def _record_order(self, param):
        if self.buckets is None:
            self._order.append(id(param))

    @property
    def owns_optimizer(self):
        """True once the buckets exist and the parameter update runs inside the bucket pipeline"""
        return self.optimizer is not None and self.buckets is not None

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
        seen = {}
        for k, pid in enumerate(self._order):              # last recorded backward wins
            seen[pid] = k
        if len(seen) >= len(live):
            live.sort(key=lambda p: seen.get(id(p), -1))    # production order of the last discovery step
        else:
            live.reverse()                                  # no hook information: last-registered parameters finish first
        # Buckets are cut from the END of the production order with doubling sizes (tail_bytes, 2x, 4x, ... capped at
        # bucket_bytes).  Measured at N = 2 (PHiSeg-7/5, ms per step): tail 0.5 / 2 / 8 / 24 MB -> 4.353 / 4.286 / 4.290 /
        # 4.250: every collective costs ~30 us of latency plus a gather launch on the communication stream, so FEW buckets
        # win over a short last one -- the default tail equals the bucket size (uniform 24 MB buckets counted from the end;
        # the parameter-heavy deep encoder levels finish ~0.8 ms before the end of backward).  A small tail remains useful
        # with ``optimizer`` (the last update / re-pack cannot hide behind backward).
        groups, cur, cur_bytes = [], [], 0
        limit = min(self.tail_bytes, self.bucket_bytes)
        for p in reversed(live):
            nbytes = p.numel() * 4
            if cur and cur_bytes + nbytes > limit:
                groups.insert(0, cur)
                cur, cur_bytes = [], 0
                limit = min(limit * 2, self.bucket_bytes)
            cur.insert(0, p)
            cur_bytes += nbytes
        if cur:
            groups.insert(0, cur)
        for h in self._hooks:
            h.remove()
        self._hooks = []
        self.buckets = []
        for plist in groups:
            flat = torch.zeros(sum(p.numel() for p in plist), dtype=torch.float32, device=plist[0].device)
            b = _Bucket(flat, plist)
            b.views = []
            off = 0
            for p in plist:
                b.views.append(flat[off:off + p.numel()].view_as(p))
                _grad_views[id(p)] = b.views[-1]
                off += p.numel()
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(b)))
            b.index = len(self.buckets)
            self.buckets.append(b)
            if self.packer is not None:
                self.packer.make_subset(('bucket', b.index), plist)
        if live and live[0].is_cuda:
            self._comm_stream = torch.cuda.Stream(device=live[0].device)
        if self.packer is not None and self.optimizer is not None:
            # from now on the packed copies are refreshed per bucket right after the update: bring them up to date once
            self.packer.refresh(force=True)
            self.packer.external = True
        self.zero_grad()

    def _make_hook(self, bucket):
        def hook(param):
            if param.is_cuda:
                st = torch.cuda.current_stream()           # the accumulation stream has waited for the producing kernels
                bucket.streams[st.cuda_stream] = st
            bucket.pending -= 1
            if bucket.pending == 0:
                self._launch(bucket)
        return hook

    def _launch(self, bucket):
        """bucket complete: gather what was not produced in place, then ONE all-reduce (average)"""
        cuda = bucket.flat.is_cuda
        if cuda:
            from . import ops
            cur = torch.cuda.current_stream()
            side = self._comm_stream
            side.wait_stream(cur)
            for st in bucket.streams.values():             # every stream a gradient of this bucket was accumulated on: the
                side.wait_stream(st)                       # layer's dgrad (reads the packed weights) precedes it there
            for st in ops.pending_aux_streams():           # weight gradients are produced on the auxiliary streams
                side.wait_stream(st)
            ctx = torch.cuda.stream(side)
        else:
            ctx = _NullCtx()
        with ctx:
            if cuda:
                from . import kern
                kern.wgrad_reducer.flush()                 # deferred split-K reductions of the weight gradients so far
            src, dst = [], []
            for p, v in zip(bucket.params, bucket.views):
                if p.grad.data_ptr() != v.data_ptr():      # conv weight gradients already live in the bucket
                    src.append(p.grad)
                    dst.append(v)
            if dst:
                torch._foreach_copy_(dst, src)
            for p, v in zip(bucket.params, bucket.views):
                p.grad = v
            if self.world > 1:
                bucket.work = dist.all_reduce(bucket.flat, op=dist.ReduceOp.AVG if cuda else dist.ReduceOp.SUM,
                                              group=self.group, async_op=True)
            elif cuda:
                bucket.work = side                         # single rank: only the gather has to be joined
            if self.optimizer is not None:
                if self.world > 1:
                    bucket.work.wait()                     # the side stream waits for the average (no host sync on CUDA)
                    if not cuda:
                        bucket.flat.div_(self.world)       # gloo has no AVG
                self.optimizer.step_params(bucket.params, ('bucket', bucket.index))
                if self.packer is not None:
                    self.packer.refresh_subset(('bucket', bucket.index))
                bucket.work = side if cuda else None

    # -- per step -------------------------------------------------------------------------------------------------
    def zero_grad(self):
        for p in self.params:
            p.grad = None
        if self.buckets is not None:
            for b in self.buckets:
                b.pending = len(b.params)
                b.work = None
                b.streams = {}

    def finish(self):
        if self.buckets is None:
            self._finish_unbucketed()
            return
        for b in self.buckets:
            if b.work is not None:
                if isinstance(b.work, torch.cuda.Stream):
                    torch.cuda.current_stream().wait_stream(b.work)
                else:
                    b.work.wait()                         # compute stream waits for the NCCL stream (no host sync on CUDA)
                    if b.flat.is_cuda:
                        torch.cuda.current_stream().wait_stream(self._comm_stream)
                    else:
                        b.flat.div_(self.world)           # gloo has no AVG
                b.work = None

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
        for b in self.buckets or []:
            for p in b.params:
                _grad_views.pop(id(p), None)


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


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
