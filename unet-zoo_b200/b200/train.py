"""Step drivers around the drop-in modules: the reference's training step (train_model.py:101-122) and its
N-sample evaluation (train_model.py:177-205) as reusable objects, optionally captured in a CUDA graph so the ~1.5 k
kernel launches of a PHiSeg step are replayed without Python in the loop.

The caller-visible semantics are the reference's: stock torch.optim.Adam(lr=1e-3, weight_decay=1e-5) on the fp32
parameters, loss = net.loss(mask) after net.forward(patch, mask, training=True).
"""
import os

import torch

from . import _lib, kern


_CONSUME_SLABS = os.environ.get('UNETZOO_ADAM_FROM_SLABS', '1') != '0'


def make_adam(net, capturable=True, fused=True, own_kernel=True):
    """The reference's optimizer (train_model.py:49: Adam, lr 1e-3, weight_decay 1e-5 as L2 on the gradient).
    own_kernel=True: b200.optim.FusedAdam, one launch for all parameters (graph capturable, same state layout as
    torch.optim.Adam).  own_kernel=False: stock torch.optim.Adam; capturable=True lets optimizer.step() live inside a CUDA
    graph, fused=True selects torch's multi-tensor implementation (its default for-each path launches two tiny pow
    kernels PER PARAMETER for the bias corrections: 1640 launches per PHiSeg step)."""
    if own_kernel:
        from .optim import FusedAdam
        return FusedAdam(net.parameters(), lr=1e-3, weight_decay=1e-5)
    return torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-5, capturable=capturable, fused=fused)


class TrainStep:
    """forward(training=True) -> loss -> zero_grad -> backward -> [gradient all-reduce] -> Adam.step.

    With the library's FusedAdam the parameter update can be pipelined with backward (``overlap_optimizer`` /
    UNETZOO_OVERLAP_OPT=1, opt-in): the parameters are grouped into gradient buckets in production order (the
    data-parallel machinery of b200.dp, also on a single GPU) and every completed bucket is [averaged,] stepped and its
    bf16 tensor-core weight copies re-packed on a side stream while the rest of backward runs -- Adam and the weight
    packing leave the critical path.  Each parameter is still updated exactly once per step from its complete gradient.
    After changing parameters from outside (load_state_dict, manual edits) call ``refresh_weights()``."""

    def __init__(self, net, optimizer, batch, image_size=(1, 128, 128), use_graph=True, dp=None, device=None,
                 overlap_optimizer=None):
        self.net, self.opt, self.dp = net, optimizer, dp
        self.device = device or torch.device('cuda', torch.cuda.current_device())
        if overlap_optimizer is None:
            # measured on B200 (PHiSeg-7/5, B=12, one GPU): 4.64 ms pipelined vs 4.60 ms with the update after backward --
            # backward is throughput-bound, the update only competes with it -- hence opt-in
            overlap_optimizer = os.environ.get('UNETZOO_OVERLAP_OPT', '0') == '1'
        from .optim import FusedAdam
        self.packer = None
        if isinstance(optimizer, FusedAdam) and hasattr(net, '_packer') and not overlap_optimizer and \
                os.environ.get('UNETZOO_ADAM_PACK', '1') != '0':
            # the optimizer re-packs the bf16 tensor-core weight copies while it updates them (uz_adam_pack_step): no
            # packing pass at the head of the step
            self.packer = net._packer()
            optimizer.attach_packer(self.packer)
            net.register_load_state_dict_post_hook(lambda module, incompatible: self.refresh_weights())
        if overlap_optimizer and isinstance(optimizer, FusedAdam) and len(optimizer.param_groups) == 1:
            self.packer = net._packer() if hasattr(net, '_packer') else None
            if self.dp is None:
                from . import dp as _dp
                self.dp = _dp.GradientAllReduce(net.parameters(), optimizer=optimizer, packer=self.packer,
                                                tail_bytes=2 << 20)    # the last update cannot hide: keep it short
            elif self.dp.optimizer is None and self.dp.buckets is None:
                self.dp.optimizer, self.dp.packer = optimizer, self.packer
            if self.packer is not None:
                net.register_load_state_dict_post_hook(lambda module, incompatible: self.refresh_weights())
        c, spatial = image_size[0], tuple(image_size[1:])          # (C,H,W) images or (C,D,H,W) volumes
        self.patch = torch.zeros((batch, c) + spatial, dtype=torch.float32, device=self.device)
        self.mask = torch.zeros((batch, 1) + spatial, dtype=torch.float32, device=self.device)
        self.loss = torch.zeros((), dtype=torch.float32, device=self.device)
        self.loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
        self.graph = None
        self.launches_per_step = None
        self.use_graph = use_graph
        net.train()

    def _body(self):
        if self.dp is not None:
            self.dp.zero_grad()
        else:
            self.opt.zero_grad(set_to_none=True)
        self.net.forward(self.patch, self.mask, training=True)
        loss = self.net.loss(self.mask)
        # single GPU, fused optimizer with an attached packer: the optimizer reads the weight-gradient slabs itself
        # (no reduction pass, no OIHW gradient tensors -- nobody else looks at .grad inside this step)
        hold = self.dp is None and self.packer is not None and getattr(self.opt, 'packer', None) is self.packer and \
            _CONSUME_SLABS
        kern.wgrad_reducer.hold = hold
        try:
            loss.backward()
            if self.dp is not None:
                self.dp.finish()
            if self.dp is None or not self.dp.owns_optimizer:
                self.opt.step()
        finally:
            kern.wgrad_reducer.hold = False
            kern.wgrad_reducer.flush()
        self.loss.copy_(loss.detach())

    def refresh_weights(self):
        """re-pack the bf16 tensor-core copies from the fp32 parameters (needed after the parameters were changed from
        outside when the optimizer pipeline keeps the copies current, see the class docstring)"""
        if self.packer is not None:
            self.packer.refresh(force=True)

    def _snapshot(self):
        """parameters, buffers (BatchNorm running statistics, num_batches_tracked) and optimizer state before the warm-up"""
        net = {k: v.detach().clone() for k, v in self.net.state_dict().items()}
        opt = {}
        for p, st in self.opt.state.items():
            opt[p] = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
        return net, opt

    def _restore(self, snap):
        """undo what the warm-up / capture steps did to the model and the optimizer -- IN PLACE, so every pointer a
        captured graph or a descriptor table holds stays valid.  Optimizer state that did not exist before (first use)
        is zeroed: the first replayed step is then step 1 from the initial / checkpoint state, like the reference loop."""
        net, opt = snap
        with torch.no_grad():
            for k, v in self.net.state_dict().items():
                v.copy_(net[k])
            for p, st in self.opt.state.items():
                prev = opt.get(p)
                for k, v in st.items():
                    if not torch.is_tensor(v):
                        if prev is not None and k in prev:
                            st[k] = prev[k]
                        continue
                    if prev is not None and k in prev and torch.is_tensor(prev[k]):
                        v.copy_(prev[k])
                    else:
                        v.zero_()
        self.refresh_weights()

    def prepare(self, warmup=3, keep_state=True):
        """eager warm-up on a side stream (also sizes every lazily-set kernel attribute), then capture.  The warm-up
        and capture steps are real optimizer steps on the zero-filled input buffers; with ``keep_state`` (default)
        the model, its BatchNorm statistics and the optimizer state are restored afterwards, so training starts from
        the state the caller handed in."""
        snap = self._snapshot() if keep_state else None
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if self.dp is not None:
            self.dp.freeze_buckets()
            if self.dp.owns_optimizer:
                # one eager step in the pipelined configuration: sizes the per-bucket descriptor tables outside a capture
                with torch.cuda.stream(s):
                    self._body()
                torch.cuda.current_stream().wait_stream(s)
                torch.cuda.synchronize()
        if self.use_graph:
            self.graph = torch.cuda.CUDAGraph()
            n0 = _lib.raw('uz_launch_count')()
            # the capture stream (and the side streams forked from it for the forward / dgrad chains) get a higher
            # priority than the auxiliary weight-gradient streams: critical-path kernels are scheduled first
            prio = int(os.environ.get('UNETZOO_MAIN_PRIORITY', '-2'))
            cap_stream = torch.cuda.Stream(device=self.device, priority=prio)
            mark = len(kern.wgrad_reducer.keep)
            with torch.cuda.graph(self.graph, stream=cap_stream):
                self._body()
            # the weight-gradient slabs the captured launches point at live (and die) with this object, not with the
            # process-wide reducer
            self._slabs = kern.wgrad_reducer.keep[mark:]
            del kern.wgrad_reducer.keep[mark:]
            self.launches_per_step = _lib.raw('uz_launch_count')() - n0
            if hasattr(self.opt, 'finish_capture'):
                self.opt.finish_capture()
        else:
            n0 = _lib.raw('uz_launch_count')()
            self._body()
            self.launches_per_step = _lib.raw('uz_launch_count')() - n0
        torch.cuda.synchronize()
        if snap is not None:
            self._restore(snap)
            torch.cuda.synchronize()

    def step_device(self):
        """inputs already resident in self.patch / self.mask"""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()

    def step_host(self, patch_pinned, mask_pinned):
        """public end-to-end call: pinned host batch in, python float loss out (H2D + step + D2H)."""
        self.patch.copy_(patch_pinned, non_blocking=True)
        self.mask.copy_(mask_pinned, non_blocking=True)
        self.step_device()
        self.loss_host.copy_(self.loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self.loss_host)


class EvalStep:
    """N-sample evaluation like train_model.py:177-222 for I images per call: N copies of every image ->
    forward(training=False) -> accumulate_output(softmax) -> argmax -> GED + NCC against M annotators + Dice of the mean
    prediction, as ONE fused tail kernel over the low-resolution level logits (uz_eval_sample_stats: accumulate, softmax,
    argmax -> bit-packed masks, per-pixel sum p / sum log p).

    ``shard`` = (rank, world) splits the N samples of every image over the ranks (13,13,13,13,12,12,12,12 for N=100 on 8
    GPUs).  The exchange is what the survey specifies: ONE all-gather of the bit-packed masks (+ their
    pixel counts; 2 KB per sample) and ONE all-reduce of the per-pixel sums (2*C*H*W floats per image); the metrics of
    image i are then computed by rank i % world only (no redundant work; results stay on the owning rank unless
    ``gather_results``).  ``images_per_step`` > 1 keeps every GPU busy when N / world is small: the encoders run on I
    images, the latent / likelihood path on I * N / world samples.  Ranks must draw DIFFERENT noise: the constructor
    offsets the CUDA generator by the rank when all ranks were seeded alike.

    ``run_host`` replays a CUDA graph of the whole evaluation; ``fused=False`` is the reference-shaped path through the
    module API and the drop-in ``utils`` functions (full-resolution logits, softmax tensor, argmax)."""

    def __init__(self, net, n_samples=100, n_classes=2, shard=None, dedup=True, use_graph=True, images_per_step=1,
                 fused=True, gather_results=False, static_weights=False):
        self.net = net
        # static_weights: the parameters do not change between calls (one validation pass over many images,
        # train_model.py:150-222): the bf16 weight packing and the BatchNorm folds run ONCE (``refresh_weights()``, also
        # called by the first run) instead of inside every call; call ``refresh_weights()`` again after an optimizer step
        self.static_weights = static_weights and hasattr(net, '_packer') and hasattr(net, '_folder')
        self.dedup = dedup
        self.use_graph = use_graph
        self.fused = fused and dedup
        self.gather_results = gather_results
        self.n = n_samples
        self.n_classes = n_classes
        self.images = images_per_step
        self.counts = None
        self.rank, self.world = 0, 1
        self.n_local = n_samples
        if shard is not None and shard[1] > 1:
            from . import dp
            self.rank, self.world = shard
            self.counts = dp.shard_counts(n_samples, self.world)
            self.n_local = self.counts[self.rank]
            dp.decorrelate_rank_seeds(self.rank, self.world)
        if not self.fused and images_per_step != 1:
            raise ValueError('images_per_step > 1 needs the fused evaluation path')
        self.graph = None
        self._shape = None
        net.eval()

    # ------------------------------------------------------------------------------------------ reference-shaped path
    def _forward_probs(self, img, masks):
        if self.dedup:
            # the N copies are identical: the encoders run once, latent sampling and likelihood on all copies
            s_list = self.net.forward(img[None, None], masks[0][None, None], training=False, replicate=self.n_local)
        else:
            patch = img[None, None].repeat(self.n_local, 1, 1, 1)
            mask = masks[0][None, None].repeat(self.n_local, 1, 1, 1)
            s_list = self.net.forward(patch, mask, training=False)
        probs = self.net.accumulate_output(s_list, use_softmax=True)
        if self.counts is not None:                                # this rank's samples -> the full set on every rank
            from . import dp
            probs = dp.gather_samples(probs.contiguous(), self.counts)
        return probs

    # ------------------------------------------------------------------------------------------ fused path
    def _fused_body(self):
        """everything on the device; results into self.out [I, 2 + C] doubles (GED, NCC, Dice per class; rows of images
        owned by other ranks stay NaN unless gather_results); no host synchronisation"""
        I, C, n = self.images, self.n_classes, self.n_local
        H, W = self.img.shape[-2:]
        hw = H * W
        masks = self.lab.permute(0, 3, 1, 2).contiguous()                 # [I, M, H, W] uint8
        pairs = self.net.forward(self.img[:, None], masks[:, 0:1].float(), training=False, replicate=n, lowres_logits=True)
        levels, factors = [p[0] for p in pairs], [p[1] for p in pairs]
        labels = list(range(1, C))
        nmax = max(self.counts) if self.counts is not None else n
        nl = len(labels)
        words = (hw + 31) // 32
        bits, cnts, sums = kern.eval_sample_stats(levels, factors, n, I, C, (H, W), labels, out=(self.flat, self.sums))
        if self.world > 1:
            import torch.distributed as dist
            dist.all_gather_into_tensor(self.flat_all, self.flat)          # masks + counts: 2 KB per sample
            dist.all_reduce(self.sums, op=dist.ReduceOp.SUM)               # per-pixel sum p, sum log p
            nb = I * n * nl * words
            bl, cl = [], []
            for r in range(self.world):
                # every rank laid its buffer out for ITS sample count; the all-gather needs equal sizes, so the buffers
                # are sized for nmax samples and rank r's rows beyond counts[r] are padding
                nr = self.counts[r]
                row = self.flat_all[r]
                bl.append(row[:I * nr * nl * words].view(I, nr, nl, words))
                cl.append(row[I * nr * nl * words:I * nr * nl * (words + 1)].view(I, nr, nl))
            bits, cnts = torch.cat(bl, dim=1), torch.cat(cl, dim=1)
        self.out.fill_(float('nan'))
        for i in range(I):
            if i % self.world != self.rank:
                continue
            ged4 = kern.ged_from_bits(bits[i].contiguous(), cnts[i].contiguous(), masks[i], labels, hw)
            self.out[i, 0:1].copy_(ged4[0:1])
            kern.ncc_dice_from_sums(self.sums[i], masks[i], self.n, dice_annotator=0, out=self.out[i, 1:])
        if self.gather_results and self.world > 1:
            import torch.distributed as dist
            dist.all_gather_into_tensor(self.out_all, self.out)
            for i in range(I):
                self.out[i].copy_(self.out_all[i % self.world, i])

    def refresh_weights(self):
        """re-pack the bf16 tensor-core weight copies and re-fold the BatchNorm running statistics (static_weights mode)"""
        if self.static_weights:
            self.net._packer().refresh(force=True)
            self.net._folder().refresh(force=True)

    def _device_body(self):
        if not self.static_weights:
            return self._device_body_inner()
        pk, fd = self.net._packer(), self.net._folder()
        prev = (pk.external, fd.external)
        pk.external, fd.external = True, True          # refreshed by refresh_weights(), not by every forward
        try:
            return self._device_body_inner()
        finally:
            pk.external, fd.external = prev

    def _device_body_inner(self):
        if self.fused:
            return self._fused_body()
        # results into self.out (double [1, 2] = GED, NCC); no host synchronisation
        masks = self.lab[0].permute(2, 0, 1).float()               # [M,H,W]
        probs = self._forward_probs(self.img[0], masks)
        pred = kern.argmax_classes(probs)                          # torch.argmax(dim=1) of train_model.py:195
        ged4 = kern.ged(pred, masks, list(range(1, self.n_classes)))
        ks = torch.arange(self.n_classes, device=masks.device, dtype=masks.dtype).view(1, self.n_classes, 1, 1)
        onehot = (masks.unsqueeze(1) == ks).long()                 # utils.convert_batch_to_onehot
        ncc = kern.variance_ncc(probs, onehot)
        self.out.fill_(float('nan'))
        self.out[0, 0:1].copy_(ged4[0:1])
        self.out[0, 1:2].copy_(ncc)

    def _prepare(self, image, labels):
        dev = torch.device('cuda', torch.cuda.current_device())
        I, C = self.images, self.n_classes
        H, W = image.shape[-2:]
        self.img = torch.zeros((I, H, W), dtype=torch.float32, device=dev)
        self.lab = torch.zeros((I, H, W, labels.shape[-1]), dtype=labels.dtype, device=dev)
        self.out = torch.zeros((I, 2 + C), dtype=torch.float64, device=dev)
        self.out_all = torch.zeros((self.world, I, 2 + C), dtype=torch.float64, device=dev)
        self.out_host = torch.zeros((I, 2 + C), dtype=torch.float64).pin_memory()
        nmax = max(self.counts) if self.counts is not None else self.n_local
        words = (H * W + 31) // 32
        per_rank = I * nmax * (C - 1) * (words + 1)
        self.flat = torch.zeros((per_rank,), dtype=torch.int32, device=dev)
        self.flat_all = torch.zeros((self.world, per_rank), dtype=torch.int32, device=dev)
        self.sums = torch.zeros((I, 2, C, H * W), dtype=torch.float32, device=dev)
        self.img.copy_(image.reshape(I, H, W))
        self.lab.copy_(labels.reshape(self.lab.shape))
        self.refresh_weights()
        if self.use_graph:
            s = torch.cuda.Stream(device=dev)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s), torch.no_grad():
                for _ in range(2):
                    self._device_body()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            n0 = _lib.raw('uz_launch_count')()
            with torch.no_grad(), torch.cuda.graph(self.graph):
                self._device_body()
            self.launches_per_step = _lib.raw('uz_launch_count')() - n0
            torch.cuda.synchronize()
        self._shape = (tuple(image.shape), tuple(labels.shape), labels.dtype)

    @torch.no_grad()
    def run_host(self, image_pinned, labels_pinned):
        """image [H,W] or [I,H,W] fp32, labels [H,W,M] or [I,H,W,M] uint8 (host, pinned) -> for one image (ged, ncc);
        for I > 1 a float64 tensor [I, 2 + C] (GED, NCC, Dice per class; NaN rows = images owned by another rank)."""
        if self._shape != (tuple(image_pinned.shape), tuple(labels_pinned.shape), labels_pinned.dtype):
            self._prepare(image_pinned, labels_pinned)
        self.img.copy_(image_pinned.reshape(self.img.shape), non_blocking=True)
        self.lab.copy_(labels_pinned.reshape(self.lab.shape), non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._device_body()
        self.out_host.copy_(self.out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        if image_pinned.dim() == 2:
            return float(self.out_host[0, 0]), float(self.out_host[0, 1])
        return self.out_host.clone()

    @torch.no_grad()
    def run_device(self, img, lab, utils=None):
        """eager, through the drop-in ``utils`` functions (the second boundary): one image [H,W], labels [H,W,M]"""
        if utils is None:
            import utils
        masks = lab.permute(2, 0, 1).float()                       # [M,H,W]
        probs = self._forward_probs(img, masks)
        pred = kern.argmax_classes(probs)                          # torch.argmax(dim=1) of train_model.py:195
        ged = utils.generalised_energy_distance(pred, masks, nlabels=self.n_classes - 1,
                                                label_range=range(1, self.n_classes))
        onehot = utils.convert_batch_to_onehot(masks.unsqueeze(1), nlabels=self.n_classes)
        ncc = utils.variance_ncc_dist(probs, onehot)
        return ged, float(ncc[0])
