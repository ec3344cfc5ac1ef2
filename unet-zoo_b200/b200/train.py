"""Step drivers around the drop-in modules: the reference's training step (train_model.py:101-122) and its
N-sample evaluation (train_model.py:177-205) as reusable objects, optionally captured in a CUDA graph so the ~1.5 k
kernel launches of a PHiSeg step are replayed without Python in the loop.

The caller-visible semantics are the reference's: stock torch.optim.Adam(lr=1e-3, weight_decay=1e-5) on the fp32
parameters, loss = net.loss(mask) after net.forward(patch, mask, training=True).
"""
import os

import torch

from . import _lib, kern


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
    """forward(training=True) -> loss -> zero_grad -> backward -> [gradient all-reduce] -> Adam.step."""

    def __init__(self, net, optimizer, batch, image_size=(1, 128, 128), use_graph=True, dp=None, device=None):
        self.net, self.opt, self.dp = net, optimizer, dp
        self.device = device or torch.device('cuda', torch.cuda.current_device())
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
        loss.backward()
        if self.dp is not None:
            self.dp.finish()
        self.opt.step()
        self.loss.copy_(loss.detach())

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
        if self.use_graph:
            self.graph = torch.cuda.CUDAGraph()
            n0 = _lib.raw('uz_launch_count')()
            # the capture stream (and the side streams forked from it for the forward / dgrad chains) get a higher
            # priority than the auxiliary weight-gradient streams: critical-path kernels are scheduled first
            prio = int(os.environ.get('UNETZOO_MAIN_PRIORITY', '-2'))
            cap_stream = torch.cuda.Stream(device=self.device, priority=prio)
            with torch.cuda.graph(self.graph, stream=cap_stream):
                self._body()
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
    """One validation image like train_model.py:177-205: N copies -> forward(training=False) -> accumulate_output
    (softmax) -> argmax -> GED + NCC against M annotators.  ``shard`` = (rank, world) splits the N samples.
    ``run_host`` replays a CUDA graph of the whole evaluation (eager it is host-bound: ~350 launches from Python take
    twice as long as the kernels they start); ``run_device`` is the eager path through the drop-in ``utils`` functions."""

    def __init__(self, net, n_samples=100, n_classes=2, shard=None, dedup=True, use_graph=True):
        self.net = net
        self.dedup = dedup
        self.use_graph = use_graph
        self.n = n_samples
        self.n_classes = n_classes
        self.counts = None
        self.n_local = n_samples
        if shard is not None and shard[1] > 1:
            from . import dp
            self.counts = dp.shard_counts(n_samples, shard[1])
            self.n_local = self.counts[shard[0]]
        self.graph = None
        self._shape = None
        net.eval()

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

    def _device_body(self):
        """everything on the device, results into self.out (double [2] = GED, NCC); no host synchronisation"""
        masks = self.lab.permute(2, 0, 1).float()                  # [M,H,W]
        probs = self._forward_probs(self.img, masks)
        pred = kern.argmax_classes(probs)                          # torch.argmax(dim=1) of train_model.py:195
        ged4 = kern.ged(pred, masks, list(range(1, self.n_classes)))
        ks = torch.arange(self.n_classes, device=masks.device, dtype=masks.dtype).view(1, self.n_classes, 1, 1)
        onehot = (masks.unsqueeze(1) == ks).long()                 # utils.convert_batch_to_onehot
        ncc = kern.variance_ncc(probs, onehot)
        self.out[0:1].copy_(ged4[0:1])
        self.out[1:2].copy_(ncc)

    def _prepare(self, image, labels):
        dev = torch.device('cuda', torch.cuda.current_device())
        self.img = torch.zeros(tuple(image.shape), dtype=torch.float32, device=dev)
        self.lab = torch.zeros(tuple(labels.shape), dtype=labels.dtype, device=dev)
        self.out = torch.zeros(2, dtype=torch.float64, device=dev)
        self.out_host = torch.zeros(2, dtype=torch.float64).pin_memory()
        self.img.copy_(image)
        self.lab.copy_(labels)
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(2):
                self._device_body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self._device_body()
        torch.cuda.synchronize()
        self._shape = (tuple(image.shape), tuple(labels.shape), labels.dtype)

    @torch.no_grad()
    def run_host(self, image_pinned, labels_pinned):
        """image [H,W] fp32, labels [H,W,M] uint8 (host, pinned) -> (ged float, ncc float)"""
        if not self.use_graph:
            import utils  # the drop-in second boundary
            dev = torch.device('cuda', torch.cuda.current_device())
            return self.run_device(image_pinned.to(dev, non_blocking=True), labels_pinned.to(dev, non_blocking=True), utils)
        if self.graph is None or self._shape != (tuple(image_pinned.shape), tuple(labels_pinned.shape), labels_pinned.dtype):
            self._prepare(image_pinned, labels_pinned)
        self.img.copy_(image_pinned, non_blocking=True)
        self.lab.copy_(labels_pinned, non_blocking=True)
        self.graph.replay()
        self.out_host.copy_(self.out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self.out_host[0]), float(self.out_host[1])

    @torch.no_grad()
    def run_device(self, img, lab, utils=None):
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
