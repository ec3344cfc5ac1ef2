"""Thin tensor-level wrappers over the C ABI (no autograd).  Activations are torch bf16 tensors of shape
[N, H, W, C] (or [N, D, H, W, C] for the volumes of models/phiseg3D.py) whose last-dim stride is 1 and whose pixel
stride may exceed C (channel slices of concat buffers); N/(D/)H/W must be densely packed over pixels.  The flat
[pixels][channels] kernels see a volume as N*D images."""
import ctypes
import struct

import torch

from . import _lib

BN_EPS = 1e-3        # reference torchlayers.py:20
BN_MOMENTUM = 0.01


def _stream():
    """raw handle of the current CUDA stream.  torch.cuda.current_stream() costs ~14 us of Python per call (device
    index resolution, availability checks, a Stream object) -- with ~800 launches per eager step that was a third of
    the host time of the path the unmodified train_model.py takes; the two C calls below cost < 1 us."""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _p(t):
    return None if t is None else t.data_ptr()


def _check_act(t):
    """-> (images, H, W, C, pixel stride); a volume [N,D,H,W,C] counts as N*D images.  One shape / stride query each:
    this runs ~5 times per launch."""
    if not t.is_cuda:
        raise _lib.UnetZooLibError('B200 path needs CUDA tensors (no CPU fallback)')
    sh, st = t.shape, t.stride()
    if len(sh) == 5:
        n, d, h, w, c = sh
        ld = st[3]
        assert t.dtype == _lib.act_dtype() and st[4] == 1, (t.dtype, sh, st)
        assert (h == 1 or st[2] == w * ld) and (d == 1 or st[1] == h * w * ld) and (n == 1 or st[0] == d * h * w * ld), st
        return n * d, h, w, c, ld
    assert t.dtype == _lib.act_dtype() and len(sh) == 4 and st[3] == 1, (t.dtype, sh, st)
    n, h, w, c = sh
    ld = st[2]
    assert (h == 1 or st[1] == w * ld) and (n == 1 or st[0] == h * w * ld), st
    return n, h, w, c, ld


def pad16(c):
    return (c + 15) // 16 * 16


def pad_channels(c, weight_dim=4):
    """stored channel count of an activation / packed weight: the next multiple of 16 (images and volumes alike)"""
    return pad16(c)


def new_act(n, h, w, c, device):
    return torch.empty((n, h, w, c), dtype=_lib.act_dtype(), device=device)


def _like(x, c):
    """fresh dense activation with x's batch / spatial dims and c channels"""
    return torch.empty(tuple(x.shape[:-1]) + (c,), dtype=_lib.act_dtype(), device=x.device)


def _taps(w):
    t = 1
    for k in w.shape[2:]:
        t *= k
    return t


def _spatial_numel(shape):
    t = 1
    for k in shape:
        t *= k
    return t


def conv_tile_geometry(n, h, w):
    tw, th, tn, nt = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.call('uz_conv_tile_geometry', n, h, w, ctypes.byref(tw), ctypes.byref(th), ctypes.byref(tn), ctypes.byref(nt))
    return tw.value, th.value, tn.value, nt.value


class WeightPacker:
    """Packs the fp32 OIHW weights of ALL tensor-core conv layers of a model into their bf16 forward / dgrad layouts
    with ONE kernel launch per forward (uz_pack_conv_weights_batched) instead of one per layer.  It always re-packs:
    tensor version counters cannot be trusted to detect updates (torch's fused Adam writes parameters without bumping
    them), and one ~30 us launch per forward is cheaper than a stale weight."""

    def __init__(self, weights):
        self.weights = [w for w in weights]
        dev = self.weights[0].device
        self.packed = {}
        self.rows = {}
        self.max_elems = 0
        for w in self.weights:
            cout, cin = w.shape[0], w.shape[1]
            taps = _taps(w)
            coutp, cinp = pad_channels(cout, w.dim()), pad_channels(cin, w.dim())
            wf = torch.empty((taps, coutp, cinp), dtype=_lib.act_dtype(), device=dev)
            wd = torch.empty((taps, cinp, coutp), dtype=_lib.act_dtype(), device=dev)
            self.packed[w.data_ptr()] = (wf, wd)
            self.rows[w.data_ptr()] = (struct.pack('<QQQiiiiii', w.data_ptr(), wf.data_ptr(), wd.data_ptr(), cout, cin, taps,
                                                   coutp, cinp, 0), 2 * taps * coutp * cinp)
            self.max_elems = max(self.max_elems, 2 * taps * coutp * cinp)
        raw = b''.join(self.rows[w.data_ptr()][0] for w in self.weights)
        self.table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        self.ptr0 = self.weights[0].data_ptr()
        self.dtype = _lib.act_dtype()
        self.versions = None
        # external=True: somebody else (the overlapped optimizer of b200.dp: Adam + re-pack per gradient bucket during
        # backward) keeps the packed copies current; the per-forward refresh() is then a no-op unless forced
        self.external = False
        self.subsets = {}

    def valid_for(self, weights_first):
        return weights_first.data_ptr() == self.ptr0 and self.dtype == _lib.act_dtype()

    def refresh(self, force=False):
        if self.external and not force:
            return
        blocks = min(64, max(1, (self.max_elems + 255) // 256 // 4))
        _lib.call('uz_pack_conv_weights_batched', _p(self.table), len(self.weights), blocks, _stream())

    def make_subset(self, key, params):
        """descriptor table for the packed layers among ``params`` (allocate outside a stream capture)"""
        rows = [self.rows[p.data_ptr()] for p in params if p.data_ptr() in self.rows]
        if not rows:
            self.subsets[key] = None
            return
        raw = b''.join(r[0] for r in rows)
        self.subsets[key] = (torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.table.device), len(rows),
                             max(r[1] for r in rows))

    def refresh_subset(self, key):
        sub = self.subsets.get(key)
        if sub is None:
            return
        table, n, max_elems = sub
        blocks = min(64, max(1, (max_elems + 255) // 256 // 4))
        _lib.call('uz_pack_conv_weights_batched', _p(table), n, blocks, _stream())

    def lookup(self, w):
        return self.packed.get(w.data_ptr())


_active_packer = None


def set_active_packer(packer):
    global _active_packer
    prev = _active_packer
    _active_packer = packer
    return prev


def pack_conv_weight(w, need_dgrad=True):
    """w fp32 [Cout,Cin,kh,kw] -> (w_fwd bf16 [taps,CoutP,CinP], w_dgrad bf16 [taps,CinP,CoutP] or None).
    Inside a model forward the active WeightPacker already holds both (packed by one batched launch)."""
    if _active_packer is not None:
        hit = _active_packer.lookup(w)
        if hit is not None:
            return hit
    cout, cin = w.shape[0], w.shape[1]
    taps = _taps(w)
    coutp, cinp = pad_channels(cout, w.dim()), pad_channels(cin, w.dim())
    wf = torch.empty((taps, coutp, cinp), dtype=_lib.act_dtype(), device=w.device)
    wd = torch.empty((taps, cinp, coutp), dtype=_lib.act_dtype(), device=w.device) if need_dgrad else None
    _lib.call('uz_pack_conv_weight', _p(w.contiguous()), cout, cin, taps, _p(wf), coutp, cinp, _p(wd), cinp, coutp,
              _stream())
    return wf, wd


class UzConvExtra(ctypes.Structure):
    """include/unetzoo_b200.h: UzConvExtra"""
    _fields_ = [('stats_rows', ctypes.c_int), ('bn_y', ctypes.c_void_p), ('bn_ldy', ctypes.c_int),
                ('bn_scale', ctypes.c_void_p), ('bn_shift', ctypes.c_void_p), ('bn_relu', ctypes.c_int),
                ('bn_sums', ctypes.c_void_p), ('residual', ctypes.c_void_p), ('ld_res', ctypes.c_int),
                ('res_sign', ctypes.c_int)]


# Deterministic mode: BatchNorm statistics are written as one row per CTA and reduced in a fixed order (three launches
# per layer instead of two, no fp32 atomics anywhere in the training step) -> bit-reproducible training.  The default
# accumulates the statistics with fp32 atomics in CTA-arrival order.  UNETZOO_DETERMINISTIC=1 or set_deterministic(True).
import os as _os
_DETERMINISTIC = _os.environ.get('UNETZOO_DETERMINISTIC', '0') == '1'


def set_deterministic(enabled):
    global _DETERMINISTIC
    prev = _DETERMINISTIC
    _DETERMINISTIC = bool(enabled)
    return prev


def is_deterministic():
    return _DETERMINISTIC


def conv_fwd(x, w_packed, out=None, scale=None, shift=None, relu=False, stats=False, bn_prev=None, residual=None,
             res_sign=1):
    """x [N,H,W,Cin] bf16, w_packed [taps,Cout,Cin] bf16 -> y [N,H,W,Cout] bf16 (+ statistics).
    stats=True: per-channel sum / sum of squares of y, as [1,2,Cout] accumulators or -- deterministic mode -- as one row
    per CTA [rows,2,Cout].  bn_prev=(y_prev, scale_prev, shift_prev, relu_prev): this call is an input-gradient and its
    result is the gradient w.r.t. the activation a_prev = relu(y_prev*scale_prev+shift_prev); the epilogue applies that
    ReLU's mask and accumulates (sum g, sum g*y_prev) -> returned as the second value [2,Cout] (fused BatchNorm backward
    reduction).  residual: out = residual + res_sign * value."""
    n, h, w, cin, ldx = _check_act(x)
    taps, cout, cin_w = w_packed.shape
    assert cin_w == cin, (cin_w, cin)
    if out is None:
        out = _like(x, cout)
    _, _, _, _, ldy = _check_act(out)
    partial = None
    extra = None
    if stats:
        if _DETERMINISTIC and x.dim() == 4:
            rows = _lib.raw('uz_conv_stats_rows')(n, h, w, cin, cout, taps)
            partial = torch.empty((rows, 2, cout), dtype=torch.float32, device=x.device)
            extra = UzConvExtra(stats_rows=1)
        else:
            partial = zero_arena.get(2 * cout, x.device).view(1, 2, cout)      # [2][Cout] accumulators, zero on entry
    keep = None
    if bn_prev is not None or residual is not None:
        assert not stats
        extra = extra or UzConvExtra()
        if bn_prev is not None:
            y_prev, sc_prev, sh_prev, relu_prev = bn_prev
            assert y_prev.shape[:-1] == out.shape[:-1] and y_prev.shape[-1] >= cout
            partial = zero_arena.get(2 * cout, x.device).view(2, cout)
            extra.bn_y, extra.bn_ldy = y_prev.data_ptr(), _check_act(y_prev)[4]
            extra.bn_scale, extra.bn_shift, extra.bn_relu = sc_prev.data_ptr(), sh_prev.data_ptr(), int(relu_prev)
            extra.bn_sums = partial.data_ptr()
        if residual is not None:
            assert residual.shape == out.shape
            extra.residual, extra.ld_res, extra.res_sign = residual.data_ptr(), _check_act(residual)[4], int(res_sign)
    if extra is not None:
        if x.dim() == 5:
            _lib.call('uz_conv3d_fwd_ex', _p(x), x.shape[0], x.shape[1], h, w, cin, ldx, _p(w_packed), cout, taps, _p(out),
                      ldy, _p(scale), _p(shift), int(relu), None if bn_prev is not None else _p(partial),
                      ctypes.byref(extra), _stream())
        else:
            _lib.call('uz_conv_fwd_ex', _p(x), n, h, w, cin, ldx, _p(w_packed), cout, taps, _p(out), ldy, _p(scale),
                      _p(shift), int(relu), None if bn_prev is not None else _p(partial), ctypes.byref(extra), _stream())
        return out, partial
    if x.dim() == 5:
        _lib.call('uz_conv3d_fwd', _p(x), x.shape[0], x.shape[1], h, w, cin, ldx, _p(w_packed), cout, taps, _p(out), ldy,
                  _p(scale), _p(shift), int(relu), _p(partial), _stream())
    else:
        _lib.call('uz_conv_fwd', _p(x), n, h, w, cin, ldx, _p(w_packed), cout, taps, _p(out), ldy, _p(scale), _p(shift),
                  int(relu), _p(partial), _stream())
    return out, partial


class WgradReducer:
    """Deferred split-K reductions of the weight gradients (uz_conv_wgrad_partial + uz_wgrad_reduce_batched): the
    tensor-core kernels of many layers leave their partial slabs behind (one slab per layer when the splits accumulate in
    L2, one per split in deterministic mode and for volumes), ONE launch per <= 64 layers reduces / transposes them into
    the OIHW gradient tensors -- or nobody does: the fused optimizer of a single-GPU TrainStep reads the slabs itself
    (``hold`` / ``take``).  ``flush()`` must run on a stream that is ordered after every producing launch
    (b200.ops joins the auxiliary streams first).  The descriptor rows are launch parameters: nothing to upload, a
    captured CUDA graph holds them by value."""
    MAX_ROWS = 64

    def __init__(self):
        self.items = []
        self.pending_bytes = 0
        self.keep = []             # slabs / gradient tensors referenced by captured launches
        # hold: the end-of-backward flush is skipped -- the optimizer that follows reads the slabs itself
        # (FusedAdam with an attached packer: uz_adam_pack_step sums the splits while it loads a tile's gradients), takes
        # the items it can use with ``take`` and flushes the rest
        self.hold = False

    def take(self, dw_ptrs):
        """-> {dw data_ptr: (slab tensor, splits, coutp, cinp)} for the pending items whose gradient tensor is in
        ``dw_ptrs``; they leave the pending list (the caller consumes the slabs directly)"""
        want = set(dw_ptrs)
        out, rest = {}, []
        for it in self.items:
            ptr = it[7].data_ptr()
            if ptr in want and ptr not in out:
                out[ptr] = (it[0], it[1], it[3], it[4])
            else:
                rest.append(it)
        self.items = rest
        self.pending_bytes = sum(it[0].numel() * 4 for it in rest)
        if out:
            cur = torch.cuda.current_stream(next(iter(out.values()))[0].device)
            if torch.cuda.is_current_stream_capturing():
                self.keep.append(list(out.values()))
            else:
                for v in out.values():
                    v[0].record_stream(cur)
        return out

    def add(self, work, splits, taps, coutp, cinp, cout, cin, dw):
        self.items.append((work, splits, taps, coutp, cinp, cout, cin, dw))
        self.pending_bytes += work.numel() * 4

    def flush(self, force=True):
        if not self.items or (self.hold and not force):
            return
        items, self.items, self.pending_bytes = self.items, [], 0
        cur = torch.cuda.current_stream(items[0][0].device)
        capturing = torch.cuda.is_current_stream_capturing()
        for k in range(0, len(items), self.MAX_ROWS):
            part = items[k:k + self.MAX_ROWS]
            raw = b''.join(struct.pack('<QQiiiiiiii', work.data_ptr(), dw.data_ptr(), splits, taps, coutp, cinp, cout,
                                       cin, 0, 0) for work, splits, taps, coutp, cinp, cout, cin, dw in part)
            buf = ctypes.create_string_buffer(raw, len(raw))
            _lib.call('uz_wgrad_reduce_batched', ctypes.cast(buf, ctypes.c_void_p), len(part), _stream())
        if capturing:
            self.keep.append(items)
        else:
            for it in items:
                it[0].record_stream(cur)           # slabs may have been allocated on an auxiliary stream
                it[7].record_stream(cur)


wgrad_reducer = WgradReducer()
_WGRAD_ACCUMULATE = _os.environ.get('UNETZOO_WGRAD_ACCUMULATE', '1') != '0'


def conv_wgrad(x, dy, taps, cin_logical, cout_logical, out=None, defer=False):
    """-> dw fp32 [cout_logical, cin_logical, taps] (written into ``out``, any contiguous fp32 tensor of that size, when
    given: data-parallel training lets the kernel produce the gradient inside its all-reduce bucket).  ``defer``: only
    the tensor-core kernel runs now; dw is filled by the next ``wgrad_reducer.flush()`` (one launch for many layers)."""
    if out is not None:
        assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == cout_logical * cin_logical * taps
        out = out.view(cout_logical, cin_logical, taps)
    n, h, w, cin, ldx = _check_act(x)
    _, _, _, cout, lddy = _check_act(dy)
    vol = x.dim() == 5 and taps == 27
    if vol:
        nb, d = x.shape[0], x.shape[1]
        ws = _lib.raw('uz_wgrad3d_workspace_floats')(nb, d, h, w, cin, cout)
    else:
        nb, d = n, 0
        ws = _lib.raw('uz_wgrad_workspace_floats')(n, h, w, cin, cout, taps)
    if ws < 0:
        raise _lib.UnetZooLibError('uz_conv_wgrad: unsupported shape Cin=%d Cout=%d' % (cin, cout))
    dw = out if out is not None else torch.empty((cout_logical, cin_logical, taps), dtype=torch.float32, device=x.device)
    if defer:
        # split-K CTAs add into ONE slab through L2 (bulk reduce-add) unless the run has to be bit-reproducible; volumes
        # keep the slabs (few parameters, thousands of pixel tiles: 45.8 ms per PHISeg3D step with slabs, 46.9 ms without)
        acc = _WGRAD_ACCUMULATE and not _DETERMINISTIC and not vol
        work = torch.zeros((taps * cout * cin,), dtype=torch.float32, device=x.device) if acc else \
            torch.empty((ws,), dtype=torch.float32, device=x.device)
        splits = ctypes.c_int(0)
        _lib.call('uz_conv_wgrad_partial', _p(x), ldx, _p(dy), lddy, nb, d, h, w, cin, cout, taps, _p(work), int(acc),
                  ctypes.byref(splits), _stream())
        wgrad_reducer.add(work, splits.value, taps, cout, cin, cout_logical, cin_logical, dw)
        return dw
    work = torch.empty((ws,), dtype=torch.float32, device=x.device)
    if vol:
        _lib.call('uz_conv3d_wgrad', _p(x), ldx, _p(dy), lddy, nb, d, h, w, cin, cout, cin_logical, cout_logical,
                  _p(work), _p(dw), _stream())
        return dw
    _lib.call('uz_conv_wgrad', _p(x), ldx, _p(dy), lddy, n, h, w, cin, cout, taps, cin_logical, cout_logical, _p(work),
              _p(dw), _stream())
    return dw


_CONV_BN_FUSED = _os.environ.get('UNETZOO_CONV_BN_FUSED', '1') != '0'


def conv_bn_fused_supported(x, w_packed):
    if not _CONV_BN_FUSED or x.dim() != 4:
        return False
    n, h, w, cin, _ = _check_act(x)
    taps, cout, _ = w_packed.shape
    return bool(_lib.raw('uz_conv_bn_fused_supported')(n, h, w, cin, cout, taps))


def conv_bn_act_fused(x, w_packed, bias, gamma, beta, running_mean, running_var, relu=True, eps=None, momentum=None,
                      stat_updates=1):
    """Conv2D (training) in one launch on a thread-block cluster -> (a, y, scale, shift, mean, invstd); see
    uz_conv_bn_act_fused.  Only for layers with conv_bn_fused_supported()."""
    n, h, w, cin, ldx = _check_act(x)
    taps, cout, cin_w = w_packed.shape
    assert cin_w == cin, (cin_w, cin)
    y = new_act(n, h, w, cout, x.device)
    a = new_act(n, h, w, cout, x.device)
    st = torch.empty((4, cout), dtype=torch.float32, device=x.device)
    _lib.call('uz_conv_bn_act_fused', _p(x), n, h, w, cin, ldx, _p(w_packed), cout, taps, _p(bias), _p(gamma), _p(beta),
              BN_EPS if eps is None else eps, BN_MOMENTUM if momentum is None else momentum, _p(running_mean),
              _p(running_var), int(stat_updates), int(relu), _p(y), cout, _p(a), cout, _p(st[0]), _p(st[1]), _p(st[2]),
              _p(st[3]), _stream())
    return a, y, st[0], st[1], st[2], st[3]


def bn_finalize(partial, count, gamma, beta, running_mean=None, running_var=None, eps=BN_EPS, momentum=BN_MOMENTUM):
    tiles, _, c = partial.shape
    dev = partial.device
    scale = torch.empty(c, dtype=torch.float32, device=dev)
    shift = torch.empty_like(scale)
    mean = torch.empty_like(scale)
    invstd = torch.empty_like(scale)
    _lib.call('uz_bn_finalize', _p(partial), tiles, c, float(count), _p(gamma), _p(beta), eps, momentum,
              _p(running_mean), _p(running_var), _p(scale), _p(shift), _p(mean), _p(invstd), _stream())
    return scale, shift, mean, invstd


def bn_apply_train(y, sums, count, gamma, beta, running_mean, running_var, relu=True, eps=BN_EPS, momentum=BN_MOMENTUM,
                   residual=None, res_sign=1, out=None, stat_updates=1):
    """finalize + normalise + ReLU in one launch -> (a, scale, shift, mean, invstd).  ``residual``: a = residual +
    res_sign * act(...) (reversible coupling, written into ``out`` -- e.g. a channel slice of the block output);
    ``stat_updates``: momentum updates of the running statistics (2 for reversible blocks, quirk Q7)."""
    n, h, w, c, ldy = _check_act(y)
    dev = y.device
    st = torch.empty((4, c), dtype=torch.float32, device=dev)
    if out is None:
        out = _like(y, c)
    ldo = _check_act(out)[4]
    if residual is None and stat_updates == 1:
        _lib.call('uz_bn_apply_train', _p(y), ldy, _p(sums), float(count), _p(gamma), _p(beta), eps, momentum,
                  _p(running_mean), _p(running_var), _p(st[0]), _p(st[1]), _p(st[2]), _p(st[3]), int(relu), _p(out), ldo,
                  n * h * w, c, _stream())
    else:
        ldr = _check_act(residual)[4] if residual is not None else 0
        _lib.call('uz_bn_apply_train_ex', _p(y), ldy, _p(sums), float(count), _p(gamma), _p(beta), eps, momentum,
                  _p(running_mean), _p(running_var), _p(st[0]), _p(st[1]), _p(st[2]), _p(st[3]), int(relu), _p(out), ldo,
                  n * h * w, c, _p(residual), ldr, int(res_sign), int(stat_updates), _stream())
    return out, st[0], st[1], st[2], st[3]


_BN_BWD_FUSED = _os.environ.get('UNETZOO_BN_BWD_FUSED', '1') != '0'
# in the captured step (gpurun_out/r2c_tl_*.json) the cluster kernel wins up to 16x16x12 pixels (4.9 vs 6.1 us at 2x2,
# 7.3 vs 8.9 us at 16x16) and loses from 32x32x12 on (13-15 vs 11.6 us)
_BN_BWD_FUSED_MAX_PIX = int(_os.environ.get('UNETZOO_BN_BWD_FUSED_MAX_PIX', '3072'))
# one cooperative launch (sums, grid-wide barrier, apply) for large maps: correct, but a cooperative grid waits until the
# whole GPU can take it -- in the multi-stream step that serialises the backward streams: 4.33 vs 4.08 ms -> opt-in
_BN_BWD_COOP = _os.environ.get('UNETZOO_BN_BWD_COOP', '0') == '1'
_BN_BWD_COOP_MIN_PIX = int(_os.environ.get('UNETZOO_BN_BWD_COOP_MIN_PIX', '49152'))


def bn_relu_bwd_train(dout, y, scale, shift, gamma, mean, invstd, relu=True, sums=None, inverse=None):
    """two launches (accumulate, apply) -> dy bf16, dgamma, dbeta; with ``sums`` ([2,C]: sum g, sum g*y, already
    accumulated by the dgrad epilogue that produced ``dout``) only the apply pass runs.  ``inverse`` = (inv_in, inv_out):
    the accumulate pass also writes inv_out = inv_in - act(y*scale+shift) (inverse of a reversible coupling)."""
    n, h, w, c, ldd = _check_act(dout)
    ldy = _check_act(y)[4]
    npix = n * h * w
    dev = y.device
    if sums is None and inverse is None and _BN_BWD_FUSED and npix <= _BN_BWD_FUSED_MAX_PIX and \
            _lib.raw('uz_bn_bwd_fused_supported')(npix, c):
        # one launch on thread-block clusters (csrc/bn_cluster.cu): sums exchanged through distributed shared memory
        dgb = torch.empty((2, c), dtype=torch.float32, device=dev)
        dy = _like(y, c)
        _lib.call('uz_bn_bwd_fused', _p(dout), ldd, _p(y), ldy, _p(scale), _p(shift), int(relu), float(npix), _p(gamma),
                  _p(mean), _p(invstd), _p(dgb[0]), _p(dgb[1]), _p(dy), c, npix, c, _stream())
        return dy, dgb[0], dgb[1]
    if sums is None and inverse is None and _BN_BWD_COOP and npix >= _BN_BWD_COOP_MIN_PIX and c <= 1024:
        # large maps: sums + apply in one cooperative launch (grid-wide barrier in between)
        sums = zero_arena.get(2 * c, dev)
        dgb = torch.empty((2, c), dtype=torch.float32, device=dev)
        dy = _like(y, c)
        _lib.call('uz_bn_bwd_coop', _p(dout), ldd, _p(y), ldy, _p(scale), _p(shift), int(relu), _p(sums), float(npix),
                  _p(gamma), _p(mean), _p(invstd), _p(dgb[0]), _p(dgb[1]), _p(dy), c, npix, c, _stream())
        return dy, dgb[0], dgb[1]
    if sums is None:
        sums = zero_arena.get(2 * c, dev)
        if inverse is None:
            _lib.call('uz_bn_bwd_reduce_sums', _p(dout), ldd, _p(y), ldy, _p(scale), _p(shift), int(relu), npix, c,
                      _p(sums), _stream())
        else:
            inv_in, inv_out = inverse
            _lib.call('uz_bn_bwd_reduce_sums_ex', _p(dout), ldd, _p(y), ldy, _p(scale), _p(shift), int(relu), npix, c,
                      _p(sums), _p(inv_in), _check_act(inv_in)[4], _p(inv_out), _check_act(inv_out)[4], _stream())
    else:
        assert inverse is None
    dgb = torch.empty((2, c), dtype=torch.float32, device=dev)
    dy = _like(y, c)
    _lib.call('uz_bn_bwd_apply_train', _p(dout), ldd, _p(y), ldy, _p(scale), _p(shift), int(relu), _p(sums), float(npix),
              _p(gamma), _p(mean), _p(invstd), _p(dgb[0]), _p(dgb[1]), _p(dy), c, npix, c, _stream())
    return dy, dgb[0], dgb[1]


class BNFolder:
    """Eval-mode BatchNorm folds (scale, shift per layer) of ALL conv units of a model with ONE launch per forward
    (uz_bn_eval_fold_batched) instead of one tiny launch per layer (106 for PHiSeg: 160 us of a 5.7 ms evaluation)."""

    def __init__(self, units):
        """units: list of (conv_bias, gamma, beta, running_mean, running_var) tensors (fp32, same device)"""
        dev = units[0][3].device
        total = sum(u[3].numel() for u in units)
        self.buf = torch.empty((2, total), dtype=torch.float32, device=dev)
        self.views = {}
        rows, off, self.maxc = [], 0, 0
        for bias, gamma, beta, rm, rv in units:
            c = rm.numel()
            sc, sh = self.buf[0, off:off + c], self.buf[1, off:off + c]
            self.views[rm.data_ptr()] = (sc, sh)
            rows.append(struct.pack('<QQQQQQQq', _p(bias) or 0, _p(gamma) or 0, _p(beta) or 0, rm.data_ptr(), rv.data_ptr(),
                                    sc.data_ptr(), sh.data_ptr(), c))
            off += c
            self.maxc = max(self.maxc, c)
        self.n = len(units)
        self.table = torch.frombuffer(bytearray(b''.join(rows)), dtype=torch.uint8).to(dev)
        self.ptr0 = units[0][3].data_ptr()
        self.external = False       # True: somebody (EvalStep with static weights) refreshes explicitly; per-forward no-op

    def valid_for(self, rm_first):
        return rm_first.data_ptr() == self.ptr0

    def refresh(self, eps=BN_EPS, force=False):
        if self.external and not force:
            return
        _lib.call('uz_bn_eval_fold_batched', _p(self.table), self.n, self.maxc, eps, _stream())


_active_folder = None


def set_active_folder(folder):
    global _active_folder
    prev = _active_folder
    _active_folder = folder
    return prev


def bn_eval_fold(conv_bias, gamma, beta, rm, rv, eps=BN_EPS):
    if _active_folder is not None:
        hit = _active_folder.views.get(rm.data_ptr())
        if hit is not None:
            return hit
    c = rm.numel()
    scale = torch.empty(c, dtype=torch.float32, device=rm.device)
    shift = torch.empty_like(scale)
    _lib.call('uz_bn_eval_fold', _p(conv_bias), _p(gamma), _p(beta), _p(rm), _p(rv), eps, c, _p(scale), _p(shift),
              _stream())
    return scale, shift


def affine_act(y, scale, shift, relu=True, out=None):
    n, h, w, c, ldy = _check_act(y)
    if out is None:
        out = _like(y, c)
    ldo = _check_act(out)[4]
    _lib.call('uz_affine_act', _p(y), ldy, _p(scale), _p(shift), int(relu), _p(out), ldo, n * h * w, c, _stream())
    return out


def bn_relu_bwd(dout, y, scale, shift, gamma, mean, invstd, relu=True):
    """-> dy bf16, dgamma, dbeta (fp32 [C])"""
    n, h, w, c, ldd = _check_act(dout)
    ldy = _check_act(y)[4]
    npix = n * h * w
    nb = _lib.raw('uz_bn_bwd_num_blocks')(npix, c)
    dev = y.device
    partial = torch.empty((nb, 2, c), dtype=torch.float32, device=dev)
    _lib.call('uz_bn_bwd_reduce', _p(dout), ldd, _p(y), ldy, _p(scale), _p(shift), int(relu), npix, c, _p(partial),
              _stream())
    coef = torch.empty((3, c), dtype=torch.float32, device=dev)
    dgamma = torch.empty(c, dtype=torch.float32, device=dev)
    dbeta = torch.empty(c, dtype=torch.float32, device=dev)
    _lib.call('uz_bn_bwd_finalize', _p(partial), nb, c, float(npix), _p(gamma), _p(mean), _p(invstd), _p(coef[0]),
              _p(coef[1]), _p(coef[2]), _p(dgamma), _p(dbeta), _stream())
    dy = _like(y, c)
    _lib.call('uz_bn_bwd_apply', _p(dout), ldd, _p(y), ldy, _p(scale), _p(shift), int(relu), _p(coef[0]), _p(coef[1]),
              _p(coef[2]), _p(dy), c, npix, c, _stream())
    return dy, dgamma, dbeta


def relu_bwd(dout, y, scale, shift):
    """plain (bias+)ReLU backward: dy = dout * [y*scale+shift > 0]"""
    n, h, w, c, ldd = _check_act(dout)
    ldy = _check_act(y)[4]
    dy = _like(y, c)
    _lib.call('uz_bn_bwd_apply', _p(dout), ldd, _p(y), ldy, _p(scale), _p(shift), 1, None, None, None, _p(dy), c,
              n * h * w, c, _stream())
    return dy


def avgpool2_fwd(x):
    n, h, w, c, ldx = _check_act(x)
    if x.dim() == 5:
        nb, d = x.shape[0], x.shape[1]
        out = torch.empty((nb, d // 2, h // 2, w // 2, c), dtype=_lib.act_dtype(), device=x.device)
        _lib.call('uz_avgpool3_fwd', _p(x), ldx, _p(out), c, nb, d // 2, h // 2, w // 2, c, _stream())
        return out
    out = new_act(n, h // 2, w // 2, c, x.device)
    _lib.call('uz_avgpool2_fwd', _p(x), ldx, _p(out), c, n, h // 2, w // 2, c, _stream())
    return out


def avgpool2_bwd(dout):
    n, ho, wo, c, ldd = _check_act(dout)
    if dout.dim() == 5:
        nb, do = dout.shape[0], dout.shape[1]
        dx = torch.empty((nb, do * 2, ho * 2, wo * 2, c), dtype=_lib.act_dtype(), device=dout.device)
        _lib.call('uz_avgpool3_bwd', _p(dout), ldd, _p(dx), c, nb, do, ho, wo, c, _stream())
        return dx
    dx = new_act(n, ho * 2, wo * 2, c, dout.device)
    _lib.call('uz_avgpool2_bwd', _p(dout), ldd, _p(dx), c, n, ho, wo, c, 0, _stream())
    return dx


def upsample2x_fwd(x, align_corners=True, out=None):
    n, h, w, c, ldx = _check_act(x)
    if x.dim() == 5:                       # trilinear, align_corners=True only (models/phiseg3D.py)
        assert align_corners
        nb, d = x.shape[0], x.shape[1]
        if out is None:
            out = torch.empty((nb, 2 * d, 2 * h, 2 * w, c), dtype=_lib.act_dtype(), device=x.device)
        ldo = _check_act(out)[4]
        _lib.call('uz_upsample3d_fwd', _p(x), ldx, _p(out), ldo, nb, d, h, w, c, _stream())
        return out
    if out is None:
        out = new_act(n, 2 * h, 2 * w, c, x.device)
    ldo = _check_act(out)[4]
    _lib.call('uz_upsample2x_fwd', _p(x), ldx, _p(out), ldo, n, h, w, c, int(align_corners), _stream())
    return out


def upsample2x_bwd(dout, align_corners=True):
    n, hh, ww, c, ldd = _check_act(dout)
    if dout.dim() == 5:
        nb, dd = dout.shape[0], dout.shape[1]
        dx = torch.empty((nb, dd // 2, hh // 2, ww // 2, c), dtype=_lib.act_dtype(), device=dout.device)
        _lib.call('uz_upsample3d_bwd', _p(dout), ldd, _p(dx), c, nb, dd // 2, hh // 2, ww // 2, c, _stream())
        return dx
    dx = new_act(n, hh // 2, ww // 2, c, dout.device)
    _lib.call('uz_upsample2x_bwd', _p(dout), ldd, _p(dx), c, n, hh // 2, ww // 2, c, int(align_corners), _stream())
    return dx


def copy_channels(src, dst, accumulate=False):
    """accumulate: False/0 copy, True/1 dst += src, 2 dst -= src"""
    n, h, w, c, lds = _check_act(src)
    nd, _, _, _, ldd = _check_act(dst)
    if dst.shape[0] > src.shape[0] and dst.shape[0] % src.shape[0] == 0 and not accumulate and \
            tuple(src.shape[1:]) == tuple(dst.shape[1:]):
        # I images replicated over a batch of copies, dst image index = copy * I + image (evaluation with shared encoders)
        _lib.call('uz_copy_channels_bcast', _p(src), lds, n * h * w, _p(dst), ldd, nd * h * w, c, _stream())
        return dst
    assert dst.shape == src.shape
    _lib.call('uz_copy_channels', _p(src), lds, _p(dst), ldd, n * h * w, c, int(accumulate), _stream())
    return dst


def add_channels(a, b, out, sign=1):
    """out = a + b (sign >= 0) or a - b (sign < 0); all three may be channel slices"""
    n, h, w, c, lda = _check_act(a)
    ldb = _check_act(b)[4]
    ldo = _check_act(out)[4]
    assert a.shape == b.shape == out.shape
    _lib.call('uz_add_channels', _p(a), lda, _p(b), ldb, _p(out), ldo, n * h * w, c, int(sign), _stream())
    return out


def global_mean_fwd(x):
    n, h, w, c, ldx = _check_act(x)
    out = new_act(n, 1, 1, c, x.device)
    _lib.call('uz_global_mean_fwd', _p(x), ldx, n, h * w, c, _p(out), c, _stream())
    return out


def global_mean_bwd(dout, h, w):
    n, _, _, c, ldd = _check_act(dout)
    dx = new_act(n, h, w, c, dout.device)
    _lib.call('uz_global_mean_bwd', _p(dout), ldd, n, h * w, c, _p(dx), c, _stream())
    return dx


def input_pack(patch, mask, nlabels=2, cp=16):
    """patch fp32 [B,Cimg,(D,)H,W], mask index map [B,1,(D,)H,W] or None -> bf16 channel-last [B,(D,)H,W,cp]"""
    b, cimg = patch.shape[0], patch.shape[1]
    sp = tuple(patch.shape[2:])
    assert cimg + (nlabels if mask is not None else 0) <= cp
    out = torch.empty((b,) + sp + (cp,), dtype=_lib.act_dtype(), device=patch.device)
    patch = patch.contiguous().float()
    if mask is not None:
        mask = mask.contiguous().float()
    _lib.call('uz_input_pack', _p(patch), _p(mask), b, cimg, _spatial_numel(sp[:-1]), sp[-1], nlabels, _p(out), cp,
              _stream())
    return out


def nchw_to_nhwc(x, ld=None):
    """fp32 [B,C,*spatial] -> bf16 [B,*spatial,ld]"""
    b, c = x.shape[0], x.shape[1]
    sp = tuple(x.shape[2:])
    ld = ld or pad16(c)
    out = torch.empty((b,) + sp + (ld,), dtype=_lib.act_dtype(), device=x.device)
    _lib.call('uz_nchw_to_nhwc', _p(x.contiguous().float()), b, c, _spatial_numel(sp), _p(out), ld, _stream())
    return out


def nhwc_to_nchw(x, c=None):
    _, _, _, cc, ld = _check_act(x)
    c = c or cc
    b = x.shape[0]
    sp = tuple(x.shape[1:-1])
    out = torch.empty((b, c) + sp, dtype=torch.float32, device=x.device)
    _lib.call('uz_nhwc_to_nchw', _p(x), ld, b, c, _spatial_numel(sp), _p(out), _stream())
    return out


def head_fwd(feat, wmu, bmu, wsig, bsig, eps):
    _, _, _, c, ld = _check_act(feat)
    n = feat.shape[0]
    sp = tuple(feat.shape[1:-1])
    hw = _spatial_numel(sp)
    zdim = wmu.shape[0]
    mu = torch.empty((n, zdim) + sp, dtype=torch.float32, device=feat.device)
    sigma = torch.empty_like(mu)
    z = torch.empty_like(mu)
    _lib.call('uz_head_fwd', _p(feat), ld, c, _p(wmu), _p(bmu), _p(wsig), _p(bsig), _p(eps), n, hw, zdim, _p(mu),
              _p(sigma), _p(z), _stream())
    return mu, sigma, z


def head_bwd(feat, wmu, wsig, eps, sigma, dmu, dsigma, dz):
    _, _, _, c, ld = _check_act(feat)
    n = feat.shape[0]
    hw = _spatial_numel(feat.shape[1:-1])
    zdim = wmu.shape[0]
    dev = feat.device
    nb = _lib.raw('uz_head_bwd_num_blocks')(n, hw)
    wpartial = torch.empty((nb, 2 * zdim, c), dtype=torch.float32, device=dev)
    bpartial = torch.empty((nb, 2 * zdim), dtype=torch.float32, device=dev)
    dw = torch.empty((2 * zdim, c), dtype=torch.float32, device=dev)
    db = torch.empty((2 * zdim,), dtype=torch.float32, device=dev)
    dfeat = _like(feat, c)
    _lib.call('uz_head_bwd', _p(feat), ld, c, _p(wmu), _p(wsig), _p(eps), _p(sigma), _p(dmu), _p(dsigma), _p(dz), n,
              hw, zdim, _p(dfeat), c, _p(wpartial), _p(bpartial), _p(dw), _p(db), _stream())
    return dfeat, dw, db


def kl_fwd(mu0, s0, mu1, s1, weight):
    b = mu0.shape[0]
    per = mu0.numel() // b
    out = torch.empty((1,), dtype=torch.float32, device=mu0.device)
    partial = torch.empty((_lib.raw('uz_kl_num_blocks')(b, per),), dtype=torch.float64, device=mu0.device)
    _lib.call('uz_kl_fwd', _p(mu0), _p(s0), _p(mu1), _p(s1), b, per, float(weight), _p(out), _p(partial), _stream())
    return out


def kl_bwd(mu0, s0, mu1, s1, weight, upstream):
    b = mu0.shape[0]
    per = mu0.numel() // b
    g = [torch.empty_like(mu0) for _ in range(4)]
    _lib.call('uz_kl_bwd', _p(mu0), _p(s0), _p(mu1), _p(s1), b, per, float(weight), _p(upstream), _p(g[0]), _p(g[1]),
              _p(g[2]), _p(g[3]), _stream())
    return g


def _kl_args(levels, level_weights):
    L = len(levels)
    cols = [[t[k] for t in levels] for k in range(4)]
    numel = (ctypes.c_longlong * L)(*[t[0].numel() for t in levels])
    wts = (ctypes.c_float * L)(*[float(w) for w in level_weights])
    return L, [_ptr_array(c) for c in cols], numel, wts


def kl_hierarchy_fwd(levels, level_weights, total_weight):
    """levels: L tuples (mu0, sigma0, mu1, sigma1) of contiguous fp32 tensors -> (total fp32 [1], per-level fp32 [L])"""
    L, ptrs, numel, wts = _kl_args(levels, level_weights)
    dev = levels[0][0].device
    nb = _lib.raw('uz_kl_hierarchy_num_blocks')(max(t[0].numel() for t in levels))
    partial = torch.empty((L * nb,), dtype=torch.float64, device=dev)
    out = torch.empty((L + 1,), dtype=torch.float32, device=dev)
    _lib.call('uz_kl_hierarchy_fwd', ptrs[0], ptrs[1], ptrs[2], ptrs[3], numel, wts, L, levels[0][0].shape[0],
              float(total_weight), _p(partial), _p(out[:L]), _p(out[L:]), _stream())
    return out[L:], out[:L]


def kl_hierarchy_bwd(levels, level_weights, total_weight, upstream):
    L, ptrs, numel, wts = _kl_args(levels, level_weights)
    grads = [torch.empty_like(t) for lv in levels for t in lv]
    _lib.call('uz_kl_hierarchy_bwd', ptrs[0], ptrs[1], ptrs[2], ptrs[3], numel, wts, L, levels[0][0].shape[0],
              float(total_weight), _p(upstream), _ptr_array(grads), _stream())
    return grads


def slayer_fwd(feat, w, bias, factor):
    n, h, wd, c, ld = _check_act(feat)
    ncls = w.shape[0]
    if feat.dim() == 5:
        nb, d = feat.shape[0], feat.shape[1]
        out = torch.empty((nb, ncls, d * factor, h * factor, wd * factor), dtype=torch.float32, device=feat.device)
        _lib.call('uz_slayer3d_fwd', _p(feat), ld, c, _p(w), _p(bias), ncls, nb, d, h, wd, factor, _p(out), _stream())
        return out
    out = torch.empty((n, ncls, h * factor, wd * factor), dtype=torch.float32, device=feat.device)
    _lib.call('uz_slayer_fwd', _p(feat), ld, c, _p(w), _p(bias), ncls, n, h, wd, factor, _p(out), _stream())
    return out


def slayer_bwd(dout, feat, w, factor):
    n, h, wd, c, ld = _check_act(feat)
    ncls = w.shape[0]
    dev = feat.device
    nb = _lib.raw('uz_slayer_bwd_num_blocks')(n, h, wd)
    wpartial = torch.empty((nb, ncls, c), dtype=torch.float32, device=dev)
    bpartial = torch.empty((nb, ncls), dtype=torch.float32, device=dev)
    dw = torch.empty((ncls, c), dtype=torch.float32, device=dev)
    db = torch.empty((ncls,), dtype=torch.float32, device=dev)
    dfeat = _like(feat, c)
    if feat.dim() == 5:
        _lib.call('uz_slayer3d_bwd', _p(dout.contiguous()), _p(feat), ld, c, _p(w), ncls, feat.shape[0], feat.shape[1],
                  h, wd, factor, _p(dfeat), c, _p(wpartial), _p(bpartial), _p(dw), _p(db), _stream())
        return dfeat, dw, db
    _lib.call('uz_slayer_bwd', _p(dout.contiguous()), _p(feat), ld, c, _p(w), ncls, n, h, wd, factor, _p(dfeat), c,
              _p(wpartial), _p(bpartial), _p(dw), _p(db), _stream())
    return dfeat, dw, db


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


_ones = {}


def one_scalar(device):
    """a cached fp32 device scalar 1.0 (created outside any stream capture on first use)"""
    key = (device.type, device.index)
    if key not in _ones:
        _ones[key] = torch.ones((1,), dtype=torch.float32, device=device)
    return _ones[key]


def residual_ce(s_list, target, need_grad=True, upstream=None):
    """s_list: L fp32 NCHW logits (index = latent level); -> ce_levels fp32 [L], grads list or None."""
    L = len(s_list)
    b, ncls = s_list[0].shape[0], s_list[0].shape[1]
    hw = _spatial_numel(s_list[0].shape[2:])
    dev = s_list[0].device
    s_list = [s.contiguous() for s in s_list]
    grads = [torch.empty_like(s) for s in s_list] if need_grad else None
    nb = _lib.raw('uz_residual_ce_num_blocks')(b, hw)
    partial = torch.empty((nb, L), dtype=torch.float32, device=dev)
    ce = torch.empty((L,), dtype=torch.float32, device=dev)
    target = target.contiguous().float()
    _lib.call('uz_residual_ce', _ptr_array(s_list), _ptr_array(grads) if need_grad else None, _p(upstream), L, ncls,
              _p(target), b, hw, _p(partial), _p(ce), _stream())
    return ce, grads


def accumulate_output(s_list, use_softmax, out):
    L = len(s_list)
    b, ncls = s_list[0].shape[0], s_list[0].shape[1]
    hw = _spatial_numel(s_list[0].shape[2:])
    _lib.call('uz_accumulate_output', _ptr_array(s_list), L, ncls, b, hw, int(use_softmax), _p(out), _stream())
    return out


_DTYPE_CODE = {torch.int64: 0, torch.float32: 1, torch.uint8: 2}


def ged(samples, gts, label_values):
    """samples [N,H,W], gts [M,H,W] (int64 / float32 / uint8) -> double tensor [4] = GED, sum d_sy, d_ss, d_yy."""
    if not samples.is_cuda:
        raise _lib.UnetZooLibError('B200 path needs CUDA tensors (no CPU fallback)')
    samples = samples.contiguous()
    gts = gts.contiguous()
    n, m = samples.shape[0], gts.shape[0]
    hw = samples[0].numel()
    assert gts[0].numel() == hw
    nl = len(label_values)
    words = (hw + 31) // 32
    dev = samples.device
    lv = (ctypes.c_int * nl)(*[int(v) for v in label_values])
    bits_s = torch.empty((n, nl, words), dtype=torch.int32, device=dev)
    cnt_s = torch.empty((n, nl), dtype=torch.int32, device=dev)
    bits_y = torch.empty((m, nl, words), dtype=torch.int32, device=dev)
    cnt_y = torch.empty((m, nl), dtype=torch.int32, device=dev)
    _lib.call('uz_ged_pack_masks', _p(samples), _DTYPE_CODE[samples.dtype], n, hw, lv, nl, _p(bits_s), _p(cnt_s),
              _stream())
    _lib.call('uz_ged_pack_masks', _p(gts), _DTYPE_CODE[gts.dtype], m, hw, lv, nl, _p(bits_y), _p(cnt_y), _stream())
    pair_d = torch.empty((n * m + n * n + m * m,), dtype=torch.float64, device=dev)
    out = torch.empty((4,), dtype=torch.float64, device=dev)
    _lib.call('uz_ged_pairwise', _p(bits_s), _p(cnt_s), n, _p(bits_y), _p(cnt_y), m, nl, hw, _p(pair_d), _p(out),
              _stream())
    return out


def eval_sample_stats(levels, factors, n, images, n_classes, hw_shape, label_values, out=None):
    """levels: L fp32 tensors [n*images, C, h_l, w_l] (batch index = sample * images + image), factors: H / h_l.
    -> bits int32 [images, n, nl, words], counts int32 [images, n, nl], sums fp32 [images, 2, C, H*W]
    (``out`` = (flat int32 buffer holding bits then counts, sums) lets the caller own the buffers, e.g. for collectives)"""
    H, W = hw_shape
    hw = H * W
    nl = len(label_values)
    words = (hw + 31) // 32
    dev = levels[0].device
    levels = [t.contiguous() for t in levels]
    for t, f in zip(levels, factors):
        assert t.dtype == torch.float32 and t.shape[0] == n * images and t.shape[1] == n_classes and \
            t.shape[2] * f == H and t.shape[3] * f == W, (tuple(t.shape), f)
    if out is None:
        flat = torch.empty((images * n * nl * (words + 1),), dtype=torch.int32, device=dev)
        sums = torch.empty((images, 2, n_classes, hw), dtype=torch.float32, device=dev)
    else:
        flat, sums = out
    nb = images * n * nl * words
    bits = flat[:nb].view(images, n, nl, words)
    counts = flat[nb:nb + images * n * nl].view(images, n, nl)
    G = _lib.raw('uz_eval_sample_groups')(n, images, hw)
    part = torch.empty((images, G, 2, n_classes, hw), dtype=torch.float32, device=dev)
    lv = (ctypes.c_int * nl)(*[int(v) for v in label_values])
    fa = (ctypes.c_int * len(factors))(*[int(f) for f in factors])
    _lib.call('uz_eval_sample_stats', _ptr_array(levels), fa, len(levels), n, images, n_classes, H, W, lv, nl, _p(bits),
              _p(counts), _p(part), _p(sums), _stream())
    return bits, counts, sums


def ged_from_bits(bits_s, cnt_s, gts, label_values, hw):
    """bits_s int32 [N, nl, words] / cnt_s [N, nl] (eval_sample_stats) vs ground-truth label maps gts [M, H, W]
    -> double [4] = GED, sum d_sy, sum d_ss, sum d_yy (bit-identical to the reference's loops)"""
    n, nl, words = bits_s.shape
    m = gts.shape[0]
    dev = bits_s.device
    gts = gts.contiguous()
    lv = (ctypes.c_int * nl)(*[int(v) for v in label_values])
    bits_y = torch.empty((m, nl, words), dtype=torch.int32, device=dev)
    cnt_y = torch.empty((m, nl), dtype=torch.int32, device=dev)
    _lib.call('uz_ged_pack_masks', _p(gts), _DTYPE_CODE[gts.dtype], m, hw, lv, nl, _p(bits_y), _p(cnt_y), _stream())
    pair_d = torch.empty((n * m + n * n + m * m,), dtype=torch.float64, device=dev)
    out = torch.empty((4,), dtype=torch.float64, device=dev)
    _lib.call('uz_ged_pairwise', _p(bits_s), _p(cnt_s), n, _p(bits_y), _p(cnt_y), m, nl, hw, _p(pair_d), _p(out),
              _stream())
    return out


def ncc_dice_from_sums(sums, gts, n_total, dice_annotator=-1, out=None):
    """sums fp32 [2, C, HW] over n_total samples, gts [M, H, W] label maps -> double [1 + C] = NCC, per-class Dice of
    argmax(mean probs) vs annotator ``dice_annotator`` (NaN-free only when dice_annotator >= 0)"""
    _, c, hw = sums.shape
    m = gts.shape[0]
    dev = sums.device
    gts = gts.contiguous()
    work = torch.empty(((1 + m) * hw + m,), dtype=torch.float64, device=dev)
    dcnt = torch.empty((3 * c,), dtype=torch.int32, device=dev) if dice_annotator >= 0 else None
    if out is None:
        out = torch.zeros((1 + c,), dtype=torch.float64, device=dev)
    _lib.call('uz_ncc_dice_from_sums', _p(sums), _p(gts), _DTYPE_CODE[gts.dtype], n_total, c, hw, m, int(dice_annotator),
              _p(work), _p(dcnt), _p(out), _stream())
    return out


def argmax_classes(x):
    n, c, h, w = x.shape
    out = torch.empty((n, h, w), dtype=torch.uint8, device=x.device)
    _lib.call('uz_argmax_classes', _p(x.contiguous()), n, c, h * w, _p(out), _stream())
    return out


def variance_ncc(probs, gt_onehot):
    """probs fp32 [N,C,H,W], gt_onehot [M,C,H,W] -> double tensor [1]"""
    if not probs.is_cuda:
        raise _lib.UnetZooLibError('B200 path needs CUDA tensors (no CPU fallback)')
    probs = probs.contiguous().float()
    gt_onehot = gt_onehot.contiguous().to(probs.device)
    n, c, h, w = probs.shape
    m = gt_onehot.shape[0]
    hw = h * w
    work = torch.empty(((1 + m) * hw + m,), dtype=torch.float64, device=probs.device)
    out = torch.empty((1,), dtype=torch.float64, device=probs.device)
    _lib.call('uz_variance_ncc', _p(probs), _p(gt_onehot), _DTYPE_CODE[gt_onehot.dtype], n, c, hw, m, _p(work), _p(out),
              _stream())
    return out


def channel_sum(g):
    """per-channel sum over pixels of a bf16 NHWC tensor -> fp32 [C] (conv bias gradient of BN-free layers)."""
    n, h, w, c, ld = _check_act(g)
    npix = n * h * w
    nb = _lib.raw('uz_bn_bwd_num_blocks')(npix, c)
    partial = torch.empty((nb, 2, c), dtype=torch.float32, device=g.device)
    one = torch.ones(c, dtype=torch.float32, device=g.device)
    _lib.call('uz_bn_bwd_reduce', _p(g), ld, _p(g), ld, _p(one), _p(one), 0, npix, c, _p(partial), _stream())
    return partial[:, 0].sum(0)


class ZeroArena:
    """Zero-filled fp32 scratch handed out as views: gradients that are exactly zero by construction (conv bias in
    front of BatchNorm) and zero-initialised accumulators share ONE fill per step instead of one launch each.
    The memory is never written by this package; a fresh block is taken every step so stale views stay valid."""

    def __init__(self, floats=1 << 19):
        self.floats = floats
        self.buf = None
        self.off = 0

    def reset(self, device=None):
        """start a new step; with ``device`` the block is allocated (and zero-filled) right away on the CURRENT stream,
        i.e. before any side stream forks off, so every later user is ordered after the fill"""
        self.buf = None
        self.off = 0
        if device is not None:
            self.buf = torch.zeros(self.floats, dtype=torch.float32, device=device)

    def get(self, n, device):
        n_al = (n + 3) // 4 * 4
        if self.buf is None or self.buf.device != device or self.off + n_al > self.buf.numel():
            self.buf = torch.zeros(max(self.floats, n_al), dtype=torch.float32, device=device)
            self.off = 0
        out = self.buf[self.off:self.off + n]
        self.off += n_al
        return out


zero_arena = ZeroArena()
