"""torch.autograd.Functions over the C ABI.  They own the saved tensors and workspaces (SURVEY.md 8b "Ownership")
and keep the parameter / gradient tensors in the reference's fp32 NCHW / OIHW layouts so the caller's stock Adam
(reference train_model.py:49) keeps working.

``Act`` is the internal activation handle that flows between the drop-in modules: a bf16 NHWC tensor whose channel
count is padded to a multiple of 16, plus the logical channel count.
"""
import torch

from . import kern


class Act:
    """bf16 channel-last activation [N,H,W,Cp] or, for volumes, [N,D,H,W,Cp] (Cp = padded c); ``c`` = logical
    channels."""
    __slots__ = ('t', 'c')

    def __init__(self, t, c):
        self.t = t
        self.c = c

    @property
    def shape(self):
        return (self.t.shape[0], self.c) + tuple(self.t.shape[1:-1])


def _dense(t):
    """grads handed over by autograd may be arbitrary views; kernels need unit channel stride + packed pixels."""
    if t.dim() == 5:
        n, d, h, w, _ = t.shape
        ld = t.stride(3)
        if t.stride(4) == 1 and ld % 8 == 0 and t.data_ptr() % 16 == 0 and (h == 1 or t.stride(2) == w * ld) and \
                (d == 1 or t.stride(1) == h * w * ld) and (n == 1 or t.stride(0) == d * h * w * ld):
            return t
        return t.contiguous()
    if t.stride(3) == 1 and (t.shape[1] == 1 or t.stride(1) == t.shape[2] * t.stride(2)) and \
            (t.shape[0] == 1 or t.stride(0) == t.shape[1] * t.shape[2] * t.stride(2)) and t.stride(2) % 8 == 0 \
            and t.data_ptr() % 16 == 0:
        return t
    return t.contiguous()


# ------------------------------------------------------------------------------------------------ auxiliary stream
# The weight gradient of a layer depends only on (x, dy) and nothing downstream in the backward pass depends on it, so
# it is issued on an auxiliary stream next to the dgrad chain; the streams are joined once at the end of backward
# (autograd end-of-pass callback) or earlier by whoever needs the gradients (dp bucket hooks).  Small layers, whose
# persistent kernels occupy a fraction of the SMs, then overlap.  UNETZOO_CONCURRENCY=0 disables it.
import os as _os

_AUX_ENABLED = _os.environ.get('UNETZOO_CONCURRENCY', '1') != '0'
# planned SM share of an overlapped wgrad: measured with the splits accumulated in L2 and the full-resolution maps at
# 100 %: 25 / 28 / 30 / 32 / 35 % -> 4.03-4.04 ms, 40 / 45 / 50 / 70 % -> 4.08 ms, 100 % -> 4.17 ms
_WGRAD_SM_PERCENT = int(_os.environ.get('UNETZOO_WGRAD_SM_PERCENT', '30'))


def set_concurrency(enabled):
    """switch the auxiliary weight-gradient streams on / off; an overlapped launch is planned for a quarter of the SMs
    (fewer split-K partials), a launch that owns the GPU for all of them"""
    global _AUX_ENABLED
    _AUX_ENABLED = bool(enabled)
    _aux['planned_for'] = None


def _plan_wgrad_share(overlapped):
    if _aux.get('planned_for') is not overlapped:         # applied lazily: importing this module never touches the library
        kern._lib.call('uz_set_wgrad_sm_percent', _WGRAD_SM_PERCENT if overlapped else 100)
        _aux['planned_for'] = overlapped

_AUX_STREAMS = max(1, int(_os.environ.get('UNETZOO_AUX_STREAMS', '3')))      # weight-gradient streams (round robin)
_aux = {'streams': {}, 'pending': [], 'keep': [], 'callback_task': None, 'next': 0}


def _aux_stream(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(device).cuda_stream, _aux['next'] % _AUX_STREAMS)
    _aux['next'] += 1
    if key not in _aux['streams']:
        _aux['streams'][key] = torch.cuda.Stream(device=device)
    return _aux['streams'][key]


def pending_aux_streams():
    """auxiliary streams with launches that nobody has joined yet (the dp bucket hooks make their side stream wait)"""
    return list(_aux['pending'])


def sync_aux_streams():
    """make the current stream wait for every outstanding auxiliary-stream launch, then reduce the weight-gradient
    partials that are still pending (one launch)"""
    cur = torch.cuda.current_stream()
    for st in _aux['pending']:
        cur.wait_stream(st)
    _aux['pending'] = []
    kern.wgrad_reducer.flush(force=False)      # unless an optimizer that reads the slabs itself follows (TrainStep)
    _aux['keep'] = []
    _aux['callback_task'] = None
    _aux['next'] = 0


# Weight gradients leave their split-K partial slabs behind; one batched launch per ~_FLUSH_BYTES of slabs (on an
# auxiliary stream, next to the dgrad chain) and one at the end of backward reduce them into the OIHW gradients.  The
# per-layer reduction kernels were ~0.7 ms of a 8.3 ms single-stream PHiSeg step (106 launches writing with a 36-byte
# stride).  UNETZOO_DEFER_WGRAD_REDUCE=0 restores them.
_DEFER_REDUCE = _os.environ.get('UNETZOO_DEFER_WGRAD_REDUCE', '1') != '0'
_FLUSH_BYTES = int(_os.environ.get('UNETZOO_WGRAD_FLUSH_MB', '4096')) << 20
# (flushing earlier -- every 16 / 64 MB of slabs, or once when backward reaches the 64x64 maps -- on an auxiliary stream was
# measured slower: 4.78 / 4.62 / 4.42 ms against 4.58 / 4.58 / 4.37 ms with the single flush at the end of backward)


def _can_defer(weight):
    """the gradient tensor is handed to autograd before it is filled: only safe when nothing reads it before the end of
    backward -- no accumulation into an existing .grad, no tensor hooks on the parameter"""
    return _DEFER_REDUCE and weight.grad is None and not weight._backward_hooks


def _flush_partials_if_large():
    if kern.wgrad_reducer.pending_bytes < _FLUSH_BYTES or not _aux['pending']:
        return
    dev = kern.wgrad_reducer.items[0][0].device
    st = _aux_stream(dev)
    for other in _aux['pending']:
        if other is not st:
            st.wait_stream(other)
    with torch.cuda.stream(st):
        kern.wgrad_reducer.flush()
    if st not in _aux['pending']:
        _aux['pending'].append(st)


_BIG_MAP_PIXELS = int(_os.environ.get('UNETZOO_WGRAD_BIG_PIXELS', str(12 * 128 * 128)))
# the full-resolution maps come last in backward, when little else is left to overlap with: whole GPU (4.105 -> 4.047 ms)
_BIG_MAP_PERCENT = int(_os.environ.get('UNETZOO_WGRAD_BIG_PERCENT', '100'))


def _run_on_aux(fn, keep, last=False):
    """run fn() on the auxiliary stream after everything already queued on the current stream.  ``last``: nothing follows
    this layer on its backward chain (the first layer of a network: no input gradient) -- its weight gradient is the tail
    of the step and is planned for the whole GPU on the chain's own stream."""
    # Only a captured step is GPU-bound: an eagerly issued step is bound by the host, where the stream switches of this
    # function (~80 us per layer) cost more than the overlap can return -- eager launches stay on the caller's stream.
    overlapped = _AUX_ENABLED and torch.cuda.is_current_stream_capturing() and not last
    if overlapped and _BIG_MAP_PERCENT != _WGRAD_SM_PERCENT and keep[0].numel() // keep[0].shape[-1] >= _BIG_MAP_PIXELS:
        # the largest maps are processed at the end of backward, when little else is left to overlap with
        kern._lib.call('uz_set_wgrad_sm_percent', _BIG_MAP_PERCENT)
        _aux['planned_for'] = None
    else:
        _plan_wgrad_share(overlapped)
    task = torch._C._current_graph_task_id()
    if task != -1 and _aux['callback_task'] != task:
        # once per backward pass (keyed by its graph-task id, so a pass that died with an exception cannot leave the flag
        # set); also joins / flushes when nothing ran on an auxiliary stream
        _aux['callback_task'] = task
        torch.autograd.Variable._execution_engine.queue_callback(sync_aux_streams)
    if not overlapped:
        return fn()
    cur = torch.cuda.current_stream()
    st = _aux_stream(keep[0].device)
    st.wait_stream(cur)
    with torch.cuda.stream(st):
        out = fn()
    _aux['keep'].append((keep, out))          # inputs stay referenced until the join: no early reuse of their memory
    if st not in _aux['pending']:
        _aux['pending'].append(st)
    _flush_partials_if_large()
    return out


# ------------------------------------------------------------------------------------------------ layout boundary
class ToNHWC(torch.autograd.Function):
    """fp32 NCHW -> bf16 NHWC (padded).  Used at module boundaries and for z -> next conv."""

    @staticmethod
    def forward(ctx, x, ld=None):
        ctx.c = x.shape[1]
        return kern.nchw_to_nhwc(x, ld)

    @staticmethod
    def backward(ctx, g):
        return kern.nhwc_to_nchw(_dense(g), ctx.c), None


class FromNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, c):
        ctx.ld = t.shape[-1]
        return kern.nhwc_to_nchw(t, c)

    @staticmethod
    def backward(ctx, g):
        return kern.nchw_to_nhwc(g, ctx.ld), None


def to_act(x):
    if isinstance(x, Act):
        return x
    if not x.is_cuda:
        raise kern._lib.UnetZooLibError('UNet-Zoo B200 modules need CUDA tensors: there is no CPU fallback path')
    return Act(ToNHWC.apply(x.float(), kern.pad_channels(x.shape[1], x.dim())), x.shape[1])


def from_act(a):
    return FromNHWC.apply(a.t, a.c)


# ------------------------------------------------------------------------------------------------ conv (+BN) (+ReLU)
# Fusing the BatchNorm-backward reduction of layer L into the input-gradient conv of layer L+1 (opt-in: UNETZOO_FUSE_BN_BWD=1;
# measured on B200 it removes 57 launches from the PHiSeg step but the extra y loads lengthen the epilogue-bound dgrad
# kernels by more than the removed reductions cost: 5.07 ms vs 4.96 ms per step, profiles/r02_knobs.md):
# ConvBNAct.forward tags its output tensor with (y, scale, shift, relu); a ConvBNAct that consumes a tagged tensor asks its
# dgrad launch to mask the result with that ReLU and to accumulate (sum g, sum g*y) in the epilogue, and tags the gradient
# it returns with those sums.  The producer's backward uses them only if the very same tensor arrives unmodified (autograd
# accumulates gradients of multi-consumer tensors in place: the version counter tells) -- otherwise it runs the separate
# reduction kernel as before.  The mask is idempotent, so a masked gradient is always safe to hand on.
_FUSE_BN_BWD = _os.environ.get('UNETZOO_FUSE_BN_BWD', '0') == '1'


def set_fuse_bn_backward(enabled):
    global _FUSE_BN_BWD
    prev = _FUSE_BN_BWD
    _FUSE_BN_BWD = bool(enabled)
    return prev


def _fused_sums_of(da, c):
    tag = getattr(da, '_uz_sums', None)
    if tag is None:
        return None
    sums, version, ptr = tag
    if da._version != version or da.data_ptr() != ptr or sums.shape[-1] != c:
        return None
    return sums


def _bucket_view(weight):
    """data-parallel training: the flat-bucket view this parameter's gradient is all-reduced in (b200.dp), or None"""
    from . import dp
    return dp.grad_view_for(weight)


class ConvBNAct(torch.autograd.Function):
    """Conv2D of the reference (torchlayers.py:7-29): conv(k=3 pad 1 | k=1) + bias -> BatchNorm(train) -> ReLU.

    forward (training): tcgen05 conv with statistics epilogue -> uz_bn_apply_train (finalize + normalise + ReLU); maps of
    up to 1024 pixels (2x2 ... 8x8 at batch 12) take ONE launch (uz_conv_bn_act_fused: statistics exchanged through the
    distributed shared memory of a thread-block cluster).
    backward: BN/ReLU backward (reduction fused into the consumer's dgrad epilogue when possible, else its own launch;
    then the apply pass) -> dgrad (same conv kernel, flipped weights) + tcgen05 wgrad.
    Deterministic mode (kern.set_deterministic): statistics as per-CTA rows reduced in fixed order by uz_bn_finalize /
    uz_bn_bwd_finalize, no atomics, no fusion.
    The conv bias gets a zero gradient: BatchNorm removes any per-channel constant, d loss / d bias == 0 exactly.
    """

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, running_mean, running_var, relu, cin_logical):
        need_dx = ctx.needs_input_grad[0]
        wf, wd = kern.pack_conv_weight(weight, need_dgrad=need_dx)
        ctx.det = kern.is_deterministic() and x.dim() == 4
        npix = kern._spatial_numel(x.shape[:-1])
        if kern.conv_bn_fused_supported(x, wf):
            # small maps: conv + batch statistics (exchanged inside a thread-block cluster, fixed order) + normalise + ReLU
            # in ONE launch
            a, y, scale, shift, mean, invstd = kern.conv_bn_act_fused(x, wf, bias, gamma, beta, running_mean, running_var,
                                                                      relu=relu)
        else:
            y, sums = kern.conv_fwd(x, wf, shift=bias, stats=True)
            if ctx.det:
                scale, shift, mean, invstd = kern.bn_finalize(sums, npix, gamma, beta, running_mean, running_var)
                a = kern.affine_act(y, scale, shift, relu=relu)
            else:
                a, scale, shift, mean, invstd = kern.bn_apply_train(y, sums, npix, gamma, beta, running_mean, running_var,
                                                                    relu=relu)
        src = getattr(x, '_uz_bn', None) if (need_dx and _FUSE_BN_BWD and not ctx.det and x.dim() == 4) else None
        ctx.fuse_src = src is not None
        if src is not None:
            ctx.save_for_backward(x, y, wd, scale, shift, mean, invstd, gamma, src[0], src[1], src[2])
            ctx.src_relu = src[3]
        else:
            ctx.save_for_backward(x, y, wd, scale, shift, mean, invstd, gamma)
        ctx.relu = relu
        ctx.cin_logical = cin_logical
        ctx.wshape = weight.shape
        ctx.weight_ref = weight
        if _FUSE_BN_BWD and not ctx.det and x.dim() == 4:
            a._uz_bn = (y, scale, shift, relu)
        return a

    @staticmethod
    def backward(ctx, da):
        if ctx.fuse_src:
            x, y, wd, scale, shift, mean, invstd, gamma, y_prev, sc_prev, sh_prev = ctx.saved_tensors
        else:
            x, y, wd, scale, shift, mean, invstd, gamma = ctx.saved_tensors
        sums = _fused_sums_of(da, y.shape[-1])
        if sums is None:
            da = _dense(da)
        if ctx.det:
            dy, dgamma, dbeta = kern.bn_relu_bwd(da, y, scale, shift, gamma, mean, invstd, relu=ctx.relu)
        else:
            dy, dgamma, dbeta = kern.bn_relu_bwd_train(da, y, scale, shift, gamma, mean, invstd, relu=ctx.relu, sums=sums)
        dx = None
        if ctx.needs_input_grad[0]:
            if ctx.fuse_src:
                dx, psums = kern.conv_fwd(dy, wd, bn_prev=(y_prev, sc_prev, sh_prev, ctx.src_relu))
                dx._uz_sums = (psums, dx._version, dx.data_ptr())
            else:
                dx, _ = kern.conv_fwd(dy, wd)
        cout, cin = ctx.wshape[0], ctx.wshape[1]
        taps = kern._spatial_numel(ctx.wshape[2:])
        defer = _can_defer(ctx.weight_ref)
        dw = _run_on_aux(lambda: kern.conv_wgrad(x, dy, taps, cin, cout, out=_bucket_view(ctx.weight_ref), defer=defer),
                         (x, dy), last=not ctx.needs_input_grad[0]).view(ctx.wshape)
        dbias = kern.zero_arena.get(cout, dy.device)
        return dx, dw, dbias, dgamma, dbeta, None, None, None, None


class ConvAffineAct(torch.autograd.Function):
    """Inference-time Conv2D (BatchNorm running statistics folded into the conv epilogue) and the BN-free
    conv + bias (+ReLU) of models/unet.py:25-30 (scale = 1, shift = bias): ONE kernel per layer."""

    @staticmethod
    def forward(ctx, x, weight, scale, shift, relu, cin_logical, bias_is_shift):
        need_dx = ctx.needs_input_grad[0]
        need_bwd = need_dx or ctx.needs_input_grad[1]
        wf, wd = kern.pack_conv_weight(weight, need_dgrad=need_dx)
        a, _ = kern.conv_fwd(x, wf, scale=scale, shift=shift, relu=relu)
        if need_bwd:
            ctx.save_for_backward(x, a, wd, scale)
        ctx.relu = relu
        ctx.cin_logical = cin_logical
        ctx.wshape = weight.shape
        ctx.weight_ref = weight
        ctx.bias_is_shift = bias_is_shift
        return a

    @staticmethod
    def backward(ctx, da):
        x, a, wd, scale = ctx.saved_tensors
        if scale is not None:
            raise NotImplementedError('backward through eval-mode (folded) BatchNorm is not on the hot path; '
                                      'call net.train() for training (reference train_model.py:95)')
        da = _dense(da)
        cout, cin = ctx.wshape[0], ctx.wshape[1]
        taps = kern._spatial_numel(ctx.wshape[2:])
        one = torch.ones(cout, dtype=torch.float32, device=da.device)
        zero = torch.zeros_like(one)
        if ctx.relu:
            dy = kern.relu_bwd(da, a, one, zero)
        else:
            dy = da
        dx = None
        if ctx.needs_input_grad[0]:
            dx, _ = kern.conv_fwd(dy, wd)
        dw = _run_on_aux(lambda: kern.conv_wgrad(x, dy, taps, cin, cout, defer=_can_defer(ctx.weight_ref)),
                         (x, dy)).view(ctx.wshape)
        dshift = None
        if ctx.bias_is_shift and ctx.needs_input_grad[3]:
            dshift = kern.channel_sum(dy)
        return dx, dw, None, dshift, None, None, None


# ------------------------------------------------------------------------------------------------ pooling / upsampling / concat
class AvgPool2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return kern.avgpool2_fwd(x)

    @staticmethod
    def backward(ctx, g):
        return kern.avgpool2_bwd(_dense(g))


class Upsample2x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, align_corners):
        ctx.align = align_corners
        return kern.upsample2x_fwd(x, align_corners)

    @staticmethod
    def backward(ctx, g):
        return kern.upsample2x_bwd(_dense(g), ctx.align), None


class Concat(torch.autograd.Function):
    """torch.cat([a, b], dim=channels) on NHWC; with up_b the second operand is bilinearly upsampled x2 on the fly
    (Likelihood top-down path, models/phiseg.py:304-315; U-Net decoder, models/unet.py:67-72)."""

    @staticmethod
    def forward(ctx, a, b, up_a, up_b, align_corners):
        ca, cb = a.shape[-1], b.shape[-1]
        sp = tuple(k * (2 if up_a else 1) for k in a.shape[1:-1])
        out = torch.empty((a.shape[0],) + sp + (ca + cb,), dtype=kern._lib.act_dtype(), device=a.device)
        for src, up, sl in ((a, up_a, out[..., :ca]), (b, up_b, out[..., ca:])):
            if up:
                kern.upsample2x_fwd(src, align_corners, out=sl)
            else:
                kern.copy_channels(src, sl)
        ctx.cfg = (ca, up_a, up_b, align_corners)
        return out

    @staticmethod
    def backward(ctx, g):
        ca, up_a, up_b, align = ctx.cfg
        g = _dense(g)
        ga, gb = g[..., :ca], g[..., ca:]
        da = kern.upsample2x_bwd(ga, align) if up_a else ga
        db = kern.upsample2x_bwd(gb, align) if up_b else gb
        return da, db, None, None, None


class GlobalMean(torch.autograd.Function):
    """spatial mean [N,H,W,C] -> [N,1,1,C] (models/probabilistic_unet.py:114-115)"""

    @staticmethod
    def forward(ctx, x):
        ctx.hw = (x.shape[1], x.shape[2])
        return kern.global_mean_fwd(x)

    @staticmethod
    def backward(ctx, g):
        return kern.global_mean_bwd(_dense(g), *ctx.hw)


# ------------------------------------------------------------------------------------------------ latent head / losses
class LatentHead(torch.autograd.Function):
    """mu, sigma = softplus(.), z = mu + sigma * eps from NHWC features (models/phiseg.py:95-106)."""

    @staticmethod
    def forward(ctx, feat, wmu, bmu, wsig, bsig, eps):
        c = feat.shape[-1]
        wm = wmu.reshape(wmu.shape[0], -1)
        ws = wsig.reshape(wsig.shape[0], -1)
        if wm.shape[1] != c:          # features are channel padded
            wm = torch.nn.functional.pad(wm, (0, c - wm.shape[1]))
            ws = torch.nn.functional.pad(ws, (0, c - ws.shape[1]))
        wm, ws = wm.contiguous(), ws.contiguous()
        mu, sigma, z = kern.head_fwd(feat, wm, bmu, ws, bsig, eps)
        ctx.save_for_backward(feat, wm, ws, eps, sigma)
        ctx.wshape = wmu.shape
        return mu, sigma, z

    @staticmethod
    def backward(ctx, dmu, dsigma, dz):
        feat, wm, ws, eps, sigma = ctx.saved_tensors
        cont = lambda t: None if t is None else t.contiguous()
        dfeat, dw, db = kern.head_bwd(feat, wm, ws, eps, sigma, cont(dmu), cont(dsigma), cont(dz))
        zdim, cl = ctx.wshape[0], ctx.wshape[1]
        dwmu = dw[:zdim, :cl].reshape(ctx.wshape)
        dwsig = dw[zdim:, :cl].reshape(ctx.wshape)
        return dfeat, dwmu, db[:zdim], dwsig, db[zdim:], None


class KLLevel(torch.autograd.Function):
    """weight * KL_two_gauss_with_diag_cov (models/phiseg.py:436-453,463-472)."""

    @staticmethod
    def forward(ctx, mu0, s0, mu1, s1, weight):
        mu0, s0, mu1, s1 = (t.contiguous() for t in (mu0, s0, mu1, s1))
        ctx.save_for_backward(mu0, s0, mu1, s1)
        ctx.weight = weight
        return kern.kl_fwd(mu0, s0, mu1, s1, weight).reshape(())

    @staticmethod
    def backward(ctx, up):
        mu0, s0, mu1, s1 = ctx.saved_tensors
        g = kern.kl_bwd(mu0, s0, mu1, s1, ctx.weight, up.reshape(1).contiguous())
        return g[0], g[1], g[2], g[3], None


class KLHierarchy(torch.autograd.Function):
    """calculate_hierarchical_KL_div_loss (models/phiseg.py:463-472) for all latent levels at once: returns
    (sum_l total_weight * w_l * KL_l  -- added in the reference's order l = L-1 ... 0 --, per-level w_l * KL_l for the
    loss dictionary).  Two launches forward, one backward, instead of four / two per level."""

    @staticmethod
    def forward(ctx, level_weights, total_weight, *tensors):
        ts = [t.contiguous() for t in tensors]
        levels = [tuple(ts[i:i + 4]) for i in range(0, len(ts), 4)]
        total, per_level = kern.kl_hierarchy_fwd(levels, level_weights, total_weight)
        ctx.save_for_backward(*ts)
        ctx.cfg = (tuple(level_weights), float(total_weight))
        ctx.mark_non_differentiable(per_level)
        return total.reshape(()), per_level

    @staticmethod
    def backward(ctx, up, _unused):
        ts = ctx.saved_tensors
        levels = [tuple(ts[i:i + 4]) for i in range(0, len(ts), 4)]
        grads = kern.kl_hierarchy_bwd(levels, ctx.cfg[0], ctx.cfg[1], up.reshape(1).contiguous().float())
        return (None, None) + tuple(grads)


class SLayerNearest(torch.autograd.Function):
    """1x1 conv to class logits + nearest upsample to the image size (models/phiseg.py:319-321)."""

    @staticmethod
    def forward(ctx, feat, weight, bias, factor):
        c = feat.shape[-1]
        w2 = weight.reshape(weight.shape[0], -1)
        if w2.shape[1] != c:
            w2 = torch.nn.functional.pad(w2, (0, c - w2.shape[1]))
        w2 = w2.contiguous()
        ctx.save_for_backward(feat, w2)
        ctx.factor = factor
        ctx.wshape = weight.shape
        return kern.slayer_fwd(feat, w2, bias, factor)

    @staticmethod
    def backward(ctx, g):
        feat, w2 = ctx.saved_tensors
        dfeat, dw, db = kern.slayer_bwd(g, feat, w2, ctx.factor)
        return dfeat, dw[:, :ctx.wshape[1]].reshape(ctx.wshape), db, None


class ResidualCE(torch.autograd.Function):
    """residual_multinoulli_loss (models/phiseg.py:492-513): returns (sum over levels, per-level values).  The
    gradient kernel is the same fused pass re-run in backward with the upstream scalar folded in."""

    @staticmethod
    def forward(ctx, target, *s_list):
        s_list = [s.contiguous() for s in s_list]
        need = any(ctx.needs_input_grad[1:])
        if need:
            # ONE pass: the gradients for an upstream of 1 come out of the same kernel run as the values (they cost 8 B per
            # logit to keep); backward only scales them.  The second, gradient-only run sat alone on the step's critical
            # path between forward and backward (~55 us of an otherwise idle GPU).
            ce, grads = kern.residual_ce(s_list, target, need_grad=True, upstream=kern.one_scalar(s_list[0].device))
            ctx.grads = grads
        else:
            ce, _ = kern.residual_ce(s_list, target, need_grad=False)
        total = ce.sum()
        ctx.mark_non_differentiable(ce)
        return total, ce

    @staticmethod
    def backward(ctx, up, _unused):
        return (None,) + tuple(torch._foreach_mul(ctx.grads, up.reshape(()).float()))
