"""Drop-in for the reference's ``torchlayers.py`` (Conv2D :7-29, Conv2DSequence :32-52, ReversibleSequence :55-82):
same class names, constructor arguments, sub-module names (=> identical state_dict keys and default initialisation,
parameters stay fp32 NCHW nn.Parameters owned by stock torch optimisers) -- but ``forward`` runs on the B200 kernels
of libunetzoo_b200.so (tcgen05 implicit-GEMM conv + fused BatchNorm/ReLU passes).

Every module accepts either a plain fp32 NCHW CUDA tensor (returns one, like the reference) or the internal
``b200.ops.Act`` handle (bf16 NHWC, returns an Act) so stacks of modules never leave NHWC.  No CPU path exists.
"""
import torch
import torch.nn as nn

from b200 import kern, ops
from b200.ops import Act


# BatchNorm.num_batches_tracked bookkeeping: inside ``deferred_batch_counts()`` the per-layer ``+= 1`` launches are
# collected and applied with a single multi-tensor add when the context exits (106 launches -> 1 per PHiSeg forward).
_deferred_counts = None


class deferred_batch_counts:
    def __enter__(self):
        global _deferred_counts
        self.prev = _deferred_counts
        _deferred_counts = []
        return self

    def __exit__(self, *exc):
        global _deferred_counts
        counts, _deferred_counts = _deferred_counts, self.prev
        if counts:
            if self.prev is not None:
                self.prev.extend(counts)
            else:
                # a counter may appear more than once (reversible blocks run F and G twice per step): one add per
                # multiplicity, every tensor at most once per multi-tensor launch
                seen = {}
                for t in counts:
                    seen.setdefault(id(t), [t, 0])[1] += 1
                for mult in sorted(set(n for _, n in seen.values())):
                    torch._foreach_add_([t for t, n in seen.values() if n == mult], mult)
        return False


def _boundary(fn):
    """forward(x) wrapper: NCHW fp32 in -> NCHW fp32 out, Act in -> Act out."""

    def wrapped(self, x, *args, **kwargs):
        if isinstance(x, Act):
            return fn(self, x, *args, **kwargs)
        return ops.from_act(fn(self, ops.to_act(x), *args, **kwargs))

    return wrapped


class Conv2D(nn.Module):
    """conv(k in {3 (pad 1), 1 (pad 0)}) + bias -> norm(eps 1e-3, momentum 0.01) -> activation.
    ``_conv_cls`` / ``_norm_cls`` / ``_granule`` let models/phiseg3D.py derive its Conv3D from the same code."""
    _conv_cls = nn.Conv2d
    _norm_cls = nn.BatchNorm2d
    _granule = 16            # output channels the tensor-core path stores per group

    def __init__(self, input_dim, output_dim, kernel_size=3, stride=1, padding=1, activation=torch.nn.ReLU,
                 norm=None, norm_before_activation=True):
        super(Conv2D, self).__init__()
        if norm is None:
            norm = self._norm_cls
        if kernel_size not in (1, 3) or stride != 1:
            raise NotImplementedError('B200 conv layers support the shapes the reference uses: kernel 1 or 3, stride 1')
        padding = 1 if kernel_size == 3 else 0
        layers = [self._conv_cls(input_dim, output_dim, kernel_size=kernel_size, stride=stride, padding=padding)]
        if norm_before_activation:
            layers.append(norm(num_features=output_dim, eps=1e-3, momentum=0.01))
            layers.append(activation())
        else:
            raise NotImplementedError('norm_before_activation=False is never used by the reference models')
        self.convolution = nn.Sequential(*layers)
        if not isinstance(self.convolution[1], (self._norm_cls, nn.Identity)) or \
                not isinstance(self.convolution[2], (nn.ReLU, nn.Identity)):
            raise NotImplementedError('B200 conv layers support norm in {BatchNorm, Identity}, activation in {ReLU, Identity}')
        self.input_dim = input_dim
        self.output_dim = output_dim

    @_boundary
    def forward(self, x):
        conv, bn, act = self.convolution[0], self.convolution[1], self.convolution[2]
        relu = isinstance(act, nn.ReLU)
        if self.output_dim % self._granule != 0:
            raise NotImplementedError('%s with %d output channels (needs a multiple of %d): use the fused logits path '
                                      '(b200.ops.SLayerNearest)' % (type(self).__name__, self.output_dim, self._granule))
        if isinstance(bn, self._norm_cls):
            if self.training:
                t = ops.ConvBNAct.apply(x.t, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean,
                                        bn.running_var, relu, x.c)
                if _deferred_counts is not None:
                    _deferred_counts.append(bn.num_batches_tracked)
                else:
                    bn.num_batches_tracked.add_(1)
            else:
                scale, shift = kern.bn_eval_fold(conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var)
                t = ops.ConvAffineAct.apply(x.t, conv.weight, scale, shift, relu, x.c, False)
        else:
            t = ops.ConvAffineAct.apply(x.t, conv.weight, None, conv.bias, relu, x.c, True)
        return Act(t, self.output_dim)


class Conv2DSequence(nn.Module):
    """depth x Conv2D (reference torchlayers.py:32-52)."""

    def __init__(self, input_dim, output_dim, kernel=3, depth=2, activation=torch.nn.ReLU, norm=torch.nn.BatchNorm2d,
                 norm_before_activation=True):
        super(Conv2DSequence, self).__init__()
        assert depth >= 1
        padding = 1 if kernel == 3 else 0
        layers = [Conv2D(input_dim, output_dim, kernel_size=kernel, padding=padding, activation=activation, norm=norm)]
        for i in range(depth - 1):
            layers.append(Conv2D(output_dim, output_dim, kernel_size=kernel, padding=padding, activation=activation,
                                 norm=norm))
        self.convolution = nn.Sequential(*layers)

    @_boundary
    def forward(self, x):
        for layer in self.convolution:
            x = layer(x)
        return x


# Fused reversible path (UNETZOO_FUSED_REVERSIBLE=0 selects the autograd-nested path below, which deterministic mode
# also uses): per F / G unit
#   forward   conv (+ statistics)  ->  ONE pass: BatchNorm + ReLU + coupling add written straight into the block output
#   backward  conv (recompute)     ->  ONE pass: BatchNorm-backward sums + coupling inverse (x2 = y2 - G(y1))
#             -> BatchNorm-backward apply -> dgrad whose epilogue adds the incoming gradient (dx1 = dy1 + dG/dy1)
#             -> wgrad (auxiliary stream)
# i.e. 2 launches forward and 4 on the backward chain per unit instead of 3 and 7, no nested torch.autograd.backward, no
# intermediate F(x2) / G(y1) tensors.  The BatchNorm statistics of the recomputation are the forward pass's (the inputs
# are the same up to bf16 reconstruction rounding); the second momentum update of the running statistics that revtorch's
# recomputation causes is applied together with the first.
import os as _os
_FUSED_REVERSIBLE = _os.environ.get('UNETZOO_FUSED_REVERSIBLE', '1') != '0'


def set_fused_reversible(enabled):
    global _FUSED_REVERSIBLE
    prev = _FUSED_REVERSIBLE
    _FUSED_REVERSIBLE = bool(enabled)
    return prev


def _count_batches(bn, n):
    for _ in range(n):
        if _deferred_counts is not None:
            _deferred_counts.append(bn.num_batches_tracked)
        else:
            bn.num_batches_tracked.add_(1)


class ReversibleBlock(nn.Module):
    """Additive coupling of revtorch 0.2.0 (reference torchlayers.py:67-75): y1 = x1 + F(x2), y2 = x2 + G(y1) on the two
    channel halves.  Sub-module names ``f_block`` / ``g_block`` are revtorch's (state_dict keys)."""

    def __init__(self, f_block, g_block):
        super(ReversibleBlock, self).__init__()
        self.f_block = f_block
        self.g_block = g_block

    # ------------------------------------------------------------------------------------------ fused path
    def fusable(self):
        if not _FUSED_REVERSIBLE or kern.is_deterministic():
            return False
        for seq in (self.f_block, self.g_block):
            if len(seq) != 1 or not isinstance(seq[0], Conv2D):
                return False
            m = seq[0]
            if not isinstance(m.convolution[1], m._norm_cls) or not isinstance(m.convolution[2], nn.ReLU):
                return False
        return True

    @staticmethod
    def _unit_forward(m, xin, res, out, updates):
        """out = res + ReLU(BatchNorm(conv(xin))); returns the statistics the backward pass needs (training mode)"""
        conv, bn = m.convolution[0], m.convolution[1]
        wf, _ = kern.pack_conv_weight(conv.weight, need_dgrad=True)
        if m.training:
            y, sums = kern.conv_fwd(xin, wf, shift=conv.bias, stats=True)
            npix = kern._spatial_numel(xin.shape[:-1])
            _, scale, shift, mean, invstd = kern.bn_apply_train(y, sums, npix, bn.weight, bn.bias, bn.running_mean,
                                                                bn.running_var, relu=True, residual=res, res_sign=1,
                                                                out=out, stat_updates=updates)
            _count_batches(bn, updates)
            return scale, shift, mean, invstd
        scale, shift = kern.bn_eval_fold(conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var)
        kern.conv_fwd(xin, wf, out=out, scale=scale, shift=shift, relu=True, residual=res, res_sign=1)
        return None

    def couple_fused(self, x, updates=1):
        c = x.shape[-1]
        half = c // 2
        y = kern._like(x, c)
        x1, x2, y1, y2 = x[..., :half], x[..., half:], y[..., :half], y[..., half:]
        sf = self._unit_forward(self.f_block[0], x2, x1, y1, updates)           # y1 = x1 + F(x2)
        sg = self._unit_forward(self.g_block[0], y1, x2, y2, updates)           # y2 = x2 + G(y1)
        return y, (sf, sg)

    @staticmethod
    def _unit_backward(m, xin, stats, g, inv, dres, grads):
        """unit u = ReLU(BatchNorm(conv(xin))) with upstream gradient g: recompute, inv[1] = inv[0] - u(xin),
        dres[1] = dres[0] + du/dxin, parameter gradients into ``grads``"""
        conv, bn = m.convolution[0], m.convolution[1]
        scale, shift, mean, invstd = stats
        wf, wd = kern.pack_conv_weight(conv.weight, need_dgrad=True)
        y, _ = kern.conv_fwd(xin, wf, shift=conv.bias)
        dy, dgamma, dbeta = kern.bn_relu_bwd_train(g, y, scale, shift, bn.weight, mean, invstd, relu=True, inverse=inv)
        kern.conv_fwd(dy, wd, out=dres[1], residual=dres[0], res_sign=1)
        wshape = conv.weight.shape
        cout, cin = wshape[0], wshape[1]
        taps = kern._spatial_numel(wshape[2:])
        defer = ops._can_defer(conv.weight)
        dw = ops._run_on_aux(lambda: kern.conv_wgrad(xin, dy, taps, cin, cout, out=ops._bucket_view(conv.weight),
                                                     defer=defer), (xin, dy)).view(wshape)
        grads[id(conv.weight)] = dw
        grads[id(conv.bias)] = kern.zero_arena.get(cout, dy.device)      # exactly zero in front of BatchNorm
        grads[id(bn.weight)] = dgamma
        grads[id(bn.bias)] = dbeta

    def backward_fused(self, y, dy, stats, grads):
        c = y.shape[-1]
        half = c // 2
        x = kern._like(y, c)
        dx = kern._like(y, c)
        y1, y2, dy1, dy2 = y[..., :half], y[..., half:], dy[..., :half], dy[..., half:]
        x1, x2, dx1, dx2 = x[..., :half], x[..., half:], dx[..., :half], dx[..., half:]
        sf, sg = stats
        self._unit_backward(self.g_block[0], y1, sg, dy2, (y2, x2), (dy1, dx1), grads)     # x2 = y2 - G(y1); dx1 = dy1 + ...
        self._unit_backward(self.f_block[0], x2, sf, dx1, (y1, x1), (dy2, dx2), grads)     # x1 = y1 - F(x2); dx2 = dy2 + ...
        return x, dx

    # ------------------------------------------------------------------------------------------ autograd-nested path
    def couple(self, x):
        """x: bf16 NHWC [N,H,W,C] -> y of the same shape; no autograd (callers handle gradients by inversion)."""
        c = x.shape[-1]
        half = c // 2
        y = kern._like(x, c)
        x1, x2, y1, y2 = x[..., :half], x[..., half:], y[..., :half], y[..., half:]
        fx2 = self.f_block(Act(x2, half)).t
        kern.add_channels(x1, fx2, y1)                               # y1 = x1 + F(x2)
        gy1 = self.g_block(Act(y1, half)).t
        kern.add_channels(x2, gy1, y2)                               # y2 = x2 + G(y1)
        return y

    def backward_pass(self, y, dy):
        """Inverse recompute (revtorch ReversibleBlock.backward_pass): returns (x, dx); parameter gradients are
        accumulated into ``.grad`` by the inner backward calls, F and G run a second time (BatchNorm running
        statistics receive their second momentum update)."""
        c = y.shape[-1]
        half = c // 2
        x = kern._like(y, c)
        dx = kern._like(y, c)
        y1, y2, dy1, dy2 = y[..., :half], y[..., half:], dy[..., :half], dy[..., half:]
        x1, x2, dx1, dx2 = x[..., :half], x[..., half:], dx[..., :half], dx[..., half:]
        y1_leaf = y1.detach().requires_grad_(True)
        with torch.enable_grad():
            gy1 = self.g_block(Act(y1_leaf, half)).t
        torch.autograd.backward(gy1, dy2)
        kern.add_channels(y2, gy1.detach(), x2, sign=-1)             # x2 = y2 - G(y1)
        kern.add_channels(dy1, ops._dense(y1_leaf.grad), dx1)        # dx1 = dy1 + dG/dy1
        x2_leaf = x2.detach().requires_grad_(True)
        with torch.enable_grad():
            fx2 = self.f_block(Act(x2_leaf, half)).t
        torch.autograd.backward(fx2, dx1)
        kern.add_channels(y1, fx2.detach(), x1, sign=-1)             # x1 = y1 - F(x2)
        kern.add_channels(dy2, ops._dense(x2_leaf.grad), dx2)        # dx2 = dy2 + dF/dx2
        return x, dx


class _ReversibleFunction(torch.autograd.Function):
    """Keeps only the output of the block chain; backward walks the blocks in reverse, regenerating their inputs."""

    @staticmethod
    def forward(ctx, x, blocks, *params):
        y = x
        ctx.fused = all(b.fusable() for b in blocks) and all(b.f_block[0].training for b in blocks)
        ctx.stats = []
        for block in blocks:
            if ctx.fused:
                y, st = block.couple_fused(y, updates=2)
                ctx.stats.append(st)
            else:
                y = block.couple(y)
        ctx.blocks = blocks
        ctx.params = params
        ctx.y = y
        ctx.packer = kern._active_packer          # weights do not change between forward and backward of a step
        return y

    @staticmethod
    def backward(ctx, dy):
        y = ctx.y
        del ctx.y
        dy = ops._dense(dy)
        prev = kern.set_active_packer(ctx.packer)
        grads = {}
        try:
            blocks = list(ctx.blocks)
            for k in range(len(blocks) - 1, -1, -1):
                if ctx.fused:
                    y, dy = blocks[k].backward_fused(y, dy, ctx.stats[k], grads)
                else:
                    y, dy = blocks[k].backward_pass(y, dy)
        finally:
            kern.set_active_packer(prev)
        return (dy, None) + tuple(grads.get(id(p)) for p in ctx.params)


class _RevtorchSequence(nn.Module):
    """Stands where ``rv.ReversibleSequence`` does in the reference (attribute ``reversible_blocks``)."""

    def __init__(self, reversible_blocks):
        super(_RevtorchSequence, self).__init__()
        self.reversible_blocks = reversible_blocks

    def forward(self, x):
        if torch.is_grad_enabled():
            t = _ReversibleFunction.apply(x.t, self.reversible_blocks, *self.reversible_blocks.parameters())
        else:
            t = x.t
            for block in self.reversible_blocks:
                if block.fusable():
                    t, _ = block.couple_fused(t, updates=1)
                else:
                    t = block.couple(t)
        return Act(t, x.c)


class ReversibleSequence(nn.Module):
    """Reversible stack of the reference (torchlayers.py:55-82): optional 1x1 Conv2D to change the channel count, then
    ``reversible_depth`` additive-coupling blocks whose F and G are 3x3 Conv2D on half the channels.  Only the stack's
    output is kept for backward; inputs are regenerated block by block (activation memory ~ 1/depth)."""

    _conv_layer = Conv2D

    def __init__(self, input_dim, output_dim, reversible_depth=3, kernel=3):
        super(ReversibleSequence, self).__init__()
        Conv = self._conv_layer
        if output_dim % (2 * Conv._granule) != 0:
            raise NotImplementedError('reversible stacks need channel halves that are multiples of %d on the B200 path'
                                      % Conv._granule)
        if input_dim != output_dim:
            self.inital_conv = Conv(input_dim, output_dim, kernel_size=1)
        else:
            self.inital_conv = nn.Identity()
        blocks = []
        for i in range(reversible_depth):
            f_func = nn.Sequential(Conv(output_dim // 2, output_dim // 2, kernel_size=kernel, padding=1))
            g_func = nn.Sequential(Conv(output_dim // 2, output_dim // 2, kernel_size=kernel, padding=1))
            blocks.append(ReversibleBlock(f_func, g_func))
        self.sequence = _RevtorchSequence(nn.ModuleList(blocks))

    @_boundary
    def forward(self, x):
        x = self.inital_conv(x)
        return self.sequence(x)
