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
                torch._foreach_add_(counts, 1)
        return False


def _boundary(fn):
    """forward(x) wrapper: NCHW fp32 in -> NCHW fp32 out, Act in -> Act out."""

    def wrapped(self, x, *args, **kwargs):
        if isinstance(x, Act):
            return fn(self, x, *args, **kwargs)
        return ops.from_act(fn(self, ops.to_act(x), *args, **kwargs))

    return wrapped


class Conv2D(nn.Module):
    """conv(k in {3 (pad 1), 1 (pad 0)}) + bias -> norm(eps 1e-3, momentum 0.01) -> activation."""

    def __init__(self, input_dim, output_dim, kernel_size=3, stride=1, padding=1, activation=torch.nn.ReLU,
                 norm=torch.nn.BatchNorm2d, norm_before_activation=True):
        super(Conv2D, self).__init__()
        if kernel_size not in (1, 3) or stride != 1:
            raise NotImplementedError('B200 Conv2D supports the shapes the reference uses: kernel 1 or 3, stride 1')
        padding = 1 if kernel_size == 3 else 0
        layers = [nn.Conv2d(input_dim, output_dim, kernel_size=kernel_size, stride=stride, padding=padding)]
        if norm_before_activation:
            layers.append(norm(num_features=output_dim, eps=1e-3, momentum=0.01))
            layers.append(activation())
        else:
            raise NotImplementedError('norm_before_activation=False is never used by the reference models')
        self.convolution = nn.Sequential(*layers)
        if not isinstance(self.convolution[1], (nn.BatchNorm2d, nn.Identity)) or \
                not isinstance(self.convolution[2], (nn.ReLU, nn.Identity)):
            raise NotImplementedError('B200 Conv2D supports norm in {BatchNorm2d, Identity}, activation in {ReLU, Identity}')
        self.input_dim = input_dim
        self.output_dim = output_dim

    @_boundary
    def forward(self, x):
        conv, bn, act = self.convolution[0], self.convolution[1], self.convolution[2]
        relu = isinstance(act, nn.ReLU)
        if self.output_dim % 16 != 0:
            raise NotImplementedError('Conv2D with %d output channels: use the fused logits path '
                                      '(b200.ops.SLayerNearest)' % self.output_dim)
        if isinstance(bn, nn.BatchNorm2d):
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


class ReversibleSequence(nn.Module):
    """Reversible stack (reference torchlayers.py:55-82 over revtorch): not built yet on the B200 path."""

    def __init__(self, input_dim, output_dim, reversible_depth=3, kernel=3):
        super(ReversibleSequence, self).__init__()
        raise NotImplementedError('reversible blocks (RevPHiSeg) are scheduled after the non-reversible hot path; '
                                  'see DESIGN.md "next"')
