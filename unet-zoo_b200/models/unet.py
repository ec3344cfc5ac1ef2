"""Drop-in for the reference's ``models/unet.py`` (DownConvBlock :12-40, UpConvBlock :43-75, Unet :78-165) on B200.

Same classes, constructor keywords, attribute names and state_dict keys (``contracting_path.N.layers.K.*`` bare
``nn.Conv2d`` + ``nn.ReLU`` entries, ``upsampling_path.N.conv_block.layers.K.*``, ``last_layer.*``).  The U-Net has no
BatchNorm: every layer is ONE tensor-core conv launch with bias + ReLU in the epilogue; the decoder's bilinear x2
(align_corners=False, unet.py:67) is written straight into the concat buffer ([up, bridge] order, unet.py:72).
"""
import torch
import torch.nn as nn

from b200 import kern, ops
from b200.ops import Act
from torchlayers import ReversibleSequence, _boundary
from utils import init_weights


class DownConvBlock(nn.Module):
    def __init__(self, input_dim, output_dim, initializers, padding, pool=True, reversible=False):
        super(DownConvBlock, self).__init__()
        layers = []
        if pool:
            layers.append(nn.AvgPool2d(kernel_size=2, stride=2, padding=0, ceil_mode=True))
        if not reversible:
            layers.append(nn.Conv2d(input_dim, output_dim, kernel_size=3, stride=1, padding=int(padding)))
            layers.append(nn.ReLU(inplace=True))
            layers.append(nn.Conv2d(output_dim, output_dim, kernel_size=3, stride=1, padding=int(padding)))
            layers.append(nn.ReLU(inplace=True))
            layers.append(nn.Conv2d(output_dim, output_dim, kernel_size=3, stride=1, padding=int(padding)))
            layers.append(nn.ReLU(inplace=True))
        else:
            layers.append(ReversibleSequence(input_dim, output_dim, reversible_depth=3))
        self.layers = nn.Sequential(*layers)
        self.layers.apply(init_weights)

    @_boundary
    def forward(self, x):
        mods = list(self.layers)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.AvgPool2d):
                x = Act(ops.AvgPool2.apply(x.t), x.c)
            elif isinstance(m, nn.Conv2d):
                relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
                if m.out_channels % 16 != 0:
                    raise NotImplementedError('U-Net filter counts must be multiples of 16 on the B200 path')
                x = Act(ops.ConvAffineAct.apply(x.t, m.weight, None, m.bias, relu, x.c, True), m.out_channels)
                if relu:
                    i += 1
            elif isinstance(m, nn.ReLU):
                raise RuntimeError('unexpected bare ReLU')
            else:
                x = m(x)
            i += 1
        return x


class UpConvBlock(nn.Module):
    def __init__(self, input_dim, output_dim, initializers, padding, bilinear=True, reversible=False):
        super(UpConvBlock, self).__init__()
        self.bilinear = bilinear
        if not self.bilinear:
            raise NotImplementedError('transposed-convolution upsampling is never used by the reference experiments')
        self.conv_block = DownConvBlock(input_dim, output_dim, initializers, padding, pool=False, reversible=reversible)

    def forward(self, x, bridge):
        plain = not isinstance(x, Act)
        x, bridge = ops.to_act(x), ops.to_act(bridge)
        assert 2 * x.t.shape[2] == bridge.t.shape[2]
        # up = interpolate(x, bilinear, x2, align_corners=False); out = cat([up, bridge]) in one buffer
        cat = Act(ops.Concat.apply(x.t, bridge.t, True, False, False), x.c + bridge.c)
        out = self.conv_block(cat)
        return ops.from_act(out) if plain else out


class Unet(nn.Module):
    """U-Net behind the reference API (models/unet.py:78-165)."""

    def __init__(self, input_channels, num_classes, num_filters, initializers=None, apply_last_layer=True,
                 padding=True, reversible=False, training=False, latent_dim=3, no_convs_fcomb=4, beta=1.0):
        super(Unet, self).__init__()
        self.input_channels = input_channels
        self.num_classes = num_classes
        self.num_filters = num_filters
        self.padding = padding
        self.activation_maps = []
        self.apply_last_layer = apply_last_layer
        self.contracting_path = nn.ModuleList()
        self.prediction = None
        for i in range(len(self.num_filters)):
            input = self.input_channels if i == 0 else output
            output = self.num_filters[i]
            pool = i != 0
            self.contracting_path.append(DownConvBlock(input, output, initializers, padding, pool=pool,
                                                       reversible=reversible))
        self.upsampling_path = nn.ModuleList()
        n = len(self.num_filters) - 2
        for i in range(n, -1, -1):
            input = output + self.num_filters[i]
            output = self.num_filters[i]
            self.upsampling_path.append(UpConvBlock(input, output, initializers, padding, reversible=reversible))
        if self.apply_last_layer:
            self.last_layer = nn.Conv2d(output, num_classes, kernel_size=1)

    def sample(self, testing=True):
        return self.prediction

    def _packer(self):
        ws = [m.weight for m in self.modules() if isinstance(m, nn.Conv2d) and m.out_channels % 16 == 0]
        pk = getattr(self, '_weight_packer', None)
        if pk is None or not pk.valid_for(ws[0]):
            pk = kern.WeightPacker(ws)
            object.__setattr__(self, '_weight_packer', pk)
        return pk

    def features(self, x):
        """encoder-decoder up to (not including) last_layer; returns the NHWC activation handle"""
        if not x.is_cuda:
            raise kern._lib.UnetZooLibError('UNet-Zoo B200 modules need CUDA tensors: there is no CPU fallback path')
        ws = [m.weight for m in self.modules() if isinstance(m, nn.Conv2d) and m.out_channels % 16 == 0]
        outer = kern._active_packer
        if outer is not None and outer.lookup(ws[0]) is not None:
            pk = outer                                   # an enclosing model (ProbabilisticUnet) already packed them
        else:
            pk = self._packer()
            pk.refresh()
        prev = kern.set_active_packer(pk)
        try:
            a = Act(kern.input_pack(x, None, cp=kern.pad16(self.input_channels)), self.input_channels)
            blocks = []
            for i, down in enumerate(self.contracting_path):
                a = down(a)
                if i != len(self.contracting_path) - 1:
                    blocks.append(a)
            for i, up in enumerate(self.upsampling_path):
                a = up(a, blocks[-i - 1])
            del blocks
        finally:
            kern.set_active_packer(prev)
        return a

    def forward(self, x, mask=None, training=True, val=False):
        a = self.features(x)
        if val:
            self.activation_maps.append(ops.from_act(a))
        if self.apply_last_layer:
            out = ops.SLayerNearest.apply(a.t, self.last_layer.weight, self.last_layer.bias, 1)
        else:
            out = ops.from_act(a)
        self.prediction = out
        return out

    def loss(self, mask):
        """nn.CrossEntropyLoss() (mean over all pixels) on mask.view(-1, 128, 128) -- reference unet.py:159-165,
        including the hard-coded 128 x 128."""
        target = mask.view(-1, 128, 128)
        total, _ = ops.ResidualCE.apply(target.unsqueeze(1).float(), self.prediction)
        return total / float(128 * 128)
