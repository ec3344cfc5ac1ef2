"""Drop-in for the reference's ``models/phiseg3D.py`` on B200.

Same classes, constructor keywords, method names, cached attributes and state_dict keys as the reference
(Conv3D :13-35, Conv3DSequence :38-58, ReversibleSequence :61-88, DownConvolutionalBlock :91-118, UpConvolutionalBlock
:121-153, SampleZBlock :156-190, Posterior :193-285, increase_resolution :288-301, Likelihood :304-400, PHISeg3D
:403-611).  Activations are bf16 NDHWC volumes; every 3x3x3 convolution (forward, input gradient, weight gradient) runs
on the same persistent tcgen05 kernel as the 2-D models with the z taps as one more factor of its K loop.

The reference file cannot complete a forward pass as shipped.  This module implements the documented
fixed specification, the one the oracle's patched reference (oracle/ref_run.py: build_reference_phiseg3d) also runs:
  (ii)  logits are nearest-upsampled to the full volume ``image_size[1:4]`` (the reference passes a 2-D size, :398);
  (iii) the posterior's conditioning mask is an index volume [B,1,D,H,W] one-hot encoded with ``num_classes`` labels
        (the reference hard-codes nlabels=2 at :253 while reserving num_classes channels at :221);
  (i)   the filter list must satisfy the reference's own channel arithmetic (num_filters[latent_levels-1] ==
        num_filters[-1]); nothing is changed here, inconsistent lists fail like in the reference.
Channel counts of the B200 path: every conv layer needs an output width that is a multiple of 16 (the reference's
BraTS widths 32/64/128 are); reversible blocks (depth 1 everywhere, :103,131,165,339,352) split it into two halves, so
they need multiples of 32.
"""
import torch
import torch.nn as nn

from b200 import kern, ops
from b200.ops import Act
import torchlayers
from torchlayers import _boundary, deferred_batch_counts
from models.phiseg import PHISeg, _Fork, _null_ctx, _CONCURRENT


class Conv3D(torchlayers.Conv2D):
    """nn.Conv3d(k in {3 (pad 1), 1}) + bias -> BatchNorm3d(eps 1e-3, momentum 0.01) -> ReLU (models/phiseg3D.py:13-35)"""
    _conv_cls = nn.Conv3d
    _norm_cls = nn.BatchNorm3d
    _granule = 16


class Conv3DSequence(nn.Module):
    """depth x Conv3D (models/phiseg3D.py:38-58)."""

    def __init__(self, input_dim, output_dim, kernel=3, depth=2, activation=torch.nn.ReLU, norm=torch.nn.BatchNorm3d,
                 norm_before_activation=True):
        super(Conv3DSequence, self).__init__()
        assert depth >= 1
        padding = 1 if kernel == 3 else 0
        layers = [Conv3D(input_dim, output_dim, kernel_size=kernel, padding=padding, activation=activation, norm=norm)]
        for i in range(depth - 1):
            layers.append(Conv3D(output_dim, output_dim, kernel_size=kernel, padding=padding, activation=activation,
                                 norm=norm))
        self.convolution = nn.Sequential(*layers)

    @_boundary
    def forward(self, x):
        for layer in self.convolution:
            x = layer(x)
        return x


class ReversibleSequence(torchlayers.ReversibleSequence):
    """models/phiseg3D.py:61-88: optional 1x1x1 Conv3D, then additive-coupling blocks of 3x3x3 Conv3D halves."""
    _conv_layer = Conv3D


def _pool(x):
    d, h, w = x.t.shape[1:4]
    if d % 2 or h % 2 or w % 2:
        raise ValueError('AvgPool3d on odd sizes is unreachable in the reference (skip-shape asserts)')
    return Act(ops.AvgPool2.apply(x.t), x.c)


def _up(x):
    return Act(ops.Upsample2x.apply(x.t, True), x.c)


class DownConvolutionalBlock(nn.Module):
    def __init__(self, input_dim, output_dim, initializers, depth=3, padding=True, pool=True, reversible=False):
        super(DownConvolutionalBlock, self).__init__()
        if depth < 1:
            raise ValueError
        layers = []
        if pool:
            layers.append(nn.AvgPool3d(kernel_size=2, stride=2, padding=0, ceil_mode=True))
        if reversible:
            layers.append(ReversibleSequence(input_dim, output_dim, reversible_depth=1))
        else:
            layers.append(Conv3D(input_dim, output_dim, kernel_size=3, stride=1, padding=int(padding)))
            if depth > 1:
                for i in range(depth - 1):
                    layers.append(Conv3D(output_dim, output_dim, kernel_size=3, stride=1, padding=int(padding)))
        self.layers = nn.Sequential(*layers)

    @_boundary
    def forward(self, x):
        for layer in self.layers:
            x = _pool(x) if isinstance(layer, nn.AvgPool3d) else layer(x)
        return x


class UpConvolutionalBlock(nn.Module):
    """trilinear x2 (align_corners=True) -> 2 x Conv3D -> cat([x, bridge])"""

    def __init__(self, input_dim, output_dim, initializers, padding, bilinear=True, reversible=False):
        super(UpConvolutionalBlock, self).__init__()
        self.bilinear = bilinear
        if self.bilinear:
            if reversible:
                self.upconv_layer = ReversibleSequence(input_dim, output_dim, reversible_depth=1)
            else:
                self.upconv_layer = nn.Sequential(
                    Conv3D(input_dim, output_dim, kernel_size=3, stride=1, padding=1),
                    Conv3D(output_dim, output_dim, kernel_size=3, stride=1, padding=1),
                )
        else:
            raise NotImplementedError

    def forward(self, x, bridge):
        plain = not isinstance(x, Act)
        x, bridge = ops.to_act(x), ops.to_act(bridge)
        if self.bilinear:
            x = _up(x)
            if isinstance(self.upconv_layer, nn.Sequential):
                for layer in self.upconv_layer:
                    x = layer(x)
            else:
                x = self.upconv_layer(x)
        assert x.t.shape[2] == bridge.t.shape[2]
        assert x.t.shape[1] == bridge.t.shape[1]
        out = Act(ops.Concat.apply(x.t, bridge.t, False, False, True), x.c + bridge.c)
        return ops.from_act(out) if plain else out


class SampleZBlock(nn.Module):
    """2 x Conv3D, then 1x1x1 heads: mu, sigma = softplus(.), z = mu + sigma * randn_like(sigma)."""

    def __init__(self, input_dim, z_dim0=2, depth=2, reversible=False):
        super(SampleZBlock, self).__init__()
        self.input_dim = input_dim
        layers = []
        if reversible:
            layers.append(ReversibleSequence(input_dim, input_dim, reversible_depth=1))
        else:
            for i in range(depth):
                layers.append(Conv3D(input_dim, input_dim, kernel_size=3, padding=1))
        self.conv = nn.Sequential(*layers)
        self.mu_conv = nn.Sequential(nn.Conv3d(input_dim, z_dim0, kernel_size=1))
        self.sigma_conv = nn.Sequential(nn.Conv3d(input_dim, z_dim0, kernel_size=1), nn.Softplus())

    def forward(self, pre_z):
        x = ops.to_act(pre_z)
        for layer in self.conv:
            x = layer(x)
        n = x.t.shape[0]
        zdim = self.mu_conv[0].out_channels
        # same call, shape, dtype and order as the reference (models/phiseg3D.py:188) => same RNG stream
        eps = torch.randn_like(torch.empty((n, zdim) + tuple(x.t.shape[1:4]), device=x.t.device, dtype=torch.float32),
                               dtype=torch.float32)
        return ops.LatentHead.apply(x.t, self.mu_conv[0].weight, self.mu_conv[0].bias, self.sigma_conv[0].weight,
                                    self.sigma_conv[0].bias, eps)


class Posterior(nn.Module):
    """Posterior network (prior when is_posterior=False); resolution levels = len(num_filters), latent levels from the
    constructor (models/phiseg3D.py:193-285)."""

    def __init__(self, input_channels, num_classes, num_filters, latent_levels, initializers=None, padding=True,
                 is_posterior=True, reversible=False):
        super(Posterior, self).__init__()
        self.input_channels = input_channels
        self.num_filters = num_filters
        self.num_classes = num_classes
        self.latent_levels = latent_levels
        self.resolution_levels = len(num_filters)
        self.lvl_diff = self.resolution_levels - self.latent_levels
        self.padding = padding
        self.activation_maps = []
        self.is_posterior = is_posterior
        if is_posterior:
            self.input_channels += num_classes

        self.contracting_path = nn.ModuleList()
        for i in range(self.resolution_levels):
            input = self.input_channels if i == 0 else output
            output = self.num_filters[i]
            pool = False if i == 0 else True
            self.contracting_path.append(DownConvolutionalBlock(input, output, initializers, depth=3, padding=padding,
                                                                pool=pool, reversible=reversible))
        self.upsampling_path = nn.ModuleList()
        for i in reversed(range(self.latent_levels)):
            input = 2
            output = self.num_filters[0] * 2
            self.upsampling_path.append(UpConvolutionalBlock(input, output, initializers, padding, reversible=reversible))
        self.sample_z_path = nn.ModuleList()
        for i in reversed(range(self.latent_levels)):
            input = 2 * self.num_filters[0] + self.num_filters[i + self.lvl_diff]
            if i == self.latent_levels - 1:
                input = self.num_filters[i + self.lvl_diff]
            self.sample_z_path.append(SampleZBlock(input, depth=2, reversible=reversible))

    def forward(self, patch, segm=None, training_prior=False, z_list=None):
        return self.latent(*self.contract(patch, segm), training_prior=training_prior, z_list=z_list)

    def contract(self, patch, segm=None):
        """Encoder half (models/phiseg3D.py:250-267): one-hot(mask) - 0.5 appended to the image channels straight in
        NDHWC bf16, then the DownConvolutionalBlocks.  No random draws."""
        if not patch.is_cuda:
            raise kern._lib.UnetZooLibError('UNet-Zoo B200 modules need CUDA tensors: there is no CPU fallback path')
        if patch.dim() != 5:
            raise ValueError('PHISeg3D expects [B,C,D,H,W] volumes')
        cp = kern.pad_channels(self.input_channels, 5)
        x = Act(kern.input_pack(patch, segm if segm is not None else None, nlabels=self.num_classes, cp=cp),
                self.input_channels)
        blocks = []
        for i, down in enumerate(self.contracting_path):
            x = down(x)
            if i != len(self.contracting_path) - 1:
                blocks.append(x)
        return x, blocks

    def latent(self, x, blocks, training_prior=False, z_list=None):
        z = [None] * self.latent_levels
        sigma = [None] * self.latent_levels
        mu = [None] * self.latent_levels
        pre_conv = x
        for i, sample_z in enumerate(self.sample_z_path):
            if i != 0:
                pre_conv = self.upsampling_path[i - 1](ops.to_act(z[-i]), blocks[-i])
            mu[-i - 1], sigma[-i - 1], z[-i - 1] = self.sample_z_path[i](pre_conv)
            if training_prior:
                z[-i - 1] = z_list[-i - 1]      # own draw discarded AFTER it was made (RNG order of the reference)
        del blocks
        return z, mu, sigma


class _UpsampleConvStack(nn.Sequential):
    """nn.Sequential of [nn.Upsample(trilinear), Conv3DSequence] * n whose forward runs the B200 kernels."""

    @_boundary
    def forward(self, x):
        for m in self:
            x = _up(x) if isinstance(m, nn.Upsample) else m(x)
        return x


def increase_resolution(times, input_dim, output_dim):
    """models/phiseg3D.py:288-301"""
    module_list = []
    for i in range(times):
        module_list.append(nn.Upsample(mode='trilinear', scale_factor=2, align_corners=True))
        if i != 0:
            input_dim = output_dim
        module_list.append(Conv3DSequence(input_dim=input_dim, output_dim=output_dim, depth=1))
    return _UpsampleConvStack(*module_list)


class Likelihood(nn.Module):
    def __init__(self, input_channels, num_classes, num_filters, latent_levels=5, image_size=(128, 128, 1),
                 reversible=False, initializers=None, apply_last_layer=True, padding=True):
        super(Likelihood, self).__init__()
        self.input_channels = input_channels
        self.num_classes = num_classes
        self.num_filters = num_filters
        self.latent_levels = latent_levels
        self.resolution_levels = len(num_filters)
        self.lvl_diff = self.resolution_levels - latent_levels
        self.image_size = image_size
        self.reversible = reversible
        self.padding = padding
        self.activation_maps = []
        self.apply_last_layer = apply_last_layer

        self.likelihood_ups_path = nn.ModuleList()
        self.likelihood_post_ups_path = nn.ModuleList()
        for i in reversed(range(self.latent_levels)):
            input = self.num_filters[i]
            if reversible:
                self.likelihood_ups_path.append(ReversibleSequence(input_dim=2, output_dim=input, reversible_depth=1))
            else:
                self.likelihood_ups_path.append(Conv3DSequence(input_dim=2, output_dim=input, depth=2))
            self.likelihood_post_ups_path.append(increase_resolution(times=self.lvl_diff, input_dim=input,
                                                                     output_dim=input))
        self.likelihood_post_c_path = nn.ModuleList()
        for i in range(latent_levels - 1):
            input = self.num_filters[i] + self.num_filters[i + 1 + self.lvl_diff]
            output = self.num_filters[i + self.lvl_diff]
            if reversible:
                self.likelihood_post_c_path.append(ReversibleSequence(input_dim=input, output_dim=output,
                                                                      reversible_depth=1))
            else:
                self.likelihood_post_c_path.append(Conv3DSequence(input_dim=input, output_dim=output, depth=2))
        self.s_layer = nn.ModuleList()
        output = self.num_classes
        for i in reversed(range(self.latent_levels)):
            input = self.num_filters[i + self.lvl_diff]
            self.s_layer.append(Conv3DSequence(input_dim=input, output_dim=output, depth=1, kernel=1,
                                               activation=torch.nn.Identity, norm=torch.nn.Identity))

    def forward(self, z):
        """z: list of latent volumes [B,2,r,r,r] fp32 (index = latent level) -> list of full-resolution logits."""
        s = [None] * self.latent_levels
        post_z = [None] * self.latent_levels
        post_c = [None] * self.latent_levels
        for i in range(self.latent_levels):
            assert z[-i - 1].shape[1] == 2
            assert z[-i - 1].shape[2] == self.image_size[1] * 2 ** (-self.resolution_levels + 1 + i)
            x = self.likelihood_ups_path[i](ops.to_act(z[-i - 1]))
            x = self.likelihood_post_ups_path[i](x)
            assert x.t.shape[1] == self.image_size[1] * 2 ** (-self.latent_levels + i + 1)
            assert x.c == self.num_filters[-i - 1 - self.lvl_diff], '{} != {}'.format(x.c, self.num_filters[-i - 1])
            post_z[-i - 1] = x
        post_c[self.latent_levels - 1] = post_z[self.latent_levels - 1]
        for i in reversed(range(self.latent_levels - 1)):
            below = post_c[i + 1]
            assert post_z[i].t.shape[1] == 2 * below.t.shape[1] and post_z[i].t.shape[2] == 2 * below.t.shape[2]
            # trilinear x2 of the level below written straight into the concat buffer
            concat = Act(ops.Concat.apply(post_z[i].t, below.t, False, True, True), post_z[i].c + below.c)
            post_c[i] = self.likelihood_post_c_path[i](concat)
        for i, block in enumerate(self.s_layer):
            feat = post_c[-i - 1]
            conv = block.convolution[0].convolution[0]
            factor = self.image_size[1] // feat.t.shape[1]
            assert all(factor * feat.t.shape[1 + k] == self.image_size[1 + k] for k in range(3)), \
                'logits are upsampled by one integer factor to image_size[1:4] (fixed specification (ii))'
            s[-i - 1] = ops.SLayerNearest.apply(feat.t, conv.weight, conv.bias, factor)
        return s


class PHISeg3D(PHISeg):
    """PHiSeg on volumes behind the reference's module API (models/phiseg3D.py:403-611).  Sampling, accumulate_output,
    KL, residual cross-entropy and the aliased loss bookkeeping are dimension independent and inherited from the 2-D
    drop-in (the reference's two files are line-for-line clones there: phiseg.py:381-537 vs phiseg3D.py:455-611)."""

    def __init__(self, input_channels, num_classes, num_filters, latent_levels=5, initializers=None, no_convs_fcomb=4,
                 beta=10.0, image_size=(128, 128, 1), reversible=False, apply_last_layer=True,
                 exponential_weighting=True, padding=True):
        nn.Module.__init__(self)
        self.input_channels = input_channels
        self.num_classes = num_classes
        self.num_filters = num_filters
        self.latent_levels = latent_levels
        self.image_size = image_size
        self.loss_tot = 0
        self.loss_dict = {}
        self.kl_divergence_loss_weight = 1.0
        self.beta = 1.0
        self.padding = padding
        self.activation_maps = []
        self.apply_last_layer = apply_last_layer
        self.exponential_weighting = exponential_weighting
        self.exponential_weight = 4
        self.residual_multinoulli_loss_weight = 1.0
        self.kl_divergence_loss = 0
        self.reconstruction_loss = 0

        self.posterior = Posterior(input_channels, num_classes, num_filters, latent_levels=latent_levels,
                                   initializers=None, padding=True, reversible=reversible)
        self.likelihood = Likelihood(input_channels, num_classes, num_filters, latent_levels=latent_levels,
                                     initializers=None, apply_last_layer=True, padding=True,
                                     image_size=self.image_size, reversible=reversible)
        self.prior = Posterior(input_channels, num_classes, num_filters, latent_levels=latent_levels,
                               initializers=None, padding=True, is_posterior=False, reversible=reversible)
        self.s_out_list = [None] * self.latent_levels
        self.s_out_list_with_softmax = [None] * self.latent_levels

    def _packer(self):
        convs = [m for m in self.modules() if isinstance(m, nn.Conv3d) and m.out_channels % 16 == 0 and m.weight.is_cuda]
        pk = getattr(self, '_weight_packer', None)
        if pk is None or not pk.valid_for(convs[0].weight):
            pk = kern.WeightPacker([m.weight for m in convs])
            object.__setattr__(self, '_weight_packer', pk)
        return pk

    def _forward(self, patch, mask, training=True, replicate=1, lowres_logits=False):
        # volumes fill the GPU on their own: one stream, the reference's order
        if replicate != 1 or lowres_logits:
            raise NotImplementedError('replicate / lowres_logits are evaluation shortcuts of the 2-D model')
        post_x, post_blocks = self.posterior.contract(patch, mask)
        self.posterior_latent_space, self.posterior_mu, self.posterior_sigma = self.posterior.latent(post_x, post_blocks)
        del post_x, post_blocks
        prior_x, prior_blocks = self.prior.contract(patch)
        if training:
            self.prior_latent_space, self.prior_mu, self.prior_sigma = self.prior.latent(
                prior_x, prior_blocks, training_prior=True, z_list=self.posterior_latent_space)
            self.s_out_list = self.likelihood(self.posterior_latent_space)
        else:
            self.prior_latent_space, self.prior_mu, self.prior_sigma = self.prior.latent(prior_x, prior_blocks)
            self.s_out_list = self.likelihood(self.prior_latent_space)
        return self.s_out_list
