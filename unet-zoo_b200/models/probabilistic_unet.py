"""Drop-in for the reference's ``models/probabilistic_unet.py`` (Encoder :20-70, AxisAlignedConvGaussian :73-130,
Fcomb :133-199, ProbabilisticUnet :202-370) on B200: same classes, constructor keywords, methods
(forward / sample / reconstruct / accumulate_output / kl_divergence / elbo / loss) and state_dict keys.

Kernel mapping: encoders = tensor-core conv + fused BatchNorm/ReLU stacks; spatial mean -> uz_global_mean; the
2*latent_dim Gaussian head and the class logits are the small 1x1-conv kernel (uz_slayer); fcomb = 1x1 tensor-core
convs on [features | tiled z]; KL = uz_kl (same sigma1*sigma0 quirk as PHiSeg, probabilistic_unet.py:292-308);
reconstruction loss = uz_residual_ce.  The latent distributions stay ``torch.distributions`` objects on [B, L] tensors
because they are part of the API surface (``prior_latent_space.rsample()``, ``.base_dist.loc`` ...).
"""
import numpy as np
import torch
import torch.nn as nn
from torch.distributions import Independent, Normal

import utils
from b200 import kern, ops
from b200.ops import Act
from models.unet import Unet
from torchlayers import Conv2D, Conv2DSequence, ReversibleSequence, _boundary, deferred_batch_counts
from utils import init_weights, init_weights_orthogonal_normal, l2_regularisation

device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')


class Encoder(nn.Module):
    """len(num_filters) blocks of no_convs_per_block Conv2D with AvgPool2d between blocks."""

    def __init__(self, input_channels, num_filters, no_convs_per_block, num_classes=2, initializers=None, padding=True,
                 posterior=False, reversible=False):
        super(Encoder, self).__init__()
        self.contracting_path = nn.ModuleList()
        self.input_channels = input_channels
        self.num_filters = num_filters
        if posterior:
            self.input_channels += num_classes
        layers = []
        for i in range(len(self.num_filters)):
            input_dim = self.input_channels if i == 0 else output_dim
            output_dim = num_filters[i]
            if i != 0:
                layers.append(nn.AvgPool2d(kernel_size=2, stride=2, padding=0, ceil_mode=True))
            if reversible:
                layers.append(ReversibleSequence(input_dim, output_dim, kernel=3, reversible_depth=no_convs_per_block - 1))
            else:
                layers.append(Conv2DSequence(input_dim, output_dim, kernel=3, depth=no_convs_per_block))
        self.layers = nn.Sequential(*layers)
        self.layers.apply(init_weights)

    @_boundary
    def forward(self, x):
        for layer in self.layers:
            if isinstance(layer, nn.AvgPool2d):
                x = Act(ops.AvgPool2.apply(x.t), x.c)
            else:
                x = layer(x)
        return x


class AxisAlignedConvGaussian(nn.Module):
    """Encoder -> spatial mean -> 1x1 conv to (mu, log sigma) -> Independent(Normal(mu, exp(log sigma)))."""

    def __init__(self, input_channels, num_filters, no_convs_per_block, latent_dim, initializers, posterior=False):
        super(AxisAlignedConvGaussian, self).__init__()
        self.input_channels = input_channels
        self.channel_axis = 1
        self.num_filters = num_filters
        self.no_convs_per_block = no_convs_per_block
        self.latent_dim = latent_dim
        self.posterior = posterior
        self.name = 'Posterior' if self.posterior else 'Prior'
        self.encoder = Encoder(self.input_channels, self.num_filters, self.no_convs_per_block,
                               initializers=initializers, posterior=self.posterior)
        self.conv_layer = nn.Conv2d(num_filters[-1], 2 * self.latent_dim, kernel_size=1, stride=1)
        self.sum_input = 0
        nn.init.kaiming_normal_(self.conv_layer.weight, mode='fan_in', nonlinearity='relu')
        nn.init.normal_(self.conv_layer.bias)

    def forward(self, input, segm=None):
        if not input.is_cuda:
            raise kern._lib.UnetZooLibError('UNet-Zoo B200 modules need CUDA tensors: there is no CPU fallback path')
        cin = self.encoder.input_channels
        x = Act(kern.input_pack(input, segm, nlabels=2, cp=kern.pad16(cin)), cin)     # one-hot(mask) - 0.5 appended
        enc = self.encoder(x)
        pooled = ops.GlobalMean.apply(enc.t)                                           # [B,1,1,C]
        mu_log_sigma = ops.SLayerNearest.apply(pooled, self.conv_layer.weight, self.conv_layer.bias, 1)   # [B,2L,1,1]
        mu_log_sigma = mu_log_sigma[:, :, 0, 0]
        mu = mu_log_sigma[:, :self.latent_dim]
        log_sigma = mu_log_sigma[:, self.latent_dim:]
        # validate_args=False: the default argument validation synchronises the host (not CUDA-graph capturable)
        return Independent(Normal(loc=mu, scale=torch.exp(log_sigma), validate_args=False), 1, validate_args=False)


class Fcomb(nn.Module):
    """1x1 convs on cat(feature_map, tiled z) (probabilistic_unet.py:133-199)."""

    def __init__(self, num_filters, latent_dim, num_output_channels, num_classes, no_convs_fcomb, initializers,
                 use_tile=True):
        super(Fcomb, self).__init__()
        self.num_channels = num_output_channels
        self.num_classes = num_classes
        self.channel_axis = 1
        self.spatial_axes = [2, 3]
        self.num_filters = num_filters
        self.latent_dim = latent_dim
        self.use_tile = use_tile
        self.no_convs_fcomb = no_convs_fcomb
        self.name = 'Fcomb'
        if self.use_tile:
            layers = [Conv2D(self.num_filters[0] + self.latent_dim, self.num_filters[0], kernel_size=1)]
            for _ in range(no_convs_fcomb - 2):
                layers.append(Conv2D(self.num_filters[0], self.num_filters[0], kernel_size=1))
            self.layers = nn.Sequential(*layers)
            self.last_layer = nn.Conv2d(self.num_filters[0], self.num_classes, kernel_size=1)
            if initializers['w'] == 'orthogonal':
                self.layers.apply(init_weights_orthogonal_normal)
                self.last_layer.apply(init_weights_orthogonal_normal)
            else:
                self.layers.apply(init_weights)
                self.last_layer.apply(init_weights)

    def forward(self, feature_map, z):
        """feature_map: NHWC Act or NCHW tensor [B,C,H,W]; z: [B, latent_dim] -> class logits fp32 NCHW."""
        if self.use_tile:
            fm = ops.to_act(feature_map)
            n, h, w, _ = fm.t.shape
            zt = ops.to_act(z[:, :, None, None].expand(n, z.shape[1], h, w))           # broadcast over H x W
            cat = Act(ops.Concat.apply(fm.t, zt.t, False, False, True), fm.c + zt.c)
            # fm.c is a multiple of 16, so z's channels sit right behind the features in the padded concat buffer and
            # the first 1x1 conv sees logical channels [features | z] like torch.cat((feature_map, z), dim=1)
            if fm.c % 16 != 0:
                raise NotImplementedError('feature maps with a channel count that is not a multiple of 16')
            x = cat
            for layer in self.layers:
                x = layer(x)
            return ops.SLayerNearest.apply(x.t, self.last_layer.weight, self.last_layer.bias, 1)


class ProbabilisticUnet(nn.Module):
    """Probabilistic U-Net behind the reference API (models/probabilistic_unet.py:202-370)."""

    def __init__(self, input_channels=1, num_classes=1, num_filters=None, latent_levels=1, latent_dim=2,
                 initializers=None, no_convs_fcomb=4, image_size=(1, 128, 128), beta=10.0, reversible=False):
        super(ProbabilisticUnet, self).__init__()
        self.input_channels = input_channels
        self.num_classes = num_classes
        self.num_filters = num_filters
        self.latent_dim = latent_dim
        self.no_convs_per_block = 3
        self.no_convs_fcomb = no_convs_fcomb
        self.initializers = {'w': 'he_normal', 'b': 'normal'}
        self.z_prior_sample = 0
        self.unet = Unet(self.input_channels, self.num_classes, self.num_filters, initializers=self.initializers,
                         apply_last_layer=False, padding=True, reversible=reversible).to(device)
        self.prior = AxisAlignedConvGaussian(self.input_channels, self.num_filters, self.no_convs_per_block,
                                             self.latent_dim, initializers=self.initializers).to(device)
        self.posterior = AxisAlignedConvGaussian(self.input_channels, self.num_filters, self.no_convs_per_block,
                                                 self.latent_dim, initializers=self.initializers, posterior=True
                                                 ).to(device)
        self.fcomb = Fcomb(self.num_filters, self.latent_dim, self.input_channels, self.num_classes,
                           self.no_convs_fcomb, initializers={'w': 'orthogonal', 'b': 'normal'}, use_tile=True
                           ).to(device)
        self.last_conv = Conv2D(32, num_classes, kernel_size=1, activation=torch.nn.Identity, norm=torch.nn.Identity)

    # ---------------------------------------------------------------- plumbing shared by the sub-networks
    def _packer(self):
        ws = [m.weight for m in self.modules()
              if isinstance(m, nn.Conv2d) and m.out_channels % 16 == 0 and m.weight.is_cuda]
        pk = getattr(self, '_weight_packer', None)
        if pk is None or not pk.valid_for(ws[0]):
            pk = kern.WeightPacker(ws)
            object.__setattr__(self, '_weight_packer', pk)
        return pk

    @property
    def unet_features(self):
        """NCHW fp32 view of the cached U-Net feature map (the reference caches the tensor itself, :254)"""
        return ops.from_act(self._unet_features)

    def forward(self, patch, segm=None, training=True):
        if not patch.is_cuda:
            raise kern._lib.UnetZooLibError('UNet-Zoo B200 modules need CUDA tensors: there is no CPU fallback path')
        pk = self._packer()
        pk.refresh()
        kern.zero_arena.reset(patch.device)
        prev = kern.set_active_packer(pk)
        try:
            with deferred_batch_counts():
                if segm is not None:
                    self.posterior_latent_space = self.posterior.forward(patch, segm)
                self.prior_latent_space = self.prior.forward(patch)
                object.__setattr__(self, '_unet_features', self.unet.features(patch))
            conv = self.last_conv.convolution[0]
            return ops.SLayerNearest.apply(self._unet_features.t, conv.weight, conv.bias, 1)
        finally:
            kern.set_active_packer(prev)

    def _with_packer(self, fn, refresh):
        """sample() / reconstruct() / elbo() run 1x1 convs outside forward(): they activate the model's packed weights
        themselves and restore whatever was active before (nothing stays installed process-wide).  ``refresh`` re-packs
        first -- needed when the parameters may have changed since forward() (sample() after an optimizer step)."""
        pk = self._packer()
        if refresh:
            pk.refresh()
        prev = kern.set_active_packer(pk)
        try:
            return fn()
        finally:
            kern.set_active_packer(prev)

    def sample(self, testing=False):
        if testing is False:
            z_prior = self.prior_latent_space.rsample()
        else:
            z_prior = self.prior_latent_space.sample()
        self.z_prior_sample = z_prior
        return self._with_packer(lambda: self.fcomb.forward(self._unet_features, z_prior), refresh=True)

    def reconstruct(self, use_posterior_mean=False, calculate_posterior=False, z_posterior=None):
        if use_posterior_mean:
            z_posterior = self.posterior_latent_space.loc
        elif calculate_posterior:
            z_posterior = self.posterior_latent_space.rsample()
        return self._with_packer(lambda: self.fcomb.forward(self._unet_features, z_posterior), refresh=False)

    def accumulate_output(self, output_list, use_softmax=False):
        s_accum = output_list
        if use_softmax:
            out = torch.empty_like(s_accum)
            return kern.accumulate_output([s_accum.contiguous()], True, out)
        return s_accum

    def KL_two_gauss_with_diag_cov(self, mu0, sigma0, mu1, sigma1):
        return ops.KLLevel.apply(mu0, sigma0, mu1, sigma1, 1.0)

    def kl_divergence(self, analytic=True, calculate_posterior=False, z_posterior=None):
        mu0 = self.posterior_latent_space.mean
        sigma0 = self.posterior_latent_space.stddev
        mu1 = self.prior_latent_space.mean
        sigma1 = self.prior_latent_space.stddev
        return self.KL_two_gauss_with_diag_cov(mu0, sigma0, mu1, sigma1)

    def multinoulli_loss(self, reconstruction, target):
        total, _ = ops.ResidualCE.apply(target.float(), reconstruction)
        return total

    def elbo(self, segm, analytic_kl=False, reconstruct_posterior_mean=False):
        z_posterior = self.posterior_latent_space.rsample()
        self.kl_divergence_loss = torch.mean(
            self.kl_divergence(analytic=analytic_kl, calculate_posterior=False, z_posterior=z_posterior))
        with deferred_batch_counts():
            self.reconstruction = self.reconstruct(use_posterior_mean=reconstruct_posterior_mean,
                                                   calculate_posterior=False, z_posterior=z_posterior)
        reconstruction_loss = self.multinoulli_loss(reconstruction=self.reconstruction, target=segm)
        self.reconstruction_loss = torch.sum(reconstruction_loss)
        self.mean_reconstruction_loss = torch.mean(reconstruction_loss)
        return -(self.reconstruction_loss + 1.0 * self.kl_divergence_loss)

    def loss(self, mask):
        elbo = self.elbo(mask)
        reg_loss = l2_regularisation(self.posterior) + l2_regularisation(self.prior) + l2_regularisation(
            self.fcomb.layers)
        loss = -elbo + 1e-5 * reg_loss
        return loss
