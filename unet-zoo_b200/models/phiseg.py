"""Drop-in for the reference's ``models/phiseg.py`` on B200.

Same public classes, constructor keywords, method names, cached attributes and state_dict keys as the reference
(DownConvolutionalBlock :14-39, UpConvolutionalBlock :42-73, SampleZBlock :76-106, Posterior :109-206,
increase_resolution :209-221, Likelihood :224-323, PHISeg :326-537), so ``train_model.py`` / ``test_model.py`` and
the experiment files run unmodified.  Underneath, activations stay bf16 NHWC between layers and every op is a kernel
of libunetzoo_b200.so; mu / sigma / z / logits / losses are fp32 NCHW tensors exactly where the reference exposes them.

Quirks reproduced on purpose: Q1 loss aliasing, Q2 in-place accumulate_output, Q3 sigma1*sigma0 in
the KL, Q4 RNG draw order (torch.randn_like is called with the reference's shapes in the reference's order, so a
seeded run consumes the same Philox stream), Q5 ignored constructor arguments.
"""
import os

import torch
import torch.nn as nn

from b200 import kern, ops
from b200.ops import Act
from torchlayers import Conv2D, Conv2DSequence, ReversibleSequence, _boundary, deferred_batch_counts


# Independent sub-graphs (prior vs posterior encoder, the likelihood's per-level branches) are issued on a second CUDA
# stream: most of their layers are too small to fill 148 SMs on their own.  UNETZOO_CONCURRENCY=0 disables it.
_CONCURRENT = os.environ.get('UNETZOO_CONCURRENCY', '1') != '0'
_EARLY_LOGITS = os.environ.get('UNETZOO_EARLY_LOGITS', '1') != '0'
_PRIOR_OVERLAP = os.environ.get('UNETZOO_PRIOR_OVERLAP', '1') != '0'    # prior latent path next to the likelihood
_LIKELIHOOD_STREAMS = max(1, int(os.environ.get('UNETZOO_LIKELIHOOD_STREAMS', '4')))   # side streams of Likelihood.forward
_side_streams = {}
# Transparent CUDA-graph capture of the training step for the UNMODIFIED caller (train_model.py:100-134 issues
# net.forward(training=True) -> net.loss(mask) -> optimizer.zero_grad() -> loss.backward() -> optimizer.step() eagerly,
# ~900 launches from Python: host-bound at ~400 images/s).  With UNETZOO_TRANSPARENT_GRAPH=1 (or net.transparent_graph =
# True; launch.py --graph sets it) the third identically-shaped training forward captures forward + ELBO + backward as ONE
# graph (exactly what b200.train.TrainStep captures, minus the optimizer); net.forward() replays it, net.loss() hands out
# a tensor whose backward() just returns the gradients the replay left in static buffers, and the caller's own optimizer
# keeps running eagerly on them.  Needs optimizer.zero_grad() to set gradients to None (the torch >= 2.0 default).
_TRANSPARENT = os.environ.get('UNETZOO_TRANSPARENT_GRAPH', '0') == '1'
_TRANSPARENT_WARMUP = 2
_TG_ATTRS = ('posterior_latent_space', 'posterior_mu', 'posterior_sigma', 'prior_latent_space', 'prior_mu', 'prior_sigma',
             's_out_list', 'loss_dict', 'loss_tot', 'kl_divergence_loss', 'reconstruction_loss')


class _GraphLoss(torch.autograd.Function):
    """The loss tensor handed to the caller in transparent-graph mode: the captured step already computed the parameter
    gradients into static buffers, backward() returns them."""

    @staticmethod
    def forward(ctx, loss_static, state, *params):
        ctx.state = state
        return loss_static.clone()

    @staticmethod
    def backward(ctx, g):
        st = ctx.state
        out, live = [], []
        for p, gr in zip(st['params'], st['grads']):
            if gr is None:
                out.append(None)
                continue
            if p.grad is not None and p.grad.data_ptr() == gr.data_ptr():
                raise RuntimeError('transparent-graph mode: .grad still aliases the captured gradient buffer -- call '
                                   'optimizer.zero_grad(set_to_none=True) (the default) before loss.backward()')
            v = gr.view_as(gr)                 # a fresh handle on the static buffer: AccumulateGrad adopts it without a copy
            out.append(v)
            live.append(v)
        # upstream gradient (1 for loss.backward()): applied on the device, in place, without reading it on the host --
        # a host read here would serialise the caller's remaining per-step work behind the whole replayed step
        torch._foreach_mul_(live, g)
        return (None, None) + tuple(out)


def _use_streams(t, hw=None):
    """concurrency pays while single layers cannot fill the GPU: training-sized batches, not 100-sample evaluation"""
    if not _CONCURRENT:
        return False
    if hw is None:
        hw = t.shape[-1] * t.shape[-2]
    return t.shape[0] * hw <= (1 << 19)


class _null_ctx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class _Fork:
    """fork the current stream into per-device side streams and join them back (CUDA-graph capturable)"""

    def __init__(self, device, n=1, tag=''):
        self.main = torch.cuda.current_stream(device)
        dev = device.index if device.index is not None else torch.cuda.current_device()
        self.streams = []
        for k in range(n):
            key = (dev, self.main.cuda_stream, tag, k)
            if key not in _side_streams:
                _side_streams[key] = torch.cuda.Stream(device=device, priority=self.main.priority)
            self.streams.append(_side_streams[key])
            self.streams[-1].wait_stream(self.main)
        self.stream = self.streams[0]

    def side(self, k=0):
        return torch.cuda.stream(self.streams[k % len(self.streams)])

    def hand_over(self, *tensors):
        for t in tensors:
            t.record_stream(self.main)

    def join(self):
        for st in self.streams:
            self.main.wait_stream(st)


class DownConvolutionalBlock(nn.Module):
    def __init__(self, input_dim, output_dim, initializers, depth=3, padding=True, pool=True, reversible=False):
        super(DownConvolutionalBlock, self).__init__()
        if depth < 1:
            raise ValueError
        layers = []
        if pool:
            layers.append(nn.AvgPool2d(kernel_size=2, stride=2, padding=0, ceil_mode=True))
        if reversible:
            layers.append(ReversibleSequence(input_dim, output_dim, reversible_depth=3))
        else:
            layers.append(Conv2D(input_dim, output_dim, kernel_size=3, stride=1, padding=int(padding)))
            if depth > 1:
                for i in range(depth - 1):
                    layers.append(Conv2D(output_dim, output_dim, kernel_size=3, stride=1, padding=int(padding)))
        self.layers = nn.Sequential(*layers)

    @_boundary
    def forward(self, x):
        for layer in self.layers:
            if isinstance(layer, nn.AvgPool2d):
                if x.t.shape[1] % 2 or x.t.shape[2] % 2:
                    raise ValueError('AvgPool2d on odd sizes is unreachable in the reference (skip-shape asserts)')
                x = Act(ops.AvgPool2.apply(x.t), x.c)
            else:
                x = layer(x)
        return x


class UpConvolutionalBlock(nn.Module):
    """bilinear x2 (align_corners=True) -> 2 x Conv2D -> cat([x, bridge])"""

    def __init__(self, input_dim, output_dim, initializers, padding, bilinear=True, reversible=False):
        super(UpConvolutionalBlock, self).__init__()
        self.bilinear = bilinear
        if self.bilinear:
            if reversible:
                self.upconv_layer = ReversibleSequence(input_dim, output_dim, reversible_depth=2)
            else:
                self.upconv_layer = nn.Sequential(
                    Conv2D(input_dim, output_dim, kernel_size=3, stride=1, padding=1),
                    Conv2D(output_dim, output_dim, kernel_size=3, stride=1, padding=1),
                )
        else:
            raise NotImplementedError

    def forward(self, x, bridge):
        plain = not isinstance(x, Act)
        x, bridge = ops.to_act(x), ops.to_act(bridge)
        if self.bilinear:
            x = Act(ops.Upsample2x.apply(x.t, True), x.c)
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
    """2 x Conv2D, then 1x1 heads: mu, sigma = softplus(.), z = mu + sigma * randn_like(sigma)."""

    def __init__(self, input_dim, z_dim0=2, depth=2, reversible=False):
        super(SampleZBlock, self).__init__()
        self.input_dim = input_dim
        layers = []
        if reversible:
            layers.append(ReversibleSequence(input_dim, input_dim, reversible_depth=3))
        else:
            for i in range(depth):
                layers.append(Conv2D(input_dim, input_dim, kernel_size=3, padding=1))
        self.conv = nn.Sequential(*layers)
        self.mu_conv = nn.Sequential(nn.Conv2d(input_dim, z_dim0, kernel_size=1))
        self.sigma_conv = nn.Sequential(nn.Conv2d(input_dim, z_dim0, kernel_size=1), nn.Softplus())

    def forward(self, pre_z):
        x = ops.to_act(pre_z)
        for layer in self.conv:
            x = layer(x)
        n, h, w, _ = x.t.shape
        zdim = self.mu_conv[0].out_channels
        # same call, shape, dtype and order as the reference (models/phiseg.py:104) => same RNG stream
        eps = torch.randn_like(torch.empty((n, zdim, h, w), device=x.t.device, dtype=torch.float32), dtype=torch.float32)
        mu, sigma, z = ops.LatentHead.apply(x.t, self.mu_conv[0].weight, self.mu_conv[0].bias,
                                            self.sigma_conv[0].weight, self.sigma_conv[0].bias, eps)
        return mu, sigma, z


class Posterior(nn.Module):
    """Posterior network (prior when is_posterior=False): 7 resolution levels, 5 latent levels (hard-coded like the
    reference, models/phiseg.py:131-132)."""

    def __init__(self, input_channels, num_classes, num_filters, initializers=None, padding=True, is_posterior=True,
                 reversible=False):
        super(Posterior, self).__init__()
        self.input_channels = input_channels
        self.num_filters = num_filters
        self.latent_levels = 5
        self.resolution_levels = 7
        self.lvl_diff = self.resolution_levels - self.latent_levels
        self.padding = padding
        self.activation_maps = []
        self.is_posterior = is_posterior
        self.image_channels = input_channels
        if is_posterior:
            self.input_channels += 2          # one-hot mask, hard-coded two labels (models/phiseg.py:140,179)

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
        """Encoder half (models/phiseg.py:175-194): input packing + the 7 DownConvolutionalBlocks.  No random draws, so
        PHISeg.forward may run the prior's encoder concurrently with the posterior's on a second stream."""
        if not patch.is_cuda:
            raise kern._lib.UnetZooLibError('UNet-Zoo B200 modules need CUDA tensors: there is no CPU fallback path')
        cp = kern.pad16(self.input_channels)
        # one-hot(mask) - 0.5 concatenated after the image channels, produced directly in NHWC bf16 on the device
        x = Act(kern.input_pack(patch, segm if segm is not None else None, nlabels=2, cp=cp), self.input_channels)
        blocks = []
        for i, down in enumerate(self.contracting_path):
            x = down(x)
            if i != len(self.contracting_path) - 1:
                blocks.append(x)
        return x, blocks

    def latent(self, x, blocks, training_prior=False, z_list=None):
        """Latent half (models/phiseg.py:196-206): per level [upsample z, 2 convs, cat skip] -> SampleZBlock."""
        z = [None] * self.latent_levels
        sigma = [None] * self.latent_levels
        mu = [None] * self.latent_levels
        pre_conv = x
        for i, sample_z in enumerate(self.sample_z_path):
            if i != 0:
                pre_conv = self.upsampling_path[i - 1](ops.to_act(z[-i]), blocks[-i])
            mu[-i - 1], sigma[-i - 1], z[-i - 1] = self.sample_z_path[i](pre_conv)
            if training_prior:
                z[-i - 1] = z_list[-i - 1]      # own draw discarded AFTER it was made (quirk Q4)
        del blocks
        return z, mu, sigma


class _UpsampleConvStack(nn.Sequential):
    """nn.Sequential of [nn.Upsample, Conv2DSequence] * n whose forward runs the B200 kernels."""

    @_boundary
    def forward(self, x):
        for m in self:
            if isinstance(m, nn.Upsample):
                x = Act(ops.Upsample2x.apply(x.t, True), x.c)
            else:
                x = m(x)
        return x


def increase_resolution(times, input_dim, output_dim):
    """Increase the resolution by n times for the beginning of the likelihood path (models/phiseg.py:209-221)."""
    module_list = []
    for i in range(times):
        module_list.append(nn.Upsample(mode='bilinear', scale_factor=2, align_corners=True))
        if i != 0:
            input_dim = output_dim
        module_list.append(Conv2DSequence(input_dim=input_dim, output_dim=output_dim, depth=1))
    return _UpsampleConvStack(*module_list)


class Likelihood(nn.Module):
    def __init__(self, input_channels, num_classes, num_filters, latent_levels=5, resolution_levels=7,
                 image_size=(128, 128, 1), reversible=False, initializers=None, apply_last_layer=True, padding=True):
        super(Likelihood, self).__init__()
        self.input_channels = input_channels
        self.num_classes = num_classes
        self.num_filters = num_filters
        self.latent_levels = latent_levels
        self.resolution_levels = resolution_levels
        self.lvl_diff = resolution_levels - latent_levels
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
                self.likelihood_ups_path.append(ReversibleSequence(input_dim=2, output_dim=input, reversible_depth=2))
            else:
                self.likelihood_ups_path.append(Conv2DSequence(input_dim=2, output_dim=input, depth=2))
            self.likelihood_post_ups_path.append(increase_resolution(times=self.lvl_diff, input_dim=input,
                                                                     output_dim=input))
        self.likelihood_post_c_path = nn.ModuleList()
        for i in range(latent_levels - 1):
            input = self.num_filters[i] + self.num_filters[i + 1 + self.lvl_diff]
            output = self.num_filters[i + self.lvl_diff]
            if reversible:
                self.likelihood_post_c_path.append(ReversibleSequence(input_dim=input, output_dim=output,
                                                                      reversible_depth=2))
            else:
                self.likelihood_post_c_path.append(Conv2DSequence(input_dim=input, output_dim=output, depth=2))
        self.s_layer = nn.ModuleList()
        output = self.num_classes
        for i in reversed(range(self.latent_levels)):
            input = self.num_filters[i + self.lvl_diff]
            self.s_layer.append(Conv2DSequence(input_dim=input, output_dim=output, depth=1, kernel=1,
                                               activation=torch.nn.Identity, norm=torch.nn.Identity))

    def forward(self, z, lowres=False):
        """z: list of latent tensors [B,2,r,r] fp32 (index = latent level) -> list of full-resolution logits.
        ``lowres`` (extension, evaluation): return the logits at the resolution of their feature maps together with the
        nearest-upsampling factors, [(s_in [B,C,h,w], factor), ...] -- the fused evaluation kernel (b200.train.EvalStep)
        does the replication of models/phiseg.py:321 as index arithmetic instead of reading five full-size tensors."""
        s = [None] * self.latent_levels
        post_z = [None] * self.latent_levels
        post_c = [None] * self.latent_levels
        # the per-level branches are independent: the small ones run on a second stream next to the full-resolution one
        fork = (_Fork(z[0].device, _LIKELIHOOD_STREAMS)
                if (z[0].is_cuda and _use_streams(z[0], self.image_size[1] * self.image_size[2])) else None)
        for i in range(self.latent_levels):
            assert z[-i - 1].shape[1] == 2
            assert z[-i - 1].shape[2] == self.image_size[1] * 2 ** (-self.resolution_levels + 1 + i)
            with (fork.side(i) if (fork is not None and i != self.latent_levels - 1) else _null_ctx()):
                x = self.likelihood_ups_path[i](ops.to_act(z[-i - 1]))
                x = self.likelihood_post_ups_path[i](x)
                if fork is not None and i != self.latent_levels - 1:
                    fork.hand_over(x.t)
            assert x.t.shape[1] == self.image_size[1] * 2 ** (-self.latent_levels + i + 1)
            assert x.c == self.num_filters[-i - 1 - self.lvl_diff], '{} != {}'.format(x.c, self.num_filters[-i - 1])
            post_z[-i - 1] = x
        if fork is not None:
            fork.join()
        # The class logits of a level (1x1 conv + nearest upsampling, models/phiseg.py:319-321) depend only on that level's
        # post_c: they are issued as soon as it exists, the low-resolution ones on side streams next to the top-down chain
        # (five of them back to back after the chain were 75 us at the tail of forward -- and, replayed by autograd on the
        # same streams, 130 us at the head of backward).
        sfork = _Fork(z[0].device, self.latent_levels - 1, tag='slayer') if (fork is not None and _EARLY_LOGITS) else None

        def emit_logits(lvl):
            feat = post_c[lvl]
            conv = self.s_layer[self.latent_levels - 1 - lvl].convolution[0].convolution[0]
            factor = self.image_size[1] // feat.t.shape[1]
            assert factor * feat.t.shape[1] == self.image_size[1] and factor * feat.t.shape[2] == self.image_size[2]
            if sfork is not None and lvl != 0:
                st = sfork.streams[lvl - 1]
                st.wait_stream(sfork.main)
                feat.t.record_stream(st)
                with torch.cuda.stream(st):
                    out = ops.SLayerNearest.apply(feat.t, conv.weight, conv.bias, 1 if lowres else factor)
                sfork.hand_over(out)
            else:
                out = ops.SLayerNearest.apply(feat.t, conv.weight, conv.bias, 1 if lowres else factor)
            s[lvl] = (out, factor) if lowres else out

        post_c[self.latent_levels - 1] = post_z[self.latent_levels - 1]
        emit_logits(self.latent_levels - 1)
        for i in reversed(range(self.latent_levels - 1)):
            below = post_c[i + 1]
            assert post_z[i].t.shape[1] == 2 * below.t.shape[1] and post_z[i].t.shape[2] == 2 * below.t.shape[2]
            # bilinear x2 of the level below written straight into the concat buffer (no intermediate tensor)
            concat = Act(ops.Concat.apply(post_z[i].t, below.t, False, True, True), post_z[i].c + below.c)
            post_c[i] = self.likelihood_post_c_path[i](concat)
            emit_logits(i)
        if sfork is not None:
            sfork.join()
        return s


class PHISeg(nn.Module):
    """PHiSeg (https://arxiv.org/abs/1906.04045) behind the reference's module API (models/phiseg.py:326-537)."""

    def __init__(self, input_channels, num_classes, num_filters, latent_levels=5, latent_dim=2, initializers=None,
                 no_convs_fcomb=4, beta=10.0, image_size=(128, 128, 1), reversible=False, apply_last_layer=True,
                 exponential_weighting=True, padding=True):
        super(PHISeg, self).__init__()
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

        self.posterior = Posterior(input_channels, num_classes, num_filters, initializers=None, padding=True,
                                   reversible=reversible)
        self.likelihood = Likelihood(input_channels, num_classes, num_filters, initializers=None,
                                     apply_last_layer=True, padding=True, image_size=self.image_size,
                                     reversible=reversible)
        self.prior = Posterior(input_channels, num_classes, num_filters, initializers=None, padding=True,
                               is_posterior=False, reversible=reversible)
        self.s_out_list = [None] * self.latent_levels
        self.s_out_list_with_softmax = [None] * self.latent_levels

    # ---------------------------------------------------------------- sampling
    def sample_posterior(self):
        z_sample = [None] * self.latent_levels
        for i, _ in enumerate(z_sample):
            z_sample[i] = self.posterior_mu[i] + self.posterior_sigma[i] * torch.randn_like(self.posterior_sigma[i])
        return z_sample

    def sample_prior(self):
        z_sample = [None] * self.latent_levels
        for i, _ in enumerate(z_sample):
            z_sample[i] = self.prior_mu[i] + self.prior_sigma[i] * torch.randn_like(self.prior_sigma[i])
        return z_sample

    def sample(self, testing=True):
        if testing:
            sample, _ = self.reconstruct(self.sample_prior(), use_softmax=False)
            return sample
        else:
            raise NotImplementedError

    def reconstruct(self, z_posterior, use_softmax=True):
        layer_recon = self.likelihood(z_posterior)
        return self.accumulate_output(layer_recon, use_softmax=use_softmax), layer_recon

    # ---------------------------------------------------------------- forward
    def _packer(self):
        # one batched launch packs every conv weight (bf16 forward + dgrad layouts) per step
        first = next(m.weight for m in self.modules() if isinstance(m, nn.Conv2d) and m.out_channels % 16 == 0)
        pk = getattr(self, '_weight_packer', None)
        if pk is None or not pk.valid_for(first):
            ws = [m.weight for m in self.modules()
                  if isinstance(m, nn.Conv2d) and m.out_channels % 16 == 0 and m.weight.is_cuda]
            pk = kern.WeightPacker(ws)
            object.__setattr__(self, '_weight_packer', pk)
        return pk

    def forward(self, patch, mask, training=True, replicate=1, lowres_logits=False):
        """``replicate`` (extension, evaluation only): the N-sample evaluation of the reference feeds N identical copies of
        one image (train_model.py:177-179).  forward(patch[I,...], mask[I,...], training=False, replicate=N) returns what
        forward(patch.repeat(N,1,1,1), mask.repeat(N,1,1,1), training=False) returns -- same values, same random draws
        (batch index = copy * I + image) -- but runs the two encoders once per image instead of N times (their eval-mode
        outputs are identical for identical inputs).  ``lowres_logits``: see Likelihood.forward (evaluation only; the
        returned list then holds (tensor, factor) pairs and ``s_out_list`` is not updated)."""
        if not patch.is_cuda:
            raise kern._lib.UnetZooLibError('UNet-Zoo B200 modules need CUDA tensors: there is no CPU fallback path')
        if (replicate != 1 or lowres_logits) and (training or self.training):
            raise ValueError('replicate=N / lowres_logits need eval mode and training=False')
        object.__setattr__(self, '_tg_active', None)
        if training and self.training and torch.is_grad_enabled() and mask is not None and \
                (_TRANSPARENT or getattr(self, 'transparent_graph', False)) and \
                not torch.cuda.is_current_stream_capturing():
            return self._transparent_forward(patch, mask)
        return self._plain_forward(patch, mask, training, replicate, lowres_logits)

    def _folder(self):
        # eval mode: the BatchNorm folds of every conv unit in one launch
        from torchlayers import Conv2D as _C2D
        units = [m for m in self.modules() if isinstance(m, _C2D) and isinstance(m.convolution[1], m._norm_cls)]
        fd = getattr(self, '_bn_folder', None)
        if fd is None or not fd.valid_for(units[0].convolution[1].running_mean):
            fd = kern.BNFolder([(m.convolution[0].bias, m.convolution[1].weight, m.convolution[1].bias,
                                 m.convolution[1].running_mean, m.convolution[1].running_var) for m in units])
            object.__setattr__(self, '_bn_folder', fd)
        return fd

    def _plain_forward(self, patch, mask, training=True, replicate=1, lowres_logits=False):
        pk = self._packer()
        pk.refresh()
        kern.zero_arena.reset(patch.device)
        prev = kern.set_active_packer(pk)
        prev_f = None
        if not self.training:
            fd = self._folder()
            fd.refresh()
            prev_f = kern.set_active_folder(fd)
        try:
            with deferred_batch_counts():
                return self._forward(patch, mask, training, replicate, lowres_logits)
        finally:
            kern.set_active_packer(prev)
            if not self.training:
                kern.set_active_folder(prev_f)

    @staticmethod
    def _replicate(x, blocks, n):
        if n == 1:
            return x, blocks
        # the deepest feature map feeds convolutions and becomes a real batch of n copies (copy-major: index = copy * I +
        # image); the skip connections stay single images -- the concat kernel replicates them while it copies them into
        # the concat buffers (kern.copy_channels)
        return Act(x.t.repeat((n,) + (1,) * (x.t.dim() - 1)), x.c), blocks

    def _forward(self, patch, mask, training=True, replicate=1, lowres_logits=False):
        # posterior and prior encoders are independent (no random draws inside): the prior's runs on a second stream.
        # The latent halves keep the reference's order -- 5 posterior draws, then 5 prior draws (quirk Q4).
        if replicate != 1:
            post_x, post_blocks = self._replicate(*self.posterior.contract(patch, mask), replicate)
            prior_x, prior_blocks = self._replicate(*self.prior.contract(patch), replicate)
        elif _use_streams(patch):
            fork = _Fork(patch.device)
            with fork.side():
                prior_x, prior_blocks = self.prior.contract(patch)
                fork.hand_over(prior_x.t, *[b.t for b in prior_blocks])
            post_x, post_blocks = self.posterior.contract(patch, mask)
            fork.join()
        else:
            post_x, post_blocks = self.posterior.contract(patch, mask)
            prior_x, prior_blocks = self.prior.contract(patch)
        self.posterior_latent_space, self.posterior_mu, self.posterior_sigma = self.posterior.latent(post_x, post_blocks)
        if training and replicate == 1 and _PRIOR_OVERLAP and _use_streams(patch):
            # teacher forcing: the prior's latent path consumes the POSTERIOR samples, the likelihood too -- they are
            # independent of each other, so the prior's runs on a side stream next to the likelihood.  Its five random
            # draws are still issued (host order = Philox offsets) right after the posterior's, before any other draw.
            fork = _Fork(patch.device, 1, tag='prior')
            with fork.side():
                self.prior_latent_space, self.prior_mu, self.prior_sigma = self.prior.latent(
                    prior_x, prior_blocks, training_prior=True, z_list=self.posterior_latent_space)
                fork.hand_over(*self.prior_mu, *self.prior_sigma)
            self.s_out_list = self.likelihood(self.posterior_latent_space)
            fork.join()
        elif training:
            self.prior_latent_space, self.prior_mu, self.prior_sigma = self.prior.latent(
                prior_x, prior_blocks, training_prior=True, z_list=self.posterior_latent_space)
            self.s_out_list = self.likelihood(self.posterior_latent_space)
        else:
            self.prior_latent_space, self.prior_mu, self.prior_sigma = self.prior.latent(prior_x, prior_blocks)
            if lowres_logits:
                return self.likelihood(self.prior_latent_space, lowres=True)
            self.s_out_list = self.likelihood(self.prior_latent_space)
        return self.s_out_list

    def accumulate_output(self, output_list, use_softmax=False):
        """s_accum = output_list[-1]; s_accum += output_list[i] -- IN PLACE like the reference (quirk Q2)."""
        s_accum = output_list[-1]
        if torch.is_grad_enabled() and any(t.requires_grad for t in output_list):
            # autograd-tracked corner (sample() during training): same aliasing through torch's in-place add
            for i in range(len(output_list) - 1):
                s_accum += output_list[i]
            return torch.nn.functional.softmax(s_accum, dim=1) if use_softmax else s_accum
        lst = [t.contiguous() for t in output_list]
        if lst[-1].data_ptr() != output_list[-1].data_ptr():
            raise RuntimeError('accumulate_output needs a contiguous last element to accumulate in place')
        if use_softmax:
            # the sum must stay in output_list[-1] (validate() calls loss() on the mutated list); softmax is a new tensor
            kern.accumulate_output(lst, False, lst[-1])
            out = torch.empty_like(lst[-1])
            return kern.accumulate_output([lst[-1]], True, out)
        return kern.accumulate_output(lst, False, lst[-1])

    # ---------------------------------------------------------------- losses
    def KL_two_gauss_with_diag_cov(self, mu0, sigma0, mu1, sigma1):
        return ops.KLLevel.apply(mu0, sigma0, mu1, sigma1, 1.0)

    def calculate_hierarchical_KL_div_loss(self):
        if self.exponential_weighting:
            level_weights = [self.exponential_weight ** i for i in list(range(self.latent_levels))]
        else:
            level_weights = [1] * self.latent_levels
        # all levels in one fused op: total = sum over ii = L-1 ... 0 of kl_weight * w_ii * KL_ii (the reference's order of
        # `loss_tot += ...`), per-level values for the loss dictionary
        flat = []
        for ii in range(self.latent_levels):
            flat += [self.posterior_mu[ii], self.posterior_sigma[ii], self.prior_mu[ii], self.prior_sigma[ii]]
        total, levels = ops.KLHierarchy.apply(tuple(float(w) for w in level_weights),
                                              float(self.kl_divergence_loss_weight), *flat)
        for ii in reversed(range(self.latent_levels)):
            self.loss_dict['KL_divergence_loss_lvl%d' % ii] = levels[ii]
        self._loss_add(total)
        return self.loss_tot

    def _loss_add(self, term):
        # ``self.loss_tot += tensor``: an int on the first add (new tensor), in place afterwards -- which is what
        # makes kl_divergence_loss, reconstruction_loss and the returned loss ONE tensor in the reference (quirk Q1)
        if torch.is_tensor(self.loss_tot):
            self.loss_tot += term
        else:
            self.loss_tot = self.loss_tot + term

    def multinoulli_loss(self, reconstruction, target):
        total, _ = ops.ResidualCE.apply(target, reconstruction)
        return total

    def residual_multinoulli_loss(self, reconstruction, target):
        total, levels = ops.ResidualCE.apply(target, *reconstruction)
        for ii in range(self.latent_levels):
            self.loss_dict['residual_multinoulli_loss_lvl%d' % ii] = levels[ii]
        self.s_accumulated = [None] * self.latent_levels     # the level sums live only inside the fused kernel
        self._loss_add(self.residual_multinoulli_loss_weight * total)
        return self.loss_tot

    def kl_divergence(self):
        return self.calculate_hierarchical_KL_div_loss()

    def elbo(self, segm, reconstruct_posterior_mean=False):
        self.loss_tot = 0
        self.kl_divergence_loss = self.kl_divergence()
        self.reconstruction_loss = self.residual_multinoulli_loss(reconstruction=self.s_out_list, target=segm)
        return self.loss_tot

    def loss(self, segm):
        st = getattr(self, '_tg_active', None)
        if st is not None and torch.is_grad_enabled():
            if segm.data_ptr() != st['mask_ptr'] and not torch.equal(segm.reshape(st['mask'].shape).float(), st['mask']):
                raise RuntimeError('transparent-graph mode: loss(mask) must receive the mask the preceding '
                                   'forward(patch, mask, training=True) was given (reference train_model.py:111-112)')
            out = _GraphLoss.apply(st['loss'], st, *st['params'])
            # quirk Q1: the three loss handles are one tensor
            self.loss_tot = out
            self.kl_divergence_loss = out
            self.reconstruction_loss = out
            object.__setattr__(self, '_tg_active', None)
            return out
        return self.elbo(segm)

    # ---------------------------------------------------------------- transparent graph capture (unmodified caller)
    def _transparent_forward(self, patch, mask):
        params = [p for p in self.parameters() if p.requires_grad]
        key = (tuple(patch.shape), tuple(mask.shape), patch.device.index, params[0].data_ptr())
        table = self.__dict__.setdefault('_tg', {})
        st = table.setdefault(key, {'calls': 0})
        if 'graph_fwd' not in st:
            st['calls'] += 1
            if st['calls'] <= _TRANSPARENT_WARMUP:
                return self._plain_forward(patch, mask, True)           # eager: the caller's first steps run as before
            self._transparent_capture(st, patch, mask, params)
        st['patch'].copy_(patch)
        st['mask'].copy_(mask.reshape(st['mask'].shape))
        st['mask_ptr'] = mask.data_ptr()
        st['graph_fwd'].replay()
        for k, v in st['attrs'].items():
            setattr(self, k, v)
        object.__setattr__(self, '_tg_active', st)
        return self.s_out_list

    def _swap_parameters(self, mapping):
        for m in self.modules():
            for name, p in list(m._parameters.items()):
                if p is not None and id(p) in mapping:
                    m._parameters[name] = mapping[id(p)]

    def _transparent_capture(self, st, patch, mask, params):
        dev = patch.device
        st['patch'] = patch.detach().float().clone()
        st['mask'] = mask.detach().float().clone()
        torch.cuda.synchronize(dev)
        cap = torch.cuda.Stream(device=dev, priority=int(os.environ.get('UNETZOO_MAIN_PRIORITY', '-2')))
        graph = torch.cuda.CUDAGraph()
        # The capture runs on ALIASES of the parameters (new leaf tensors on the same storage): the caller's eager
        # warm-up steps created the parameters' AccumulateGrad nodes on the legacy default stream and the caller still
        # holds the previous loss (hence those nodes) -- gradients accumulated there would tie the legacy stream to the
        # capture and invalidate it.  Fresh leaves get fresh nodes on the capture stream.
        alias = {id(p): nn.Parameter(p.detach(), requires_grad=True) for p in params}
        back = {id(a): p for p, a in ((p, alias[id(p)]) for p in params)}
        self._swap_parameters(alias)
        try:
            mark = len(kern.wgrad_reducer.keep)
            with torch.cuda.graph(graph, stream=cap):
                self._plain_forward(st['patch'], st['mask'], True)
                loss = self.elbo(st['mask'])
                st['attrs'] = {k: getattr(self, k) for k in _TG_ATTRS}
                st['attrs']['loss_dict'] = dict(self.loss_dict)
                loss.backward()
            grads = [alias[id(p)].grad for p in params]
        finally:
            self._swap_parameters(back)         # the caller's Parameter objects (and its optimizer's references) are back
        torch.cuda.synchronize(dev)
        st.update(graph_fwd=graph, loss=loss.detach(), grads=grads, params=params,
                  slabs=kern.wgrad_reducer.keep[mark:])        # weight-gradient slabs of the captured launches
        del kern.wgrad_reducer.keep[mark:]
