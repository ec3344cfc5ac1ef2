"""CPU restatement (plain torch fp32 ops) of the reference PHiSeg forward / ELBO -- TEST INFRASTRUCTURE.

Follows, function by function:
  Conv2D                      reference torchlayers.py:7-29   (conv + bias -> BN(eps 1e-3, momentum 0.01) -> ReLU)
  DownConvolutionalBlock      models/phiseg.py:14-39          ([AvgPool2d(2,2,ceil)] + 3 x Conv2D)
  UpConvolutionalBlock        models/phiseg.py:42-73          (bilinear x2 align_corners=True -> 2 x Conv2D -> cat)
  SampleZBlock                models/phiseg.py:76-106         (2 x Conv2D; mu = 1x1; sigma = softplus(1x1); z = mu + sigma*eps)
  Posterior.forward           models/phiseg.py:175-206        (one-hot(mask)-0.5 concat; teacher forcing AFTER the draw)
  increase_resolution         models/phiseg.py:209-221
  Likelihood.forward          models/phiseg.py:286-323
  PHISeg.forward              models/phiseg.py:414-426
  accumulate_output           models/phiseg.py:428-434
  KL_two_gauss_with_diag_cov  models/phiseg.py:436-453        (sigma1*sigma0 quirk)
  hierarchical KL / residual CE / elbo   models/phiseg.py:455-537
It operates on a state_dict with the reference's key names, so it runs where
``/root/reference`` is absent (the GPU box).  It is pinned by tests/test_oracle_vs_reference.py (which
runs the real reference modules in the build container on the same seeded weights / inputs / noise)
and by the committed fixtures under tests/golden/ generated from the real reference
(oracle/make_golden.py).

``Rounding`` lets the oracle emulate the storage precision of the CUDA path (bf16 activations and
conv weights, fp32 accumulation and fp32 heads / losses) so kernels can be checked tightly; the
default is exact fp32 = the reference's arithmetic.
"""
import torch
import torch.nn.functional as F

LATENT_LEVELS = 5          # hard-coded in the reference: models/phiseg.py:131-132
RESOLUTION_LEVELS = 7
LVL_DIFF = RESOLUTION_LEVELS - LATENT_LEVELS
BN_EPS = 1e-3              # torchlayers.py:20
BN_MOMENTUM = 0.01


class Rounding:
    """fp32 (reference arithmetic) or storage-precision emulation of the CUDA path: bf16 (product library) or IEEE half
    (libunetzoo_b200_fp16.so, the tolerance-matched parity mode)."""

    def __init__(self, bf16=False, fp16=False):
        self.bf16 = bf16
        self.fp16 = fp16

    def act(self, x):
        if self.bf16:
            return x.to(torch.bfloat16).to(torch.float32)
        if self.fp16:
            return x.to(torch.float16).to(torch.float32)
        return x

    weight = act


FP32 = Rounding(False)

# baseline arm only (oracle/torch_cuda_arm.py): BatchNorm through torch's own fused kernel (F.batch_norm -> cuDNN / ATen),
# which is what the reference's nn.BatchNorm2d dispatches to; the default restates it with elementwise ops so that the
# statistics can be captured and rounding points emulated.
NATIVE_BN = False


def onehot_minus_half(mask, nlabels=2):
    """utils.py:289-311 + models/phiseg.py:176-183: channels k<nlabels = (mask==k) - 0.5."""
    return torch.cat([(mask == k).to(torch.float32) for k in range(nlabels)], dim=1) - 0.5


def conv2d_unit(x, sd, prefix, training, rnd=FP32, kernel=3, norm=True, act=True, stats=None, updates=1):
    """torchlayers.py:7-29.  ``prefix`` is the Conv2D module path; keys prefix.convolution.{0,1}.*.
    In training mode the BN running stats in ``sd`` are updated in place like nn.BatchNorm2d does."""
    w = sd[prefix + '.convolution.0.weight']
    b = sd[prefix + '.convolution.0.bias']
    y = F.conv2d(rnd.act(x), rnd.weight(w), b, padding=1 if kernel == 3 else 0)
    if norm:
        y = rnd.act(y)
        g = sd[prefix + '.convolution.1.weight']
        beta = sd[prefix + '.convolution.1.bias']
        rm = sd[prefix + '.convolution.1.running_mean']
        rv = sd[prefix + '.convolution.1.running_var']
        if NATIVE_BN:
            y = F.batch_norm(y, rm, rv, g, beta, training, BN_MOMENTUM, BN_EPS)
            if training:
                nbt = prefix + '.convolution.1.num_batches_tracked'
                if nbt in sd:
                    with torch.no_grad():
                        sd[nbt] += 1
            return F.relu(y) if act else y
        if training:
            mean = y.mean(dim=(0, 2, 3))
            var = y.var(dim=(0, 2, 3), unbiased=False)
            n = y.numel() // y.shape[1]
            with torch.no_grad():
                for _ in range(updates):      # reversible blocks run F and G twice per training step (quirk Q7)
                    rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean)
                    rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var * n / max(n - 1, 1))
                    nbt = prefix + '.convolution.1.num_batches_tracked'
                    if nbt in sd:
                        sd[nbt] += 1
        else:
            mean, var = rm, rv
        if stats is not None:
            stats[prefix] = (mean.detach().clone(), var.detach().clone())
        scale = g * torch.rsqrt(var + BN_EPS)
        shift = beta - mean * scale
        y = y * scale[None, :, None, None] + shift[None, :, None, None]
    if act:
        y = F.relu(y)
    return rnd.act(y) if (norm or act) else y


def rev_sequence(x, sd, prefix, depth, training, rnd=FP32, stats=None, recompute=True):
    """ReversibleSequence of the reference (torchlayers.py:55-82) over revtorch 0.2.0 additive coupling:
    optional 1x1 Conv2D, then depth x {y1 = x1 + F(x2); y2 = x2 + G(y1)}.  Gradients through this plain forward equal
    those of revtorch's inverse-recompute backward; with ``recompute`` (a training step that runs backward) the
    BatchNorm running statistics of F and G receive two momentum updates."""
    if prefix + '.inital_conv.convolution.0.weight' in sd:
        x = conv2d_unit(x, sd, prefix + '.inital_conv', training, rnd, kernel=1, stats=stats)
    upd = 2 if (training and recompute) else 1
    for d in range(depth):
        x1, x2 = torch.chunk(x, 2, dim=1)
        base = '%s.sequence.reversible_blocks.%d' % (prefix, d)
        y1 = rnd.act(x1 + conv2d_unit(x2, sd, base + '.f_block.0', training, rnd, stats=stats, updates=upd))
        y2 = rnd.act(x2 + conv2d_unit(y1, sd, base + '.g_block.0', training, rnd, stats=stats, updates=upd))
        x = torch.cat([y1, y2], dim=1)
    return x


def _is_rev(sd, prefix):
    return any(k.startswith(prefix + '.sequence.reversible_blocks.') for k in sd)


def up2(x):
    return F.interpolate(x, mode='bilinear', scale_factor=2, align_corners=True)


def down_block(x, sd, prefix, pool, training, rnd, stats=None):
    """models/phiseg.py:14-39 (non-reversible): layer indices shift by one when a pool is present."""
    off = 0
    if pool:
        x = rnd.act(F.avg_pool2d(x, 2, 2, 0, ceil_mode=True))
        off = 1
    if _is_rev(sd, '%s.layers.%d' % (prefix, off)):
        return rev_sequence(x, sd, '%s.layers.%d' % (prefix, off), 3, training, rnd, stats)
    for k in range(3):
        x = conv2d_unit(x, sd, '%s.layers.%d' % (prefix, k + off), training, rnd, stats=stats)
    return x


def sample_z_block(x, sd, prefix, eps, training, rnd, stats=None):
    """models/phiseg.py:76-106; ``eps`` replaces torch.randn_like(sigma)."""
    if _is_rev(sd, prefix + '.conv.0'):
        x = rev_sequence(x, sd, prefix + '.conv.0', 3, training, rnd, stats)
    else:
        for k in range(2):
            x = conv2d_unit(x, sd, '%s.conv.%d' % (prefix, k), training, rnd, stats=stats)
    mu = F.conv2d(x, sd[prefix + '.mu_conv.0.weight'], sd[prefix + '.mu_conv.0.bias'])
    sigma = F.softplus(F.conv2d(x, sd[prefix + '.sigma_conv.0.weight'], sd[prefix + '.sigma_conv.0.bias']))
    z = mu + sigma * eps
    return mu, sigma, z


def encoder_decoder(patch, sd, name, eps_list, training, rnd=FP32, segm=None, z_forced=None, stats=None):
    """Posterior.forward (models/phiseg.py:175-206); ``name`` is 'posterior' or 'prior'.
    eps_list: 5 noise tensors in DRAW order (latent level 4 first).  Returns z, mu, sigma lists indexed by level."""
    if segm is not None:
        patch = torch.cat([patch, onehot_minus_half(segm, 2)], dim=1)
    blocks = []
    x = patch
    for i in range(RESOLUTION_LEVELS):
        x = down_block(x, sd, '%s.contracting_path.%d' % (name, i), i != 0, training, rnd, stats)
        if i != RESOLUTION_LEVELS - 1:
            blocks.append(x)
    z = [None] * LATENT_LEVELS
    mu = [None] * LATENT_LEVELS
    sigma = [None] * LATENT_LEVELS
    pre = x
    for i in range(LATENT_LEVELS):
        lvl = LATENT_LEVELS - 1 - i
        if i != 0:
            u = rnd.act(up2(rnd.act(z[lvl + 1])))
            up_prefix = '%s.upsampling_path.%d.upconv_layer' % (name, i - 1)
            if _is_rev(sd, up_prefix):
                u = rev_sequence(u, sd, up_prefix, 2, training, rnd, stats)
            else:
                for k in range(2):
                    u = conv2d_unit(u, sd, '%s.%d' % (up_prefix, k), training, rnd, stats=stats)
            pre = torch.cat([u, blocks[-i]], dim=1)
        mu[lvl], sigma[lvl], z[lvl] = sample_z_block(pre, sd, '%s.sample_z_path.%d' % (name, i), eps_list[i], training, rnd, stats)
        if z_forced is not None:
            z[lvl] = z_forced[lvl]
    return z, mu, sigma


def likelihood(z, sd, image_hw, training, rnd=FP32, stats=None):
    """Likelihood.forward (models/phiseg.py:286-323).  Returns the list s[level] of full-res logits."""
    L = LATENT_LEVELS
    post_z = [None] * L
    for i in range(L):
        lvl = L - 1 - i
        x = rnd.act(z[lvl])
        if _is_rev(sd, 'likelihood.likelihood_ups_path.%d' % i):
            x = rev_sequence(x, sd, 'likelihood.likelihood_ups_path.%d' % i, 2, training, rnd, stats)
        else:
            for k in range(2):
                x = conv2d_unit(x, sd, 'likelihood.likelihood_ups_path.%d.convolution.%d' % (i, k), training, rnd,
                                stats=stats)
        for t in range(LVL_DIFF):
            x = rnd.act(up2(x))
            x = conv2d_unit(x, sd, 'likelihood.likelihood_post_ups_path.%d.%d.convolution.0' % (i, 2 * t + 1), training, rnd, stats=stats)
        post_z[lvl] = x
    post_c = [None] * L
    post_c[L - 1] = post_z[L - 1]
    for lvl in reversed(range(L - 1)):
        x = torch.cat([post_z[lvl], rnd.act(up2(post_c[lvl + 1]))], dim=1)
        if _is_rev(sd, 'likelihood.likelihood_post_c_path.%d' % lvl):
            x = rev_sequence(x, sd, 'likelihood.likelihood_post_c_path.%d' % lvl, 2, training, rnd, stats)
        else:
            for k in range(2):
                x = conv2d_unit(x, sd, 'likelihood.likelihood_post_c_path.%d.convolution.%d' % (lvl, k), training, rnd,
                                stats=stats)
        post_c[lvl] = x
    s = [None] * L
    for i in range(L):
        lvl = L - 1 - i
        p = 'likelihood.s_layer.%d.convolution.0.convolution.0' % i
        s_in = F.conv2d(post_c[lvl], sd[p + '.weight'], sd[p + '.bias'])
        s[lvl] = F.interpolate(s_in, size=list(image_hw), mode='nearest')
    return s


def kl_two_gauss(mu0, sigma0, mu1, sigma1):
    """models/phiseg.py:436-453 -- note sigma1*sigma0 where the textbook has sigma1^2 (quirk Q3)."""
    s0 = sigma0.flatten(1) * sigma0.flatten(1)
    s1 = sigma1.flatten(1) * sigma0.flatten(1)
    d = mu1.flatten(1) - mu0.flatten(1)
    return torch.mean(0.5 * torch.sum((s0 + d * d) / (s1 + 1e-10) + torch.log(s1 + 1e-10) - torch.log(s0 + 1e-10) - 1, dim=1))


def multinoulli(logits, target):
    """models/phiseg.py:481-490: CE summed over pixels, mean over batch."""
    b, c = logits.shape[:2]
    ce = F.cross_entropy(logits.reshape(b, c, -1), target.reshape(b, -1).long(), reduction='none')
    return ce.sum(dim=1).mean()


def elbo(out, segm):
    """models/phiseg.py:455-537.  Returns dict with total, kl (sum of weighted levels), recon, per-level terms."""
    kl_lvls = []
    for lvl in reversed(range(LATENT_LEVELS)):
        kl_lvls.append((lvl, (4 ** lvl) * kl_two_gauss(out['post_mu'][lvl], out['post_sigma'][lvl],
                                                     out['prior_mu'][lvl], out['prior_sigma'][lvl])))
    ce_lvls = []
    acc = None
    for lvl in reversed(range(LATENT_LEVELS)):
        acc = out['s'][lvl] if acc is None else acc + out['s'][lvl]
        ce_lvls.append((lvl, multinoulli(acc, segm)))
    kl = sum(v for _, v in kl_lvls)
    recon = sum(v for _, v in ce_lvls)
    return {'total': kl + recon, 'kl': kl, 'recon': recon,
            'kl_levels': dict(kl_lvls), 'ce_levels': dict(ce_lvls)}


def phiseg_forward(sd, patch, mask, eps_list, training=True, rnd=FP32, stats=None):
    """PHISeg.forward (models/phiseg.py:414-426).  eps_list: 10 tensors, posterior draws (levels 4..0)
    then prior draws (levels 4..0) -- quirk Q4.  Returns dict of lists indexed by latent level."""
    h, w = patch.shape[-2:]
    pz, pmu, psig = encoder_decoder(patch, sd, 'posterior', eps_list[:5], training, rnd, segm=mask, stats=stats)
    if training:
        qz, qmu, qsig = encoder_decoder(patch, sd, 'prior', eps_list[5:], training, rnd, z_forced=pz, stats=stats)
        s = likelihood(pz, sd, (h, w), training, rnd, stats)
    else:
        qz, qmu, qsig = encoder_decoder(patch, sd, 'prior', eps_list[5:], training, rnd, stats=stats)
        s = likelihood(qz, sd, (h, w), training, rnd, stats)
    return {'post_z': pz, 'post_mu': pmu, 'post_sigma': psig,
            'prior_z': qz, 'prior_mu': qmu, 'prior_sigma': qsig, 's': s}


def accumulate_output(s, use_softmax=False):
    """models/phiseg.py:428-434 (out-of-place here; the in-place aliasing quirk Q2 is a property of the
    module wrapper, tested at the boundary)."""
    acc = s[-1]
    for i in range(len(s) - 1):
        acc = acc + s[i]
    return F.softmax(acc, dim=1) if use_softmax else acc
