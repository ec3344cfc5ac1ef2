"""CPU restatement (plain torch fp32 ops) of the reference PHISeg3D forward / ELBO -- TEST INFRASTRUCTURE.

Follows models/phiseg3D.py of the reference under the FIXED SPECIFICATION of SURVEY.md 8c (the file as shipped cannot
finish a forward pass):
  Conv3D                      models/phiseg3D.py:13-35    (conv3d + bias -> BatchNorm3d(eps 1e-3, momentum 0.01) -> ReLU)
  ReversibleSequence          models/phiseg3D.py:61-88    (1x1x1 Conv3D if widths differ, then depth-1 additive coupling)
  DownConvolutionalBlock      models/phiseg3D.py:91-118   ([AvgPool3d(2,2,ceil)] + 3 x Conv3D | rev depth 1)
  UpConvolutionalBlock        models/phiseg3D.py:121-153  (trilinear x2 align_corners=True -> 2 x Conv3D -> cat)
  SampleZBlock                models/phiseg3D.py:156-190
  Posterior.forward           models/phiseg3D.py:248-285  (fix iii: one-hot with num_classes labels of an index volume)
  increase_resolution         models/phiseg3D.py:288-301
  Likelihood.forward          models/phiseg3D.py:372-400  (fix ii: nearest upsample to image_size[1:4])
  PHISeg3D.forward / elbo     models/phiseg3D.py:488-611  (same sigma1*sigma0 KL and residual CE as the 2-D file)
Pinned by tests/test_oracle_phiseg3d.py against the reference module itself (with the three patches of
oracle/ref_run.py) and by tests/golden/phiseg3d_small.npz generated from it.
"""
import torch
import torch.nn.functional as F

from .phiseg_oracle import BN_EPS, BN_MOMENTUM, FP32, kl_two_gauss, multinoulli, onehot_minus_half


def conv3d_unit(x, sd, prefix, training, rnd=FP32, kernel=3, updates=1):
    w = sd[prefix + '.convolution.0.weight']
    b = sd[prefix + '.convolution.0.bias']
    y = rnd.act(F.conv3d(rnd.act(x), rnd.weight(w), b, padding=1 if kernel == 3 else 0))
    g = sd[prefix + '.convolution.1.weight']
    beta = sd[prefix + '.convolution.1.bias']
    rm = sd[prefix + '.convolution.1.running_mean']
    rv = sd[prefix + '.convolution.1.running_var']
    if training:
        mean = y.mean(dim=(0, 2, 3, 4))
        var = y.var(dim=(0, 2, 3, 4), unbiased=False)
        n = y.numel() // y.shape[1]
        with torch.no_grad():
            for _ in range(updates):
                rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean)
                rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var * n / max(n - 1, 1))
                nbt = prefix + '.convolution.1.num_batches_tracked'
                if nbt in sd:
                    sd[nbt] += 1
    else:
        mean, var = rm, rv
    scale = g * torch.rsqrt(var + BN_EPS)
    shift = beta - mean * scale
    y = y * scale[None, :, None, None, None] + shift[None, :, None, None, None]
    return rnd.act(F.relu(y))


def _is_rev(sd, prefix):
    return any(k.startswith(prefix + '.sequence.reversible_blocks.') for k in sd)


def rev_sequence(x, sd, prefix, training, rnd=FP32, recompute=True):
    """depth-1 reversible stack (models/phiseg3D.py:61-88 over revtorch additive coupling)"""
    if prefix + '.inital_conv.convolution.0.weight' in sd:
        x = conv3d_unit(x, sd, prefix + '.inital_conv', training, rnd, kernel=1)
    upd = 2 if (training and recompute) else 1
    x1, x2 = torch.chunk(x, 2, dim=1)
    base = prefix + '.sequence.reversible_blocks.0'
    y1 = rnd.act(x1 + conv3d_unit(x2, sd, base + '.f_block.0', training, rnd, updates=upd))
    y2 = rnd.act(x2 + conv3d_unit(y1, sd, base + '.g_block.0', training, rnd, updates=upd))
    return torch.cat([y1, y2], dim=1)


def up2(x):
    return F.interpolate(x, mode='trilinear', scale_factor=2, align_corners=True)


def conv_stack(x, sd, prefix_fmt, count, training, rnd, rev_prefix):
    """either ``count`` Conv3D layers at prefix_fmt % k or one reversible stack at rev_prefix"""
    if _is_rev(sd, rev_prefix):
        return rev_sequence(x, sd, rev_prefix, training, rnd)
    for k in range(count):
        x = conv3d_unit(x, sd, prefix_fmt % k, training, rnd)
    return x


def encoder_decoder(patch, sd, name, eps_list, training, latent_levels, resolution_levels, rnd=FP32, segm=None,
                    num_classes=2, z_forced=None):
    if segm is not None:
        patch = torch.cat([patch, onehot_minus_half(segm, num_classes)], dim=1)
    blocks = []
    x = patch
    for i in range(resolution_levels):
        p = '%s.contracting_path.%d' % (name, i)
        off = 0
        if i != 0:
            x = rnd.act(F.avg_pool3d(x, 2, 2, 0, ceil_mode=True))
            off = 1
        if _is_rev(sd, '%s.layers.%d' % (p, off)):
            x = rev_sequence(x, sd, '%s.layers.%d' % (p, off), training, rnd)
        else:
            x = _plain_layers(x, sd, p, off, training, rnd)
        if i != resolution_levels - 1:
            blocks.append(x)
    z = [None] * latent_levels
    mu = [None] * latent_levels
    sigma = [None] * latent_levels
    pre = x
    for i in range(latent_levels):
        lvl = latent_levels - 1 - i
        if i != 0:
            u = rnd.act(up2(rnd.act(z[lvl + 1])))
            up_prefix = '%s.upsampling_path.%d.upconv_layer' % (name, i - 1)
            u = conv_stack(u, sd, up_prefix + '.%d', 2, training, rnd, up_prefix)
            pre = torch.cat([u, blocks[-i]], dim=1)
        sp = '%s.sample_z_path.%d' % (name, i)
        h = conv_stack(pre, sd, sp + '.conv.%d', 2, training, rnd, sp + '.conv.0')
        mu[lvl] = F.conv3d(h, sd[sp + '.mu_conv.0.weight'], sd[sp + '.mu_conv.0.bias'])
        sigma[lvl] = F.softplus(F.conv3d(h, sd[sp + '.sigma_conv.0.weight'], sd[sp + '.sigma_conv.0.bias']))
        z[lvl] = mu[lvl] + sigma[lvl] * eps_list[i]
        if z_forced is not None:
            z[lvl] = z_forced[lvl]
    return z, mu, sigma


def _plain_layers(x, sd, p, off, training, rnd):
    for k in range(3):
        x = conv3d_unit(x, sd, '%s.layers.%d' % (p, k + off), training, rnd)
    return x


def likelihood(z, sd, image_dhw, training, latent_levels, resolution_levels, rnd=FP32):
    L = latent_levels
    lvl_diff = resolution_levels - latent_levels
    post_z = [None] * L
    for i in range(L):
        lvl = L - 1 - i
        p = 'likelihood.likelihood_ups_path.%d' % i
        x = conv_stack(rnd.act(z[lvl]), sd, p + '.convolution.%d', 2, training, rnd, p)
        for t in range(lvl_diff):
            x = rnd.act(up2(x))
            x = conv3d_unit(x, sd, 'likelihood.likelihood_post_ups_path.%d.%d.convolution.0' % (i, 2 * t + 1), training,
                            rnd)
        post_z[lvl] = x
    post_c = [None] * L
    post_c[L - 1] = post_z[L - 1]
    for lvl in reversed(range(L - 1)):
        x = torch.cat([post_z[lvl], rnd.act(up2(post_c[lvl + 1]))], dim=1)
        p = 'likelihood.likelihood_post_c_path.%d' % lvl
        post_c[lvl] = conv_stack(x, sd, p + '.convolution.%d', 2, training, rnd, p)
    s = [None] * L
    for i in range(L):
        lvl = L - 1 - i
        p = 'likelihood.s_layer.%d.convolution.0.convolution.0' % i
        s_in = F.conv3d(post_c[lvl], sd[p + '.weight'], sd[p + '.bias'])
        s[lvl] = F.interpolate(s_in, size=list(image_dhw), mode='nearest')
    return s


def phiseg3d_forward(sd, patch, mask, eps_list, latent_levels, resolution_levels, num_classes, training=True, rnd=FP32):
    """eps_list: 2 * latent_levels tensors: posterior draws (deepest level first) then prior draws."""
    L = latent_levels
    dhw = patch.shape[-3:]
    pz, pmu, psig = encoder_decoder(patch, sd, 'posterior', eps_list[:L], training, L, resolution_levels, rnd,
                                    segm=mask, num_classes=num_classes)
    if training:
        qz, qmu, qsig = encoder_decoder(patch, sd, 'prior', eps_list[L:], training, L, resolution_levels, rnd,
                                        z_forced=pz)
        s = likelihood(pz, sd, dhw, training, L, resolution_levels, rnd)
    else:
        qz, qmu, qsig = encoder_decoder(patch, sd, 'prior', eps_list[L:], training, L, resolution_levels, rnd)
        s = likelihood(qz, sd, dhw, training, L, resolution_levels, rnd)
    return {'post_z': pz, 'post_mu': pmu, 'post_sigma': psig, 'prior_z': qz, 'prior_mu': qmu, 'prior_sigma': qsig,
            's': s}


def elbo(out, segm):
    """models/phiseg3D.py:529-611 (clone of the 2-D ELBO with latent_levels from the model)"""
    L = len(out['s'])
    kl_lvls = [(lvl, (4 ** lvl) * kl_two_gauss(out['post_mu'][lvl], out['post_sigma'][lvl], out['prior_mu'][lvl],
                                               out['prior_sigma'][lvl])) for lvl in reversed(range(L))]
    ce_lvls = []
    acc = None
    for lvl in reversed(range(L)):
        acc = out['s'][lvl] if acc is None else acc + out['s'][lvl]
        ce_lvls.append((lvl, multinoulli(acc, segm)))
    kl = sum(v for _, v in kl_lvls)
    recon = sum(v for _, v in ce_lvls)
    return {'total': kl + recon, 'kl': kl, 'recon': recon, 'kl_levels': dict(kl_lvls), 'ce_levels': dict(ce_lvls)}
