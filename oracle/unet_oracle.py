"""CPU restatement (plain torch fp32) of the reference U-Net and Probabilistic U-Net -- TEST INFRASTRUCTURE.

  Unet                      reference models/unet.py:12-165   (conv3x3 + bias + ReLU, no BN; AvgPool2d(2,2,ceil);
                            bilinear x2 align_corners=False; cat([up, bridge]); 1x1 last layer; CE mean)
  Encoder / AxisAlignedConvGaussian / Fcomb / ProbabilisticUnet
                            reference models/probabilistic_unet.py:20-370
Operates on state_dicts with the reference's key names; pinned by tests/test_oracle_unet.py against the live reference
(build container) and fixtures generated from it (oracle/make_golden.py).
"""
import torch
import torch.nn.functional as F

from .phiseg_oracle import FP32, conv2d_unit, kl_two_gauss, multinoulli, onehot_minus_half


def _plain_block(x, sd, prefix, pool, rnd):
    """DownConvBlock (unet.py:12-40): [pool] + 3 x (conv + ReLU); Sequential indices shift by one with a pool."""
    off = 0
    if pool:
        x = rnd.act(F.avg_pool2d(x, 2, 2, 0, ceil_mode=True))
        off = 1
    for k in range(3):
        i = off + 2 * k
        x = F.relu(F.conv2d(rnd.act(x), rnd.weight(sd['%s.layers.%d.weight' % (prefix, i)]),
                            sd['%s.layers.%d.bias' % (prefix, i)], padding=1))
        x = rnd.act(x)
    return x


def unet_features(sd, x, num_levels, prefix='', rnd=FP32):
    blocks = []
    for i in range(num_levels):
        x = _plain_block(x, sd, '%scontracting_path.%d' % (prefix, i), i != 0, rnd)
        if i != num_levels - 1:
            blocks.append(x)
    for i in range(num_levels - 1):
        up = rnd.act(F.interpolate(x, mode='bilinear', scale_factor=2, align_corners=False))
        x = torch.cat([up, blocks[-i - 1]], 1)
        x = _plain_block(x, sd, '%supsampling_path.%d.conv_block' % (prefix, i), False, rnd)
    return x


def unet_forward(sd, x, num_levels, rnd=FP32):
    f = unet_features(sd, x, num_levels, '', rnd)
    return F.conv2d(f, sd['last_layer.weight'], sd['last_layer.bias'])


def unet_loss(logits, mask):
    """unet.py:159-165: CrossEntropyLoss() mean over all pixels."""
    return F.cross_entropy(logits, mask.view(-1, logits.shape[2], logits.shape[3]).long())


def gaussian_head(sd, name, x, num_levels, latent_dim, training, segm=None, rnd=FP32):
    """AxisAlignedConvGaussian.forward (probabilistic_unet.py:102-130) -> (mu, sigma)."""
    if segm is not None:
        x = torch.cat([x, onehot_minus_half(segm, 2)], dim=1)
    for i in range(num_levels):
        if i != 0:
            x = rnd.act(F.avg_pool2d(x, 2, 2, 0, ceil_mode=True))
        li = 2 * i                                   # Sequential index: pools sit at odd positions
        for k in range(3):
            x = conv2d_unit(x, sd, '%s.encoder.layers.%d.convolution.%d' % (name, li, k), training, rnd)
    enc = rnd.act(x.mean(dim=2, keepdim=True).mean(dim=3, keepdim=True))
    mls = F.conv2d(enc, sd[name + '.conv_layer.weight'], sd[name + '.conv_layer.bias'])[:, :, 0, 0]
    return mls[:, :latent_dim], torch.exp(mls[:, latent_dim:])


def fcomb(sd, features, z, n_layers, training, rnd=FP32):
    """Fcomb.forward (probabilistic_unet.py:185-199): tile z, cat, 1x1 Conv2D stack, 1x1 last layer."""
    b, _, h, w = features.shape
    x = torch.cat([features, z[:, :, None, None].expand(b, z.shape[1], h, w)], dim=1)
    for k in range(n_layers):
        x = conv2d_unit(x, sd, 'fcomb.layers.%d' % k, training, rnd, kernel=1)
    return F.conv2d(x, sd['fcomb.last_layer.weight'], sd['fcomb.last_layer.bias'])


def probunet_step(sd, patch, mask, eps_post, num_levels, latent_dim, no_convs_fcomb, training=True, rnd=FP32):
    """forward(patch, mask) + loss(mask) of ProbabilisticUnet (probabilistic_unet.py:246-370) with the posterior
    rsample noise injected.  Returns dict of the quantities the caller can observe."""
    mu_q, sig_q = gaussian_head(sd, 'posterior', patch, num_levels, latent_dim, training, segm=mask, rnd=rnd)
    mu_p, sig_p = gaussian_head(sd, 'prior', patch, num_levels, latent_dim, training, rnd=rnd)
    feats = unet_features(sd, patch, num_levels, 'unet.', rnd)
    out = F.conv2d(feats, sd['last_conv.convolution.0.weight'], sd['last_conv.convolution.0.bias'])
    z = mu_q + sig_q * eps_post
    kl = kl_two_gauss(mu_q, sig_q, mu_p, sig_p)
    recon = fcomb(sd, feats, z, no_convs_fcomb - 1, training, rnd)
    rec_loss = multinoulli(recon, mask)
    elbo = -(rec_loss + 1.0 * kl)
    reg = 0
    for k, v in sd.items():
        if (k.startswith('posterior.') or k.startswith('prior.') or k.startswith('fcomb.layers.')) and \
                v.dtype == torch.float32 and 'running_' not in k:
            reg = reg + v.norm(2)
    return {'forward': out, 'mu_q': mu_q, 'sigma_q': sig_q, 'mu_p': mu_p, 'sigma_p': sig_p, 'kl': kl,
            'reconstruction': recon, 'reconstruction_loss': rec_loss, 'loss': -elbo + 1e-5 * reg}
