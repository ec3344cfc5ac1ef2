"""Seeded synthetic inputs and weights (host side, numpy): the benchmark's workload generator, also shared by the
oracle, the golden generator and the tests (re-exported as ``oracle.synth``).  Shapes follow SURVEY.md section 8(d):
LIDC-shaped patches fp32 [B,1,128,128] in [-0.5,0.5] (reference data/lidc_data_loader.py:92 stores
[0,1]-0.5), 4 annotator masks per image made of jittered filled ellipses with ~25 % empty
annotations; the training mask is one random annotator as float [B,1,H,W]
(reference data/batch_provider.py:61-63,131-137, train_model.py:103-106).

Weights are NOT stored in fixtures: ``synth_state_dict`` fills a state_dict from (key name, seed)
with numpy's frozen legacy ``RandomState`` stream, so generator and tests rebuild identical tensors.
"""
import zlib

import numpy as np
import torch


def _rs(seed, tag):
    return np.random.RandomState((zlib.crc32(tag.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)


def lidc_like_batch(batch, size=128, annotators=4, seed=0, empty_frac=0.25):
    """Returns patch fp32 [B,1,S,S], labels uint8 [B,S,S,M], train mask float32 [B,1,S,S]."""
    rs = _rs(seed, 'lidc')
    patch = np.clip(rs.standard_normal((batch, 1, size, size)) * 0.25, -0.5, 0.5).astype(np.float32)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    labels = np.zeros((batch, size, size, annotators), np.uint8)
    for b in range(batch):
        cy, cx = rs.uniform(0.3 * size, 0.7 * size, 2)
        ry, rx = rs.uniform(0.06 * size, 0.18 * size, 2)
        for m in range(annotators):
            if rs.uniform() < empty_frac:
                continue
            jy, jx = rs.normal(0, 0.02 * size, 2)
            sy, sx = rs.uniform(0.8, 1.25, 2)
            labels[b, :, :, m] = (((yy - cy - jy) / (ry * sy)) ** 2 + ((xx - cx - jx) / (rx * sx)) ** 2) <= 1.0
        # imprint the lesion on the image so the task is learnable
        patch[b, 0] += 0.2 * labels[b].mean(-1)
    pick = rs.randint(0, annotators, batch)
    mask = np.stack([labels[b, :, :, pick[b]] for b in range(batch)])[:, None].astype(np.float32)
    return torch.from_numpy(patch), torch.from_numpy(labels), torch.from_numpy(mask)


def noise_list(shapes, seed=0):
    """Pre-generated standard-normal tensors, consumed in call order."""
    rs = _rs(seed, 'noise')
    return [torch.from_numpy(rs.standard_normal(s).astype(np.float32)) for s in shapes]


def phiseg_noise_shapes(batch, size=128, latent_levels=5, resolution_levels=7, z_dim=2):
    """Shapes of the randn_like draws of one PHISeg.forward: posterior levels 4..0 then prior 4..0
    (reference models/phiseg.py:104,197-202)."""
    one = []
    for i in range(latent_levels):
        lvl = latent_levels - 1 - i
        r = size >> (lvl + resolution_levels - latent_levels)
        one.append((batch, z_dim, r, r))
    return one + one


def synth_state_dict(template, seed=0):
    """Fill every entry of ``template`` (a state_dict: name -> tensor, used for names/shapes/dtypes)
    with seeded values: conv weights ~ N(0, 2/fan_in) (Kaiming), conv bias ~ 0.05 N, BN gamma ~ 1+0.1 N,
    BN beta ~ 0.1 N, running_mean ~ 0.1 N, running_var ~ U(0.5,1.5), num_batches_tracked = 3."""
    out = {}
    for name, t in template.items():
        rs = _rs(seed, name)
        shape = tuple(t.shape)
        if name.endswith('num_batches_tracked'):
            v = np.asarray(3, np.int64)
        elif name.endswith('running_mean'):
            v = 0.1 * rs.standard_normal(shape)
        elif name.endswith('running_var'):
            v = rs.uniform(0.5, 1.5, shape)
        elif len(shape) >= 4:
            fan_in = int(np.prod(shape[1:]))
            v = rs.standard_normal(shape) * np.sqrt(2.0 / fan_in)
            if 'mu_conv' in name or 'sigma_conv' in name:
                v = v * 0.25                       # keep sigma away from 0 so KL / z stay well conditioned
        elif name.endswith('.1.weight'):          # BatchNorm gamma (Conv2D.convolution.1)
            v = 1.0 + 0.1 * rs.standard_normal(shape)
            if 'reversible_blocks' in name:       # small residual branches: y = x + F(x) stays bounded in eval mode,
                v = 0.2 + 0.02 * rs.standard_normal(shape)   # where the synthetic running statistics do not normalise
        elif name.endswith('.1.bias'):            # BatchNorm beta
            v = 0.1 * rs.standard_normal(shape)
        else:                                     # conv bias
            v = 0.05 * rs.standard_normal(shape)
        out[name] = torch.from_numpy(np.asarray(v)).to(t.dtype).reshape(shape).clone()
    return out


def brats_like_batch(batch, size=32, channels=4, seed=0):
    """BraTS-shaped synthetic volumes: volume fp32 [B,4,S,S,S] ~ N(0,1)*0.5 with the lesion imprinted,
    label index volume float32 [B,1,S,S,S] in {0,1,2} from two nested ellipsoids (reference data/bratsDataset.py:88-89,
    125-131 only fixes these shapes)."""
    rs = _rs(seed, 'brats')
    vol = (rs.standard_normal((batch, channels, size, size, size)) * 0.5).astype(np.float32)
    zz, yy, xx = np.mgrid[0:size, 0:size, 0:size].astype(np.float32)
    lab = np.zeros((batch, 1, size, size, size), np.float32)
    for b in range(batch):
        c = rs.uniform(0.35 * size, 0.65 * size, 3)
        r = rs.uniform(0.15 * size, 0.3 * size, 3)
        d = ((zz - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 + ((xx - c[2]) / r[2]) ** 2
        lab[b, 0] = (d <= 1.0).astype(np.float32) + (d <= 0.3).astype(np.float32)
        vol[b] += 0.3 * lab[b]
    return torch.from_numpy(vol), torch.from_numpy(lab)


def phiseg3d_noise_shapes(batch, size, latent_levels, resolution_levels, z_dim=2):
    """randn_like draws of one PHISeg3D.forward: posterior levels deepest first, then the prior's
    (reference models/phiseg3D.py:188,276-281)."""
    one = []
    for i in range(latent_levels):
        lvl = latent_levels - 1 - i
        r = size >> (lvl + resolution_levels - latent_levels)
        one.append((batch, z_dim, r, r, r))
    return one + one
