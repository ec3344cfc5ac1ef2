"""Parity at the CONFIGURED sizes of BASELINE.json: U-Net [32,64,128,192] batch 12, ProbUNet
7 x filters latent 6 batch 12, PHISeg3D [32,64,128] / 3 latent levels on a 64^3 BraTS-shaped volume (128^3 in the
benchmark; 64^3 keeps the fp32 oracle at a few seconds), each against the oracle executed in fp32 on the same GPU
(TF32 off; the oracle is pinned to the real reference by the CPU tests and the fixtures of tests/golden/).
Training mode (batch statistics), random synthetic weights; tolerances = measured level of the bf16 path with margin,
next to the same-rounding oracle's own distance from fp32."""
import numpy as np
import pytest
import torch
import torch.distributions.normal as tdn

from oracle import phiseg3d_oracle as o3
from oracle import phiseg_oracle as po
from oracle import synth
from oracle import unet_oracle as uo
from oracle.ref_run import injected_noise
from tests.gpu_util import PKG  # noqa: F401

pytestmark = pytest.mark.gpu
FILTERS = [32, 64, 128, 192, 192, 192, 192]


def _fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def _cuda(sd):
    return {k: v.cuda() for k, v in sd.items()}


def test_unet_baseline_size():
    """BASELINE configs[0]: vanilla U-Net [32,64,128,192], batch 12, 128^2, forward + loss + backward"""
    from b200 import build
    _fp32()
    filters = [32, 64, 128, 192]
    net = build.unet(filters)
    sd = synth.synth_state_dict(net.state_dict(), seed=2)
    net.load_state_dict(sd)
    net = net.cuda()
    patch, _, mask = synth.lidc_like_batch(12, seed=4)
    logits = net.forward(patch.cuda())
    loss = net.loss(mask.cuda())
    loss.backward()
    res = {}
    for tag, bf16 in (('emu', True), ('fp32', False)):
        sd2 = {k: v.clone().requires_grad_(True) for k, v in _cuda(sd).items()}
        lg = uo.unet_forward(sd2, patch.cuda(), len(filters), rnd=po.Rounding(bf16))
        ls = uo.unet_loss(lg, mask.cuda())
        ls.backward()
        res[tag] = (lg.detach(), float(ls), sd2)
    named = dict(net.named_parameters())
    gerr = [_rel(named[n].grad, p.grad) for n, p in res['fp32'][2].items() if p.grad is not None]
    agree = float((logits.argmax(1) == res['fp32'][0].argmax(1)).float().mean())
    print('\nU-Net B=12: logits rel-L2 vs fp32 %.3e (same-rounding oracle vs fp32 %.3e), loss %.6f vs %.6f, argmax agreement '
          '%.5f, gradient rel-L2 median %.3e' % (_rel(logits.detach(), res['fp32'][0]), _rel(res['emu'][0], res['fp32'][0]),
                                                 float(loss), res['fp32'][1], agree, float(np.median(gerr))))
    assert _rel(logits.detach(), res['fp32'][0]) < 1e-2
    assert abs(float(loss) - res['fp32'][1]) / abs(res['fp32'][1]) < 1e-3
    assert agree > 0.995
    assert float(np.median(gerr)) < 3e-2


def test_probunet_baseline_size():
    """BASELINE configs[1]: ProbUNet, 7 filters, latent dim 6, batch 12, training step (posterior / prior sampling + KL)"""
    from b200 import build
    _fp32()
    net = build.probunet(FILTERS, latent_dim=6)
    sd = synth.synth_state_dict(net.state_dict(), seed=3)
    net.load_state_dict(sd)
    net = net.cuda().train(True)
    patch, _, mask = synth.lidc_like_batch(12, seed=4)
    eps = synth.noise_list([(12, 6)], seed=8)[0]
    orig = tdn._standard_normal
    tdn._standard_normal = lambda shape, dtype, device: eps.to(device)
    try:
        net.forward(patch.cuda(), mask.cuda(), training=True)
        loss = net.loss(mask.cuda())
    finally:
        tdn._standard_normal = orig
    with torch.no_grad():
        ref = uo.probunet_step({k: v.clone() for k, v in _cuda(sd).items()}, patch.cuda(), mask.cuda(), eps.cuda(), 7, 6, 3,
                               training=True)
        emu = uo.probunet_step({k: v.clone() for k, v in _cuda(sd).items()}, patch.cuda(), mask.cuda(), eps.cuda(), 7, 6, 3,
                               training=True, rnd=po.Rounding(True))
    r_rec, r_gap = _rel(net.reconstruction.detach(), ref['reconstruction']), _rel(emu['reconstruction'], ref['reconstruction'])
    l_err = abs(float(loss) - float(ref['loss'])) / abs(float(ref['loss']))
    print('\nProbUNet B=12: reconstruction logits rel-L2 vs fp32 %.3e (same-rounding oracle vs fp32 %.3e), loss %.6g vs %.6g '
          '(rel %.2e), KL %.5g vs %.5g' % (r_rec, r_gap, float(loss), float(ref['loss']), l_err,
                                           float(net.kl_divergence_loss), float(ref['kl'])))
    assert r_rec < 2.0 * r_gap + 5e-3
    # random-weight fixture: the yardstick is the distance of the same-rounding oracle from the fp32 one
    l_gap = abs(float(emu['loss']) - float(ref['loss'])) / abs(float(ref['loss']))
    assert l_err < 2.0 * l_gap + 5e-3


@pytest.mark.parametrize('reversible', [False, True])
def test_phiseg3d_64cubed(reversible):
    """BASELINE configs[4] shapes ([32,64,128], 3 latent levels, 4 input channels, 3 classes, batch 1) on a 64^3 volume"""
    from b200 import build
    _fp32()
    filters, L, size = [32, 64, 128], 3, 64
    net = build.phiseg3d(filters, L, (4, size, size, size), reversible=reversible)
    sd = synth.synth_state_dict(net.state_dict(), seed=5)
    net.load_state_dict(sd)
    net = net.cuda().train(True)
    vol, lab = synth.brats_like_batch(1, size=size, seed=6)
    eps = synth.noise_list(synth.phiseg3d_noise_shapes(1, size, L, len(filters)), seed=7)
    with injected_noise(eps), torch.no_grad():
        s = [t.clone() for t in net.forward(vol.cuda(), lab.cuda(), training=True)]
        loss = float(net.loss(lab.cuda()))
    with torch.no_grad():
        ref = o3.phiseg3d_forward({k: v.clone() for k, v in _cuda(sd).items()}, vol.cuda(), lab.cuda(),
                                  [e.cuda() for e in eps], L, len(filters), 3, training=True)
        e_ref = float(o3.elbo(ref, lab.cuda())['total'])
        emu = o3.phiseg3d_forward({k: v.clone() for k, v in _cuda(sd).items()}, vol.cuda(), lab.cuda(),
                                  [e.cuda() for e in eps], L, len(filters), 3, training=True, rnd=po.Rounding(True))
    acc, acc_ref, acc_emu = sum(s), sum(ref['s']), sum(emu['s'])
    agree = float((acc.argmax(1) == acc_ref.argmax(1)).float().mean())
    print('\nPHISeg3D%s 64^3: logits rel-L2 vs fp32 %.3e (same-rounding oracle vs fp32 %.3e), ELBO %.6g vs %.6g (rel %.2e), '
          'argmax agreement %.5f' % (' (reversible)' if reversible else '', _rel(acc, acc_ref), _rel(acc_emu, acc_ref), loss,
                                     e_ref, abs(loss - e_ref) / abs(e_ref), agree))
    assert _rel(acc, acc_ref) < 2.0 * _rel(acc_emu, acc_ref) + 1e-2
    assert abs(loss - e_ref) / abs(e_ref) < 2e-2
    assert agree > 0.98
