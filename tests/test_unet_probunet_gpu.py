"""GPU parity of the drop-in U-Net and Probabilistic U-Net (CUDA path) against the CPU oracle (same bf16 storage
rounding: tight; fp32 reference arithmetic and reference-generated fixtures: bf16 envelope)."""
import os

import numpy as np
import pytest
import torch
import torch.distributions.normal as tdn

from oracle import phiseg_oracle as po
from oracle import synth
from oracle import unet_oracle as uo
from tests.gpu_util import PKG  # noqa: F401

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


def test_unet_forward_loss_backward(golden_dir):
    from models.unet import Unet
    g = np.load(os.path.join(golden_dir, 'unet_probunet.npz'))
    filters = [int(v) for v in g['unet_filters']]
    net = Unet(1, 2, filters)
    sd = synth.synth_state_dict(net.state_dict(), seed=2)
    net.load_state_dict(sd)
    net = net.cuda()
    patch, labels, mask = synth.lidc_like_batch(3, seed=4)
    logits = net.forward(patch.cuda())
    loss = net.loss(mask.cuda())
    assert net.sample() is net.prediction
    loss.backward()
    outs = {}
    for tag, bf16 in (('emu', True), ('fp32', False)):
        sd2 = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        lg = uo.unet_forward(sd2, patch, len(filters), rnd=po.Rounding(bf16))
        ls = uo.unet_loss(lg, mask)
        ls.backward()
        outs[tag] = (lg.detach(), float(ls), sd2)
    print('\nU-Net logits rel-L2 vs same-rounding oracle %.3e, vs fp32 oracle %.3e; loss cuda %.6f emu %.6f fp32 %.6f '
          'golden %.6f' % (_rel(logits.detach().cpu(), outs['emu'][0]), _rel(logits.detach().cpu(), outs['fp32'][0]),
                           float(loss), outs['emu'][1], outs['fp32'][1], float(g['unet_loss'])))
    assert _rel(logits.detach().cpu(), outs['emu'][0]) < 1e-2
    assert _rel(logits.detach().cpu(), outs['fp32'][0]) < 2e-2
    assert float(loss) == pytest.approx(float(g['unet_loss']), rel=5e-3)       # reference fixture
    np.testing.assert_allclose(logits.detach().cpu()[:, :, ::4, ::4].numpy(), g['unet_logits_ds4'], rtol=0.1, atol=0.05)
    named = dict(net.named_parameters())
    errs = sorted(((_rel(named[n].grad.cpu(), p.grad), n) for n, p in outs['emu'][2].items()), reverse=True)
    med = float(np.median([e for e, _ in errs]))
    print('U-Net gradient rel-L2 vs same-rounding oracle: median %.3e worst %s' % (med, errs[:3]))
    assert med < 3e-2 and errs[0][0] < 0.2


@pytest.mark.parametrize('training', [True, False])
def test_probunet_step(golden_dir, training):
    from models.probabilistic_unet import ProbabilisticUnet
    g = np.load(os.path.join(golden_dir, 'unet_probunet.npz'))
    key = 'train' if training else 'eval'
    filters = [int(v) for v in g['prob_filters']]
    net = ProbabilisticUnet(input_channels=1, num_classes=2, num_filters=filters, latent_dim=6, no_convs_fcomb=3)
    sd = synth.synth_state_dict(net.state_dict(), seed=3)
    net.load_state_dict(sd)
    net = net.cuda().train(training)
    patch, labels, mask = synth.lidc_like_batch(3, seed=4)
    eps = synth.noise_list([(3, 6)], seed=8)[0]
    orig = tdn._standard_normal
    tdn._standard_normal = lambda shape, dtype, device: eps.to(device)
    try:
        fwd = net.forward(patch.cuda(), mask.cuda(), training=training)
        loss = net.loss(mask.cuda())
    finally:
        tdn._standard_normal = orig
    with torch.no_grad():
        emu = uo.probunet_step({k: v.clone() for k, v in sd.items()}, patch, mask, eps, 7, 6, 3, training=training,
                               rnd=po.Rounding(True))
    print('\nProbUNet[%s] loss cuda %.6g emu %.6g golden %.6g | kl cuda %.5g emu %.5g golden %.5g | forward rel-L2 vs emu '
          '%.3e, reconstruction rel-L2 vs emu %.3e' %
          (key, float(loss), float(emu['loss']), float(g['prob_%s_loss' % key]), float(net.kl_divergence_loss),
           float(emu['kl']), float(g['prob_%s_kl' % key]), _rel(fwd.detach().cpu(), emu['forward']),
           _rel(net.reconstruction.detach().cpu(), emu['reconstruction'])))
    assert _rel(fwd.detach().cpu(), emu['forward']) < 2e-2
    assert _rel(net.reconstruction.detach().cpu(), emu['reconstruction']) < (8e-2 if training else 3e-2)
    assert float(loss) == pytest.approx(float(emu['loss']), rel=2e-2)
    assert float(loss) == pytest.approx(float(g['prob_%s_loss' % key]), rel=3e-2)
    assert float(net.reconstruction_loss) == pytest.approx(float(g['prob_%s_rec' % key]), rel=3e-2)
    assert float(net.kl_divergence_loss) == pytest.approx(float(g['prob_%s_kl' % key]), rel=0.1, abs=0.05)
    np.testing.assert_allclose(net.posterior_latent_space.mean.detach().cpu().numpy(), g['prob_%s_mu_q' % key],
                               rtol=0.1, atol=0.05)
    # API surface: distributions, sample(), accumulate_output
    assert tuple(net.prior_latent_space.rsample().shape) == (3, 6)
    smp = net.sample(testing=True)
    assert tuple(smp.shape) == (3, 2, 128, 128)
    probs = net.accumulate_output(smp, use_softmax=True)
    torch.testing.assert_close(probs, torch.softmax(smp, 1), rtol=1e-5, atol=1e-6)
    if training:
        loss.backward()
        nograd = set(str(n) for n in g['prob_nograd_names'])
        for n, p in net.named_parameters():
            assert (p.grad is None) == (n in nograd), n       # last_conv.* gets no gradient
        gd = dict(zip((str(n) for n in g['prob_grad_names']), g['prob_grad_norms']))
        rel = []
        for n, p in net.named_parameters():
            if p.grad is None or gd[n] < 1e-6 * g['prob_grad_norms'].max():
                continue
            if n.endswith('convolution.0.bias') and (n[:-len('0.bias')] + '1.weight') in gd:
                continue                                       # conv bias in front of BatchNorm: exactly zero here
            rel.append(abs(float(p.grad.norm()) - gd[n]) / gd[n])
        print('ProbUNet gradient-norm relative deviation from the reference fixture: median %.3e max %.3e' %
              (float(np.median(rel)), max(rel)))
        assert float(np.median(rel)) < 0.15


def test_unet_rgb_non_square_input():
    """3-channel 96 x 160 input, batch 1: generic tile geometry of the small-shape conv kernel, padded input channels."""
    from models.unet import Unet
    filters = [16, 32, 32, 48]
    net = Unet(3, 2, filters)
    sd = synth.synth_state_dict(net.state_dict(), seed=9)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    rs = np.random.RandomState(3)
    x = torch.from_numpy((rs.standard_normal((1, 3, 96, 160)) * 0.3).astype(np.float32))
    with torch.no_grad():
        got = net.forward(x.cuda()).cpu()
        ref = uo.unet_forward(sd, x, len(filters), rnd=po.Rounding(True))
    assert tuple(got.shape) == (1, 2, 96, 160)
    rel = float((got - ref).norm() / ref.norm())
    assert rel < 3e-2, rel
