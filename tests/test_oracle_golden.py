"""The CPU oracle (oracle/phiseg_oracle.py) against fixtures produced by the REAL reference
(oracle/make_golden.py) and, when /root/reference is present, against the reference run live."""
import os

import numpy as np
import pytest
import torch

from oracle import phiseg_oracle as po
from oracle import synth
from oracle.ref_loader import have_reference


def _template(filters, reversible=False):
    """state_dict names/shapes of the reference PHISeg without importing it: built by the drop-in's
    key grammar helper (pure host logic)."""
    from tests.keygrammar import phiseg_state_template
    return phiseg_state_template(filters, reversible=reversible)


def _run_case(g, training):
    filters = [int(v) for v in g['filters']]
    batch = int(g['batch'])
    rev = bool(int(g['reversible'])) if 'reversible' in g else False
    sd = synth.synth_state_dict(_template(filters, rev), seed=int(g['wseed']))
    patch, labels, mask = synth.lidc_like_batch(batch, seed=int(g['dseed']))
    eps = synth.noise_list(synth.phiseg_noise_shapes(batch), seed=int(g['nseed']))
    out = po.phiseg_forward(sd, patch, mask, eps, training=training)
    return out, po.elbo(out, mask), sd, mask


@pytest.mark.parametrize('case', ['phiseg_small', 'phiseg_lidc', 'phiseg_rev_small'])
@pytest.mark.parametrize('training', [True, False])
def test_oracle_matches_reference_fixture(golden_dir, case, training):
    g = np.load(os.path.join(golden_dir, case + '.npz'))
    key = 'train' if training else 'eval'
    with torch.no_grad():
        out, e, sd, mask = _run_case(g, training)
    assert float(e['total']) == pytest.approx(float(g[key + '_loss']), rel=2e-5)
    for lvl in range(5):
        assert float(e['kl_levels'][lvl]) == pytest.approx(float(g['%s_KL_divergence_loss_lvl%d' % (key, lvl)]), rel=1e-4)
        assert float(e['ce_levels'][lvl]) == pytest.approx(float(g['%s_residual_multinoulli_loss_lvl%d' % (key, lvl)]), rel=1e-4)
        for who, name in (('post', 'post'), ('prior', 'prior')):
            np.testing.assert_allclose(out[who + '_mu'][lvl].numpy(), g['%s_%s_mu%d' % (key, name, lvl)], rtol=1e-3, atol=2e-4)
            np.testing.assert_allclose(out[who + '_sigma'][lvl].numpy(), g['%s_%s_sigma%d' % (key, name, lvl)], rtol=1e-3, atol=2e-4)
    acc = po.accumulate_output(out['s']).numpy()
    if key + '_logits' in g:
        np.testing.assert_allclose(acc, g[key + '_logits'], rtol=1e-3, atol=1e-3)
    else:
        np.testing.assert_allclose(acc[:, :, ::8, ::8], g[key + '_logits_ds8'], rtol=1e-3, atol=1e-3)


def test_oracle_gradients_and_running_stats_match_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, 'phiseg_small.npz'))
    filters = [int(v) for v in g['filters']]
    batch = int(g['batch'])
    sd = synth.synth_state_dict(_template(filters), seed=int(g['wseed']))
    params = {k: v.requires_grad_(True) for k, v in sd.items()
              if v.dtype == torch.float32 and 'running_' not in k}
    patch, labels, mask = synth.lidc_like_batch(batch, seed=int(g['dseed']))
    eps = synth.noise_list(synth.phiseg_noise_shapes(batch), seed=int(g['nseed']))
    out = po.phiseg_forward(sd, patch, mask, eps, training=True)
    po.elbo(out, mask)['total'].backward()
    names = [str(n) for n in g['train_grad_names']]
    norms = g['train_grad_norms']
    nograd = set(str(n) for n in g['train_nograd_names'])
    # conv biases in front of a BatchNorm have a mathematically zero gradient (pure rounding noise in both paths)
    floor = 1e-5 * float(norms.max())
    for n, ref in zip(names, norms):
        got = float(params[n].grad.norm())
        assert got == pytest.approx(float(ref), rel=5e-3, abs=floor), n
    for n in nograd:                      # SURVEY.md 8e caveat (3): upsampling_path.4.* never gets a gradient
        assert params[n].grad is None, n
    k = 'posterior.contracting_path.3.layers.2.convolution.1.running_var'
    np.testing.assert_allclose(sd[k].detach().numpy(), g['train_running_var_probe'], rtol=1e-5)


def test_kl_known_answer(golden_dir):
    """proves the sigma1*sigma0 quirk (textbook KL would be 2.84265)."""
    g = np.load(os.path.join(golden_dir, 'metrics.npz'))
    kl = po.kl_two_gauss(torch.tensor([[[[0.5, -1.0]]]]), torch.tensor([[[[0.8, 1.5]]]]),
                         torch.tensor([[[[0.0, 0.25]]]]), torch.tensor([[[[1.2, 0.7]]]]))
    assert float(kl) == pytest.approx(1.1006804704666138, rel=1e-6)
    assert float(kl) == pytest.approx(float(g['kl1']), rel=1e-6)


def test_bilinear_known_answers():
    """2x2 ramp, align_corners=True first row [0, 1/3, 2/3, 1]."""
    x = torch.tensor([[[[0., 1.], [2., 3.]]]])
    np.testing.assert_allclose(po.up2(x)[0, 0, 0].numpy(), [0, 1 / 3, 2 / 3, 1], rtol=1e-6)


@pytest.mark.skipif(not have_reference(), reason='/root/reference only exists in the build container')
@pytest.mark.parametrize('training', [True, False])
def test_oracle_matches_live_reference(training):
    from oracle.ref_run import build_reference_phiseg, injected_noise
    filters = [16, 32, 48, 32, 48, 16, 48]
    net = build_reference_phiseg(filters)
    sd = synth.synth_state_dict(net.state_dict(), seed=9)
    net.load_state_dict(sd)
    net.train(training)
    patch, labels, mask = synth.lidc_like_batch(3, seed=8)
    eps = synth.noise_list(synth.phiseg_noise_shapes(3), seed=7)
    with injected_noise(eps), torch.no_grad():
        s = [t.clone() for t in net.forward(patch, mask, training=training)]
        loss = net.loss(mask)
    sd2 = {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        out = po.phiseg_forward(sd2, patch, mask, eps, training=training)
        e = po.elbo(out, mask)
    assert float(e['total']) == pytest.approx(float(loss), rel=2e-5)
    for a, b in zip(s, out['s']):
        np.testing.assert_allclose(b.numpy(), a.numpy(), rtol=1e-3, atol=1e-3)
    # the template grammar used on the GPU box must equal the real key set
    from tests.keygrammar import phiseg_state_template
    tpl = phiseg_state_template(filters)
    ref = net.state_dict()
    assert list(tpl.keys()) == list(ref.keys())
    for k in ref:
        assert tuple(tpl[k].shape) == tuple(ref[k].shape) and tpl[k].dtype == ref[k].dtype, k


def test_reversible_oracle_gradients_and_double_running_stat_update(golden_dir):
    """RevPHiSeg (reference torchlayers.py:55-82 over revtorch, restated in oracle/shims/revtorch -- PARITY UNPINNED for
    that third-party package): gradients through the plain forward equal the inverse-recompute backward; BatchNorm
    running statistics of F / G get two momentum updates per training step (quirk Q7)."""
    g = np.load(os.path.join(golden_dir, 'phiseg_rev_small.npz'))
    filters = [int(v) for v in g['filters']]
    batch = int(g['batch'])
    sd = synth.synth_state_dict(_template(filters, True), seed=int(g['wseed']))
    params = {k: v.requires_grad_(True) for k, v in sd.items() if v.dtype == torch.float32 and 'running_' not in k}
    patch, labels, mask = synth.lidc_like_batch(batch, seed=int(g['dseed']))
    eps = synth.noise_list(synth.phiseg_noise_shapes(batch), seed=int(g['nseed']))
    out = po.phiseg_forward(sd, patch, mask, eps, training=True)
    po.elbo(out, mask)['total'].backward()
    norms = g['train_grad_norms']
    floor = 1e-5 * float(norms.max())
    for n, ref in zip(g['train_grad_names'], norms):
        assert float(params[str(n)].grad.norm()) == pytest.approx(float(ref), rel=1e-2, abs=floor), n
    k = str(g['train_running_var_probe_key'])
    np.testing.assert_allclose(sd[k].detach().numpy(), g['train_running_var_probe'], rtol=1e-5)
    assert int(sd[k.replace('running_var', 'num_batches_tracked')]) == int(g['train_num_batches_tracked_probe']) == 5


@pytest.mark.skipif(not have_reference(), reason='/root/reference only exists in the build container')
@pytest.mark.parametrize('reversible', [False, True])
def test_checkpoint_round_trip_with_the_reference(tmp_path, reversible):
    """a `.pth` written by the reference (`torch.save(net.state_dict())`, train_model.py:558-564) loads
    into the drop-in module with strict=True, and a checkpoint written by the drop-in loads back into the reference --
    same keys, shapes, dtypes and values in both directions (host-side: no kernel involved)."""
    from oracle.ref_run import build_reference_phiseg
    from tests.keygrammar import dropin_phiseg
    filters = [32, 64, 96, 32, 64, 32, 64]      # reversible halves must be multiples of 16 on the B200 path
    ref = build_reference_phiseg(filters, reversible=reversible)
    ref.load_state_dict(synth.synth_state_dict(ref.state_dict(), seed=21))
    path = tmp_path / 'reference.pth'
    torch.save(ref.state_dict(), path)
    ours = dropin_phiseg(filters, reversible=reversible)
    missing, unexpected = ours.load_state_dict(torch.load(path), strict=True)
    assert not missing and not unexpected
    for (k, v), (k2, v2) in zip(ref.state_dict().items(), ours.state_dict().items()):
        assert k == k2 and v.dtype == v2.dtype and torch.equal(v, v2), k
    back = tmp_path / 'dropin.pth'
    torch.save(ours.state_dict(), back)
    ref2 = build_reference_phiseg(filters, reversible=reversible)
    ref2.load_state_dict(torch.load(back), strict=True)
    for (k, v), (_, v2) in zip(ref.state_dict().items(), ref2.state_dict().items()):
        assert torch.equal(v, v2), k
