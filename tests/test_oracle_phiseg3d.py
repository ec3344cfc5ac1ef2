"""CPU: the PHISeg3D restatement (oracle/phiseg3d_oracle.py) against the reference-generated fixture and, when
/root/reference is present, against the live reference module under the three documented patches;
state_dict key grammar of the 3-D drop-in."""
import os

import numpy as np
import pytest
import torch

from oracle import phiseg3d_oracle as o3
from oracle import synth
from oracle.ref_loader import have_reference
from tests.keygrammar import dropin_phiseg3d


CASES = ['phiseg3d_small', 'phiseg3d_rev_small']


def _case(golden_dir, case='phiseg3d_small'):
    g = np.load(os.path.join(golden_dir, case + '.npz'))
    filters = [int(v) for v in g['filters']]
    L, size, batch = int(g['latent_levels']), int(g['size']), int(g['batch'])
    net = dropin_phiseg3d(filters, L, (4, size, size, size), reversible=bool(int(g['reversible'])))
    sd = synth.synth_state_dict(net.state_dict(), seed=int(g['wseed']))
    vol, lab = synth.brats_like_batch(batch, size=size, seed=int(g['dseed']))
    eps = synth.noise_list(synth.phiseg3d_noise_shapes(batch, size, L, len(filters)), seed=int(g['nseed']))
    return g, filters, L, sd, vol, lab, eps


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('training', [True, False])
def test_oracle_matches_reference_fixture(golden_dir, training, case):
    g, filters, L, sd, vol, lab, eps = _case(golden_dir, case)
    key = 'train' if training else 'eval'
    with torch.no_grad():
        out = o3.phiseg3d_forward({k: v.clone() for k, v in sd.items()}, vol, lab, eps, L, len(filters), 3,
                                  training=training)
        e = o3.elbo(out, lab)
    assert float(e['total']) == pytest.approx(float(g[key + '_loss']), rel=1e-5)
    for lvl in range(L):
        assert float(e['kl_levels'][lvl]) == pytest.approx(float(g['%s_KL_divergence_loss_lvl%d' % (key, lvl)]), rel=1e-4)
        assert float(e['ce_levels'][lvl]) == pytest.approx(float(g['%s_residual_multinoulli_loss_lvl%d' % (key, lvl)]),
                                                           rel=1e-5)
        st = 2 if lvl == 0 else 1
        np.testing.assert_allclose(out['post_mu'][lvl][:, :, ::st, ::st, ::st].numpy(), g['%s_post_mu%d' % (key, lvl)],
                                   rtol=1e-3, atol=2e-5)
        np.testing.assert_allclose(out['prior_sigma'][lvl][:, :, ::st, ::st, ::st].numpy(),
                                   g['%s_prior_sigma%d' % (key, lvl)], rtol=1e-3, atol=2e-5)
    acc = sum(out['s'])
    np.testing.assert_allclose(acc[:, :, ::2, ::2, ::2].numpy(), g[key + '_logits_ds2'], rtol=1e-3, atol=2e-4)


@pytest.mark.parametrize('case', CASES)
def test_oracle_gradients_match_reference_fixture(golden_dir, case):
    g, filters, L, sd, vol, lab, eps = _case(golden_dir, case)
    sd2 = {k: v.clone() for k, v in sd.items()}
    params = {k: v.requires_grad_(True) for k, v in sd2.items() if v.dtype == torch.float32 and 'running_' not in k}
    out = o3.phiseg3d_forward(sd2, vol, lab, eps, L, len(filters), 3, training=True)
    o3.elbo(out, lab)['total'].backward()
    names = [str(n) for n in g['train_grad_names']]
    nograd = set(str(n) for n in g['train_nograd_names'])
    gmax = float(np.max(g['train_grad_norms']))
    for n, ref in zip(names, g['train_grad_norms']):
        got = float(params[n].grad.norm())
        assert got == pytest.approx(float(ref), rel=2e-3, abs=1e-5 * gmax), n
    for n in nograd:
        assert params[n].grad is None or float(params[n].grad.abs().max()) == 0.0, n
    # BatchNorm3d running statistics: one momentum-0.01 update
    k = str(g['train_running_var_probe_key'])
    np.testing.assert_allclose(sd2[k].numpy(), g['train_running_var_probe'], rtol=1e-5)


def test_dropin_state_dict_keys_match_fixture_and_reference(golden_dir):
    g, filters, L, sd, vol, lab, eps = _case(golden_dir)
    params = set(k for k in sd if 'running_' not in k and 'num_batches' not in k)
    assert params == set(str(n) for n in g['train_grad_names']) | set(str(n) for n in g['train_nograd_names'])
    if have_reference():
        from oracle.ref_run import build_reference_phiseg3d
        for rev, f in ((False, filters), (True, [32, 64, 128])):
            ref = build_reference_phiseg3d(f, (4, 32, 32, 32), L, reversible=rev).state_dict()
            mine = dropin_phiseg3d(f, L, (4, 32, 32, 32), reversible=rev).state_dict()
            assert list(ref.keys()) == list(mine.keys())
            assert all(tuple(ref[k].shape) == tuple(mine[k].shape) and ref[k].dtype == mine[k].dtype for k in ref)


@pytest.mark.skipif(not have_reference(), reason='reference checkout not present (GPU box)')
@pytest.mark.parametrize('reversible', [False, True])
def test_oracle_matches_live_patched_reference(reversible):
    from oracle.ref_run import build_reference_phiseg3d, injected_noise, phiseg3d_patches
    filters, L, size, batch = ([64, 64, 64] if reversible else [32, 64, 64]), 3, 16, 2
    net = build_reference_phiseg3d(filters, (4, size, size, size), L, reversible=reversible)
    sd = synth.synth_state_dict(net.state_dict(), seed=7)
    net.load_state_dict(sd)
    vol, lab = synth.brats_like_batch(batch, size=size, seed=2)
    eps = synth.noise_list(synth.phiseg3d_noise_shapes(batch, size, L, 3), seed=9)
    net.train()
    with phiseg3d_patches(net), injected_noise(eps):
        s = [t.clone() for t in net.forward(vol, lab, training=True)]
        loss = net.loss(lab)
    loss.backward()
    sd2 = {k: v.clone() for k, v in sd.items()}
    params = {k: v.requires_grad_(True) for k, v in sd2.items() if v.dtype == torch.float32 and 'running_' not in k}
    out = o3.phiseg3d_forward(sd2, vol, lab, eps, L, 3, 3, training=True)
    e = o3.elbo(out, lab)
    e['total'].backward()
    assert float(e['total']) == pytest.approx(float(loss), rel=1e-5)
    for lvl in range(L):
        torch.testing.assert_close(out['s'][lvl], s[lvl], rtol=1e-3, atol=1e-4)
    ref_p = dict(net.named_parameters())
    gmax = max(float(p.grad.norm()) for p in ref_p.values() if p.grad is not None)
    for n, p in ref_p.items():
        if p.grad is None:
            continue
        assert float(params[n].grad.norm()) == pytest.approx(float(p.grad.norm()), rel=5e-3, abs=1e-5 * gmax), n
    # running statistics after the step (reversible blocks: two updates)
    after = net.state_dict()
    for k in after:
        if k.endswith('running_var'):
            torch.testing.assert_close(sd2[k], after[k], rtol=1e-4, atol=1e-6)
