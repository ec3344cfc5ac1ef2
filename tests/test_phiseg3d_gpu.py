"""End-to-end parity of the drop-in PHISeg3D (CUDA path through the C ABI) against the CPU oracle with the same storage
rounding, the fp32 oracle (= the patched reference's arithmetic) and the reference-generated fixture
tests/golden/phiseg3d_small.npz (32^3 volumes, filters [32,64,64], 3 latent levels: 32^3 / 16^3 / 8^3 -- the last one
exercises partial 16x16 tiles)."""
import os

import numpy as np
import pytest
import torch

from oracle import phiseg3d_oracle as o3
from oracle import phiseg_oracle as po
from oracle import synth
from oracle.ref_run import injected_noise
from tests.keygrammar import dropin_phiseg3d

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


def _setup(golden_dir, reversible=False):
    g = np.load(os.path.join(golden_dir, 'phiseg3d_small.npz'))
    filters = [int(v) for v in g['filters']]          # [32, 64, 64]: reversible halves of 16 channels at full resolution
    L, size, batch = int(g['latent_levels']), int(g['size']), int(g['batch'])
    net = dropin_phiseg3d(filters, L, (4, size, size, size), reversible=reversible)
    sd = synth.synth_state_dict(net.state_dict(), seed=int(g['wseed']))
    net.load_state_dict(sd)
    net = net.cuda()
    vol, lab = synth.brats_like_batch(batch, size=size, seed=int(g['dseed']))
    eps = synth.noise_list(synth.phiseg3d_noise_shapes(batch, size, L, len(filters)), seed=int(g['nseed']))
    return g, filters, L, net, sd, vol, lab, eps


@pytest.mark.parametrize('training', [True, False])
def test_forward_and_losses(golden_dir, training):
    g, filters, L, net, sd, vol, lab, eps = _setup(golden_dir)
    key = 'train' if training else 'eval'
    net.train(training)
    with injected_noise(eps), torch.no_grad():
        s = [t.clone() for t in net.forward(vol.cuda(), lab.cuda(), training=training)]
        loss = net.loss(lab.cuda())
    assert net.kl_divergence_loss is net.loss_tot and net.reconstruction_loss is net.loss_tot and loss is net.loss_tot
    with torch.no_grad():
        emu = o3.phiseg3d_forward({k: v.clone() for k, v in sd.items()}, vol, lab, eps, L, len(filters), 3,
                                  training=training, rnd=po.Rounding(True))
        e_emu = o3.elbo(emu, lab)
        ref = o3.phiseg3d_forward({k: v.clone() for k, v in sd.items()}, vol, lab, eps, L, len(filters), 3,
                                  training=training)
        e_ref = o3.elbo(ref, lab)
    acc, acc_emu, acc_ref = sum(t.cpu() for t in s), sum(emu['s']), sum(ref['s'])
    rel_emu, rel_ref = _rel(acc, acc_emu), _rel(acc, acc_ref)
    agree = float((acc.argmax(1) == acc_ref.argmax(1)).float().mean())
    print('\n[phiseg3d %s] logits rel-L2 vs bf16-emulating oracle %.3e, vs fp32 oracle %.3e (oracles apart %.3e); argmax '
          'agreement %.5f; loss cuda %.6g emu %.6g fp32 %.6g fixture %.6g'
          % (key, rel_emu, rel_ref, _rel(acc_emu, acc_ref), agree, float(loss), float(e_emu['total']),
             float(e_ref['total']), float(g[key + '_loss'])))
    # same rounding points -> summation order / bf16 ties only; fp32 reference arithmetic -> bf16 storage envelope
    assert rel_emu < 2e-2
    assert rel_ref < 5e-2
    for lvl in range(L):
        assert _rel(net.prior_mu[lvl].cpu(), emu['prior_mu'][lvl]) < 3e-2
        assert _rel(net.posterior_sigma[lvl].cpu(), emu['post_sigma'][lvl]) < 3e-2
        assert float(net.loss_dict['KL_divergence_loss_lvl%d' % lvl]) == pytest.approx(
            float(e_emu['kl_levels'][lvl]), rel=3e-2)
    assert float(loss) == pytest.approx(float(e_emu['total']), rel=1e-2)
    assert float(loss) == pytest.approx(float(g[key + '_loss']), rel=2e-2)
    assert agree > 0.98
    st_full = sum(t.cpu() for t in s)[:, :, ::2, ::2, ::2]
    assert _rel(st_full, torch.from_numpy(g[key + '_logits_ds2'])) < 5e-2


def _oracle_grads(sd, vol, lab, eps, L, R, bf16):
    sd2 = {k: v.clone() for k, v in sd.items()}
    params = {k: v.requires_grad_(True) for k, v in sd2.items() if v.dtype == torch.float32 and 'running_' not in k}
    out = o3.phiseg3d_forward(sd2, vol, lab, eps, L, R, 3, training=True, rnd=po.Rounding(bf16))
    o3.elbo(out, lab)['total'].backward()
    return params, sd2


@pytest.mark.parametrize('reversible', [False, True])
def test_training_step_gradients(golden_dir, reversible):
    g, filters, L, net, sd, vol, lab, eps = _setup(golden_dir, reversible)
    net.train(True)
    with injected_noise(eps):
        net.forward(vol.cuda(), lab.cuda(), training=True)
        loss = net.loss(lab.cuda())
    loss.backward()
    torch.cuda.synchronize()
    p_fp32, _ = _oracle_grads(sd, vol, lab, eps, L, len(filters), False)
    p_emu, sd_after = _oracle_grads(sd, vol, lab, eps, L, len(filters), True)
    named = dict(net.named_parameters())
    gmax = max(float(p.grad.norm()) for p in p_fp32.values() if p.grad is not None)
    e_emu, e_fp32, o_gap = [], [], []
    for n, p in p_fp32.items():
        if p.grad is None:
            assert named[n].grad is None, n
            continue
        got = named[n].grad.cpu()
        if n.endswith('convolution.0.bias') and (n[:-len('0.bias')] + '1.weight') in p_fp32:
            assert float(got.abs().max()) == 0.0     # conv bias in front of BatchNorm3d: exactly zero
            continue
        if float(p.grad.norm()) < 1e-6 * gmax:
            continue
        e_emu.append((_rel(got, p_emu[n].grad), n))
        e_fp32.append(_rel(got, p.grad))
        o_gap.append(_rel(p_emu[n].grad, p.grad))
    e_emu.sort(reverse=True)
    med_emu, med_fp32, med_gap = (float(np.median([w for w, _ in e_emu])), float(np.median(e_fp32)),
                                  float(np.median(o_gap)))
    print('\n[phiseg3d rev=%s] parameter-gradient rel-L2: cuda vs same-rounding oracle median %.3e (worst %s); cuda vs '
          'fp32 oracle median %.3e; oracles apart median %.3e' % (reversible, med_emu, e_emu[:3], med_fp32, med_gap))
    assert med_emu < 0.2
    assert med_fp32 < 1.25 * med_gap + 0.02
    # BatchNorm3d running statistics after the step (two momentum updates inside reversible blocks, quirk Q7)
    after = net.state_dict()
    worst = max(_rel(after[k].cpu(), sd_after[k]) for k in after if k.endswith('running_var'))
    assert worst < 2e-2
    if not reversible:
        k = str(g['train_running_var_probe_key'])
        np.testing.assert_allclose(after[k].cpu().numpy(), g['train_running_var_probe'], rtol=5e-3)


def test_sample_and_accumulate_volume(golden_dir):
    g, filters, L, net, sd, vol, lab, eps = _setup(golden_dir)
    net.eval()
    with torch.no_grad():
        s = net.forward(vol.cuda(), lab.cuda(), training=False)
        want = sum(t.clone() for t in s)
        acc = net.accumulate_output(s, use_softmax=False)
        assert acc.data_ptr() == s[-1].data_ptr()                      # in place like the reference (quirk Q2)
        torch.testing.assert_close(acc, want, rtol=1e-5, atol=1e-5)
        sample = net.sample(testing=True)
        assert tuple(sample.shape) == (vol.shape[0], 3) + tuple(vol.shape[2:])
