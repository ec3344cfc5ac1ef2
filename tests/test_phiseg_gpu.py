"""End-to-end parity of the drop-in PHISeg (CUDA path through the C ABI) against
  (a) the CPU oracle with the SAME storage rounding (bf16 activations / conv weights, fp32 heads and losses): tight;
  (b) the fp32 oracle == the reference's arithmetic, and the reference-generated golden fixtures: at the tolerance the
      bf16 storage allows (stated per assert).
Weights, inputs and noise are re-synthesised from the fixture seeds (oracle/synth.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import phiseg_oracle as po
from oracle import synth
from oracle.ref_run import injected_noise
from tests.keygrammar import dropin_phiseg

pytestmark = pytest.mark.gpu


def _setup(case, golden_dir):
    g = np.load(os.path.join(golden_dir, case + '.npz'))
    filters = [int(v) for v in g['filters']]
    batch = int(g['batch'])
    net = dropin_phiseg(filters)
    sd = synth.synth_state_dict(net.state_dict(), seed=int(g['wseed']))
    net.load_state_dict(sd)
    net = net.cuda()
    patch, labels, mask = synth.lidc_like_batch(batch, seed=int(g['dseed']))
    eps = synth.noise_list(synth.phiseg_noise_shapes(batch), seed=int(g['nseed']))
    return g, net, sd, patch, mask, eps


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.mark.parametrize('case', ['phiseg_small', 'phiseg_lidc'])
@pytest.mark.parametrize('training', [True, False])
def test_forward_and_losses(golden_dir, case, training):
    g, net, sd, patch, mask, eps = _setup(case, golden_dir)
    key = 'train' if training else 'eval'
    net.train(training)
    with injected_noise(eps), torch.no_grad():
        s = net.forward(patch.cuda(), mask.cuda(), training=training)
        s = [t.clone() for t in s]
        loss = net.loss(mask.cuda())
    # quirk Q1: the three loss handles are one tensor
    assert net.kl_divergence_loss is net.loss_tot and net.reconstruction_loss is net.loss_tot and loss is net.loss_tot
    with torch.no_grad():
        emu = po.phiseg_forward({k: v.clone() for k, v in sd.items()}, patch, mask, eps, training=training,
                                rnd=po.Rounding(True))
        e_emu = po.elbo(emu, mask)
        ref = po.phiseg_forward({k: v.clone() for k, v in sd.items()}, patch, mask, eps, training=training)
        e_ref = po.elbo(ref, mask)
    acc = sum(t.cpu() for t in s)
    acc_emu = po.accumulate_output(emu['s'])
    acc_ref = po.accumulate_output(ref['s'])
    rel_emu, rel_ref = _rel(acc, acc_emu), _rel(acc, acc_ref)
    agree_ref = float((acc.argmax(1) == acc_ref.argmax(1)).float().mean())
    print('\n[%s %s] logits rel-L2 vs bf16-emulating oracle %.3e, vs fp32 oracle %.3e (oracle-emu vs oracle-fp32 %.3e); '
          'argmax agreement vs fp32 %.5f' % (case, key, rel_emu, rel_ref, _rel(acc_emu, acc_ref), agree_ref))
    print('  loss cuda %.6g  emu %.6g  fp32 %.6g  golden(reference) %.6g' %
          (float(loss), float(e_emu['total']), float(e_ref['total']), float(g[key + '_loss'])))
    # (a) same rounding points => only summation order / bf16 tie differences remain
    assert rel_emu < 2e-2
    for lvl in range(5):
        assert _rel(net.prior_mu[lvl].cpu(), emu['prior_mu'][lvl]) < 2e-2
        assert _rel(net.posterior_sigma[lvl].cpu(), emu['post_sigma'][lvl]) < 2e-2
    assert float(loss) == pytest.approx(float(e_emu['total']), rel=2e-2)
    # (b) reference arithmetic (fp32): bf16 storage through ~25 layers
    assert rel_ref < 6e-2
    assert float(loss) == pytest.approx(float(g[key + '_loss']), rel=6e-2)
    assert agree_ref > 0.97


def test_training_step_gradients(golden_dir):
    g, net, sd, patch, mask, eps = _setup('phiseg_small', golden_dir)
    net.train(True)
    with injected_noise(eps):
        net.forward(patch.cuda(), mask.cuda(), training=True)
        loss = net.loss(mask.cuda())
    loss.backward()
    # oracle gradients (fp32 == reference arithmetic)
    sd2 = {k: v.clone() for k, v in sd.items()}
    params = {k: v.requires_grad_(True) for k, v in sd2.items() if v.dtype == torch.float32 and 'running_' not in k}
    out = po.phiseg_forward(sd2, patch, mask, eps, training=True)
    po.elbo(out, mask)['total'].backward()
    named = dict(net.named_parameters())
    nograd = set(str(n) for n in g['train_nograd_names'])
    worst = []
    gmax = max(float(p.grad.norm()) for p in params.values() if p.grad is not None)
    for n, p in params.items():
        if n in nograd:
            assert named[n].grad is None, n          # never-used upsampling_path.4.*
            continue
        got = named[n].grad.cpu()
        if n.endswith('convolution.0.bias') and (n[:-len('0.bias')] + '1.weight') in params:
            assert float(got.abs().max()) == 0.0     # conv bias in front of BatchNorm: exactly zero by construction
            continue
        ref = p.grad
        if float(ref.norm()) < 1e-6 * gmax:
            continue
        worst.append((_rel(got, ref), n))
    worst.sort(reverse=True)
    print('\nworst gradient rel-L2 errors vs fp32 oracle:', worst[:5])
    med = float(np.median([w for w, _ in worst]))
    print('median %.3e' % med)
    assert med < 5e-2
    assert worst[0][0] < 0.35
    # running statistics were updated once, like nn.BatchNorm2d(momentum=0.01)
    k = 'posterior.contracting_path.3.layers.2.convolution.1.running_var'
    np.testing.assert_allclose(net.state_dict()[k].cpu().numpy(), g['train_running_var_probe'], rtol=2e-3)
    kk = 'posterior.contracting_path.3.layers.2.convolution.1.num_batches_tracked'
    assert int(net.state_dict()[kk]) == 4


def test_accumulate_output_aliasing_and_sample(golden_dir):
    """quirk Q2: accumulate_output sums IN PLACE into output_list[-1]; sample() works from cached prior mu/sigma."""
    g, net, sd, patch, mask, eps = _setup('phiseg_small', golden_dir)
    net.eval()
    with injected_noise(eps), torch.no_grad():
        s = net.forward(patch.cuda(), mask.cuda(), training=False)
        before = [t.clone() for t in s]
        probs = net.accumulate_output(s, use_softmax=True)
    total = sum(before)
    torch.testing.assert_close(s[-1], total, rtol=1e-5, atol=1e-5)            # mutated in place
    torch.testing.assert_close(probs, torch.softmax(total, 1), rtol=1e-5, atol=1e-6)
    assert net.s_out_list[-1].data_ptr() == s[-1].data_ptr()
    with torch.no_grad():
        smp = net.sample(testing=True)
    assert tuple(smp.shape) == (int(g['batch']), 2, 128, 128)
    with pytest.raises(NotImplementedError):
        net.sample(testing=False)


def test_rng_stream_matches_reference_call_pattern():
    """quirk Q4: 10 randn_like draws per forward with the reference's shapes and order."""
    net = dropin_phiseg([16, 32, 32, 32, 32, 32, 32]).cuda()
    patch, labels, mask = synth.lidc_like_batch(2, seed=1)
    torch.manual_seed(123)
    with torch.no_grad():
        net.forward(patch.cuda(), mask.cuda(), training=True)
    after = torch.randn(4, device='cuda')
    torch.manual_seed(123)
    for shp in synth.phiseg_noise_shapes(2):
        torch.randn_like(torch.empty(shp, device='cuda'), dtype=torch.float32)
    expect = torch.randn(4, device='cuda')
    assert torch.equal(after, expect)


def test_cpu_tensors_fail_loudly():
    net = dropin_phiseg([16, 32, 32, 32, 32, 32, 32])
    patch, labels, mask = synth.lidc_like_batch(1, seed=1)
    with pytest.raises(RuntimeError):
        net.forward(patch, mask, training=True)
