"""Generate tests/golden/*.npz from the REAL reference modules (run in the build container only):

    python -m oracle.make_golden

Fixtures hold seeds + reference OUTPUTS only; weights and inputs are re-synthesised from the seeds by
oracle/synth.py, so the files stay small.  Everything here executes /root/reference code unmodified
(through oracle/ref_loader.py shims) -- TEST INFRASTRUCTURE.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import synth                                    # noqa: E402
from oracle.ref_loader import load_reference                # noqa: E402
from oracle.ref_run import (build_reference_phiseg, build_reference_phiseg3d, injected_noise,  # noqa: E402
                            phiseg3d_patches)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def phiseg_case(tag, filters, batch, wseed, dseed, nseed, keep_logits, reversible=False):
    net = build_reference_phiseg(filters, reversible=reversible)
    sd = synth.synth_state_dict(net.state_dict(), seed=wseed)
    patch, labels, mask = synth.lidc_like_batch(batch, seed=dseed)
    eps = synth.noise_list(synth.phiseg_noise_shapes(batch), seed=nseed)
    out = {'filters': np.asarray(filters), 'batch': batch, 'wseed': wseed, 'dseed': dseed, 'nseed': nseed,
           'reversible': int(reversible)}
    for training in (True, False):
        net.load_state_dict(sd)
        net.train(training)
        key = 'train' if training else 'eval'
        with injected_noise(eps):
            s = net.forward(patch, mask, training=training)
            s = [t.clone() for t in s]
            net.loss_dict = {}
            loss = net.loss(mask)
        out[key + '_loss'] = float(loss)
        for k, v in net.loss_dict.items():
            out['%s_%s' % (key, k)] = float(v)
        for lvl in range(5):
            out['%s_post_mu%d' % (key, lvl)] = net.posterior_mu[lvl].detach().numpy()
            out['%s_post_sigma%d' % (key, lvl)] = net.posterior_sigma[lvl].detach().numpy()
            out['%s_prior_mu%d' % (key, lvl)] = net.prior_mu[lvl].detach().numpy()
            out['%s_prior_sigma%d' % (key, lvl)] = net.prior_sigma[lvl].detach().numpy()
        acc = sum(s)
        out[key + '_logit_absmean'] = float(acc.abs().mean())
        if keep_logits:
            out[key + '_logits'] = acc.detach().numpy().astype(np.float32)
        else:
            out[key + '_logits_ds8'] = acc.detach()[:, :, ::8, ::8].numpy().astype(np.float32)
        if training:
            net.zero_grad()
            loss.backward()
            gn = {n: float(p.grad.norm()) for n, p in net.named_parameters() if p.grad is not None}
            names = sorted(gn)
            out['train_grad_names'] = np.asarray(names)
            out['train_grad_norms'] = np.asarray([gn[n] for n in names])
            out['train_nograd_names'] = np.asarray(sorted(n for n, p in net.named_parameters() if p.grad is None))
            rs = net.state_dict()
            k = ('posterior.contracting_path.3.layers.1.sequence.reversible_blocks.1.f_block.0.convolution.1.running_var'
                 if reversible else 'posterior.contracting_path.3.layers.2.convolution.1.running_var')
            out['train_running_var_probe'] = rs[k].numpy().copy()
            out['train_running_var_probe_key'] = k
            out['train_num_batches_tracked_probe'] = int(rs[k.replace('running_var', 'num_batches_tracked')])
    np.savez_compressed(os.path.join(GOLDEN, tag + '.npz'), **out)
    print(tag, out['train_loss'], out['eval_loss'])


def phiseg3d_case(tag, filters, latent_levels, size, batch, wseed, dseed, nseed, reversible=False):
    """PHISeg3D of the reference under the three documented patches (oracle/ref_run.py)."""
    net = build_reference_phiseg3d(filters, (4, size, size, size), latent_levels, reversible=reversible)
    sd = synth.synth_state_dict(net.state_dict(), seed=wseed)
    vol, lab = synth.brats_like_batch(batch, size=size, seed=dseed)
    eps = synth.noise_list(synth.phiseg3d_noise_shapes(batch, size, latent_levels, len(filters)), seed=nseed)
    out = {'filters': np.asarray(filters), 'latent_levels': latent_levels, 'size': size, 'batch': batch, 'wseed': wseed,
           'dseed': dseed, 'nseed': nseed, 'reversible': int(reversible)}
    for training in (True, False):
        net.load_state_dict(sd)
        net.train(training)
        key = 'train' if training else 'eval'
        with phiseg3d_patches(net), injected_noise(eps):
            s = net.forward(vol, lab, training=training)
            s = [t.clone() for t in s]
            net.loss_dict = {}
            loss = net.loss(lab)
        out[key + '_loss'] = float(loss)
        for k, v in net.loss_dict.items():
            out['%s_%s' % (key, k)] = float(v)
        for lvl in range(latent_levels):
            st = 2 if lvl == 0 else 1            # the full-resolution level is stored at every second voxel
            for nm, lst in (('post_mu', net.posterior_mu), ('post_sigma', net.posterior_sigma),
                            ('prior_mu', net.prior_mu), ('prior_sigma', net.prior_sigma)):
                out['%s_%s%d' % (key, nm, lvl)] = lst[lvl].detach()[:, :, ::st, ::st, ::st].numpy()
        acc = sum(s)
        out[key + '_logit_absmean'] = float(acc.abs().mean())
        out[key + '_logits_ds2'] = acc.detach()[:, :, ::2, ::2, ::2].numpy().astype(np.float32)
        if training:
            net.zero_grad()
            loss.backward()
            gn = {n: float(p.grad.norm()) for n, p in net.named_parameters() if p.grad is not None}
            names = sorted(gn)
            out['train_grad_names'] = np.asarray(names)
            out['train_grad_norms'] = np.asarray([gn[n] for n in names])
            out['train_nograd_names'] = np.asarray(sorted(n for n, p in net.named_parameters() if p.grad is None))
            k = 'posterior.contracting_path.1.layers.2.convolution.1.running_var'
            if reversible:
                k = 'posterior.contracting_path.1.layers.1.sequence.reversible_blocks.0.f_block.0.convolution.1.running_var'
            out['train_running_var_probe'] = net.state_dict()[k].numpy().copy()
            out['train_running_var_probe_key'] = k
    np.savez_compressed(os.path.join(GOLDEN, tag + '.npz'), **out)
    print(tag, out['train_loss'], out['eval_loss'])


def metrics_case():
    ns = load_reference()
    rs = np.random.RandomState(11)
    out = {}
    # GED: N samples, M annotators, binary + 3-class, with empty masks
    for tag, (N, M, C, S) in {'bin': (7, 4, 2, 32), 'tri': (5, 3, 3, 24)}.items():
        samples = (rs.uniform(size=(N, S, S)) < 0.3).astype(np.int64) * rs.randint(1, C, (N, S, S))
        gts = (rs.uniform(size=(M, S, S)) < 0.3).astype(np.int64) * rs.randint(1, C, (M, S, S))
        samples[0] = 0
        gts[-1] = 0
        ged = ns.utils.generalised_energy_distance(torch.from_numpy(samples), torch.from_numpy(gts).float(),
                                                   nlabels=C - 1, label_range=range(1, C))
        logits = rs.standard_normal((N, C, S, S)).astype(np.float32) * 2
        probs = torch.softmax(torch.from_numpy(logits), 1)
        onehot = ns.utils.convert_batch_to_onehot(torch.from_numpy(gts).float().unsqueeze(1), nlabels=C)
        ncc = ns.utils.variance_ncc_dist(probs, onehot)
        out[tag + '_samples'] = samples.astype(np.uint8)
        out[tag + '_gts'] = gts.astype(np.uint8)
        out[tag + '_logits'] = logits
        out[tag + '_ged'] = float(ged)
        out[tag + '_ncc'] = np.asarray(ncc, np.float64)
        out[tag + '_onehot'] = onehot.numpy().astype(np.uint8)
    # KL known answer from the reference method
    net = build_reference_phiseg([16] * 7)
    kl = net.KL_two_gauss_with_diag_cov(torch.tensor([[[[0.5, -1.0]]]]), torch.tensor([[[[0.8, 1.5]]]]),
                                        torch.tensor([[[[0.0, 0.25]]]]), torch.tensor([[[[1.2, 0.7]]]]))
    out['kl1'] = float(kl)
    np.savez_compressed(os.path.join(GOLDEN, 'metrics.npz'), **out)
    print('metrics', out['bin_ged'], out['bin_ncc'], out['tri_ged'], out['tri_ncc'], out['kl1'])


def unet_cases():
    import torch.distributions.normal as tdn
    ns = load_reference()
    out = {}
    # U-Net (BASELINE configs[0] architecture at reduced width): forward + CE-mean loss + gradient norms
    filters = [16, 32, 32, 48]
    net = ns.unet.Unet(1, 2, filters)
    sd = synth.synth_state_dict(net.state_dict(), seed=2)
    net.load_state_dict(sd)
    patch, labels, mask = synth.lidc_like_batch(3, seed=4)
    logits = net.forward(patch)
    loss = net.loss(mask)
    loss.backward()
    out['unet_filters'] = np.asarray(filters)
    out['unet_logits_ds4'] = logits.detach()[:, :, ::4, ::4].numpy()
    out['unet_loss'] = float(loss)
    names = sorted(n for n, p in net.named_parameters())
    out['unet_grad_names'] = np.asarray(names)
    gd = dict(net.named_parameters())
    out['unet_grad_norms'] = np.asarray([float(gd[n].grad.norm()) for n in names])
    # Probabilistic U-Net, latent_dim 6 (BASELINE configs[1]) at reduced width
    f7 = [32, 32, 32, 32, 32, 32, 32]
    pn = ns.probabilistic_unet.ProbabilisticUnet(input_channels=1, num_classes=2, num_filters=f7, latent_dim=6,
                                                 no_convs_fcomb=3)
    sdp = synth.synth_state_dict(pn.state_dict(), seed=3)
    eps = synth.noise_list([(3, 6)], seed=8)[0]
    orig = tdn._standard_normal
    tdn._standard_normal = lambda shape, dtype, device: eps.to(device)
    try:
        for training in (True, False):
            key = 'train' if training else 'eval'
            pn.load_state_dict(sdp)
            pn.train(training)
            fwd = pn.forward(patch, mask, training=training)
            loss = pn.loss(mask)
            out['prob_%s_forward_ds4' % key] = fwd.detach()[:, :, ::4, ::4].numpy()
            out['prob_%s_loss' % key] = float(loss)
            out['prob_%s_kl' % key] = float(pn.kl_divergence_loss)
            out['prob_%s_rec' % key] = float(pn.reconstruction_loss)
            out['prob_%s_mu_q' % key] = pn.posterior_latent_space.mean.detach().numpy()
            out['prob_%s_sigma_p' % key] = pn.prior_latent_space.stddev.detach().numpy()
            out['prob_%s_reconstruction_ds4' % key] = pn.reconstruction.detach()[:, :, ::4, ::4].numpy()
            if training:
                pn.zero_grad()
                loss.backward()
                gd = {n: p.grad for n, p in pn.named_parameters()}
                names = sorted(n for n, g in gd.items() if g is not None)
                out['prob_grad_names'] = np.asarray(names)
                out['prob_grad_norms'] = np.asarray([float(gd[n].norm()) for n in names])
                out['prob_nograd_names'] = np.asarray(sorted(n for n, g in gd.items() if g is None))
    finally:
        tdn._standard_normal = orig
    out['prob_filters'] = np.asarray(f7)
    np.savez_compressed(os.path.join(GOLDEN, 'unet_probunet.npz'), **out)
    print('unet', out['unet_loss'], 'probunet', out['prob_train_loss'], out['prob_eval_loss'])


if __name__ == '__main__':
    torch.manual_seed(0)
    os.makedirs(GOLDEN, exist_ok=True)
    phiseg_case('phiseg_small', [16, 32, 32, 32, 32, 32, 32], 4, 1, 3, 5, keep_logits=True)
    phiseg_case('phiseg_lidc', [32, 64, 128, 192, 192, 192, 192], 2, 2, 4, 6, keep_logits=False)
    phiseg_case('phiseg_rev_small', [32, 64, 64, 64, 64, 64, 64], 4, 1, 3, 5, keep_logits=False, reversible=True)
    phiseg3d_case('phiseg3d_small', [32, 64, 64], 3, 32, 2, 1, 3, 5)
    phiseg3d_case('phiseg3d_rev_small', [32, 64, 64], 3, 16, 2, 2, 4, 6, reversible=True)
    metrics_case()
    unet_cases()
