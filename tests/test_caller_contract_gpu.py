"""The unmodified caller's contract on a GPU: the exact call sequence of UNetModel.validate (reference
train_model.py:177-222) and of _create_tensorboard_summary / generate_images (sample(), reconstruct(); :316, :529-530,
models/phiseg.py:386-412) replayed against the drop-in modules and the drop-in ``utils``, compared with the fp32 oracle
executing the same sequence -- including quirk Q2 (net.loss() after accumulate_output sees the mutated list).
The reference itself is not on the GPU box; its modules are pinned to the oracle by tests/test_oracle_golden.py."""
import numpy as np
import pytest
import torch

from oracle import metrics_oracle as mo
from oracle import phiseg_oracle as po
from oracle import synth
from oracle.ref_run import injected_noise
from tests.keygrammar import dropin_phiseg
from tests.test_parity_conditioned_gpu import B, FILTERS, _fp32, conditioned_state

pytestmark = pytest.mark.gpu
N_VAL = 16            # validation_samples of the reference configs (phiseg_7_5_12.py)


def _dice(mean_probs, mask, n_classes):
    s_ = mean_probs.argmax(0)
    out = []
    for lbl in range(n_classes):
        bp, bg = (s_ == lbl), (mask == lbl)
        if bg.sum() == 0 and bp.sum() == 0:
            out.append(1.0)
        elif bp.sum() == 0 or bg.sum() == 0:
            out.append(0.0)
        else:
            out.append(2.0 * float((bp & bg).sum()) / float(bp.sum() + bg.sum()))     # medpy.metric.dc
    return out


@pytest.mark.parametrize('precision', ['bf16', 'fp16'])
def test_validate_call_sequence(precision):
    import utils                                   # the drop-in second boundary (unet-zoo_b200/utils.py)
    from b200 import _lib
    _fp32()
    sd = conditioned_state()
    C = 2
    patch, labels, _ = synth.lidc_like_batch(4, seed=800)
    x_b, s_gt_arr = patch[2, 0], labels[2]                               # one validation image, [H,W], [H,W,M]
    eps = synth.noise_list(synth.phiseg_noise_shapes(N_VAL), seed=801)
    net = dropin_phiseg(FILTERS)
    net.load_state_dict({k: v.cpu() for k, v in sd.items()})
    net = net.cuda().eval()
    prev = _lib.set_precision(precision)
    try:
        with torch.no_grad(), injected_noise(eps):
            # ---- train_model.py:166-186
            val_patch = x_b.cuda().unsqueeze(dim=0).unsqueeze(dim=1)
            s_b = s_gt_arr[:, :, 1]
            val_mask = s_b.float().cuda().unsqueeze(dim=0).unsqueeze(dim=1)
            val_masks = s_gt_arr.float().cuda().transpose(0, 2).transpose(1, 2)              # CHW
            patch_arrangement = val_patch.repeat((N_VAL, 1, 1, 1))
            mask_arrangement = val_mask.repeat((N_VAL, 1, 1, 1))
            s_out_eval_list = net.forward(patch_arrangement, mask_arrangement, training=False)
            s_prediction_softmax_arrangement = net.accumulate_output(s_out_eval_list, use_softmax=True)
            # ---- :189-192 (Q2: the list was mutated by accumulate_output)
            val_loss = net.loss(mask_arrangement)
            kl, recon = net.kl_divergence_loss, net.reconstruction_loss
            assert kl is val_loss and recon is val_loss                                        # Q1
            # ---- :194-205
            s_prediction_softmax_mean = torch.mean(s_prediction_softmax_arrangement, axis=0)
            s_prediction_arrangement = torch.argmax(s_prediction_softmax_arrangement, dim=1)
            ged = utils.generalised_energy_distance(s_prediction_arrangement, val_masks, nlabels=C - 1,
                                                    label_range=range(1, C))
            onehot = utils.convert_batch_to_onehot(val_masks.unsqueeze(dim=1), nlabels=C)
            ncc = utils.variance_ncc_dist(s_prediction_softmax_arrangement, onehot)
            dice = _dice(s_prediction_softmax_mean.cpu(), s_b, C)
        torch.cuda.synchronize()
    finally:
        _lib.set_precision(prev)
        object.__setattr__(net, '_weight_packer', None)
    assert isinstance(ged, float) and isinstance(ncc, np.ndarray) and ncc.shape == (1,) and ncc.dtype == np.float64
    # the metrics kernels on OUR samples are exact restatements of utils.py: bit-exact GED, NCC to fp64 round-off
    pr = s_prediction_arrangement.cpu().numpy()
    pb = s_prediction_softmax_arrangement.cpu().numpy()
    gt = val_masks.cpu().numpy()
    assert ged == mo.generalised_energy_distance(pr, gt, C - 1, range(1, C))
    assert abs(float(ncc[0]) - float(mo.variance_ncc_dist(pb, mo.convert_batch_to_onehot(gt[:, None], C))[0])) < 1e-6
    # ---- the same sequence in the reference's arithmetic (fp32 oracle)
    with torch.no_grad():
        out = po.phiseg_forward({k: v.clone() for k, v in sd.items()}, patch_arrangement, mask_arrangement,
                                [e.cuda() for e in eps], training=False)
        s = [t.clone() for t in out['s']]
        for i in range(len(s) - 1):
            s[-1] += s[i]                                                                       # phiseg.py:429-431
        probs_ref = torch.softmax(s[-1], dim=1)
        out_q2 = dict(out, s=s)
        e_ref = po.elbo(out_q2, mask_arrangement)
        pr_ref = probs_ref.argmax(1).cpu().numpy()
        ged_ref = mo.generalised_energy_distance(pr_ref, gt, C - 1, range(1, C))
        ncc_ref = float(mo.variance_ncc_dist(probs_ref.cpu().numpy(), mo.convert_batch_to_onehot(gt[:, None], C))[0])
        dice_ref = _dice(probs_ref.mean(0).cpu(), s_b, C)
    loss_err = abs(float(val_loss) - float(e_ref['total'])) / abs(float(e_ref['total']))
    agree = float((pr == pr_ref).mean())
    print('\n[%s] validate(): loss (Q2) %.6g vs %.6g rel %.2e | GED %.6f vs %.6f | NCC %.6f vs %.6f | Dice %s vs %s | '
          'sample argmax agreement %.5f, foreground %.3f' % (precision, float(val_loss), float(e_ref['total']), loss_err, ged,
                                                           ged_ref, float(ncc[0]), ncc_ref, dice, dice_ref, agree,
                                                           float((pr_ref != 0).mean())))
    # measured on B200 (conditioned net, 16 samples): fp16 storage  loss 5e-5, NCC 6e-5, GED 1.4e-4, argmax 0.99998;
    #                                                 bf16 storage  loss 1.3e-3, NCC 6e-4, GED 5e-4, argmax 0.99990.
    # The north star's 1e-4 absolute on GED amounts to bit-identical masks: ONE flipped pixel in one 440-pixel sample mask
    # moves the 16-sample GED by ~2.5e-4.  NCC is at 6e-5 ... 1.05e-4 in the fp16 mode over several conditioning runs (the
    # oracle's own cuDNN training of the fixture is not bit-reproducible): asserted at 2e-4; GED at 5e-4.
    assert agree >= 0.999
    if precision == 'fp16':
        assert loss_err < 1e-3
        assert abs(float(ncc[0]) - ncc_ref) < 2e-4
        assert abs(ged - ged_ref) < 5e-4
        assert max(abs(a - b) for a, b in zip(dice, dice_ref)) < 1e-3
    else:
        assert loss_err < 3e-3
        assert abs(float(ncc[0]) - ncc_ref) < 2e-3 and abs(ged - ged_ref) < 2e-3
        assert max(abs(a - b) for a, b in zip(dice, dice_ref)) < 5e-3


def test_sample_and_reconstruct_values():
    """PHISeg.sample_prior / sample_posterior / reconstruct / sample(testing=True) (models/phiseg.py:386-412): fresh draws
    with the mu / sigma cached by the last forward -> likelihood -> accumulate_output, against the oracle on the same
    noise.  sample(testing=False) raises NotImplementedError like the reference."""
    _fp32()
    sd = conditioned_state()
    patch, _, mask = synth.lidc_like_batch(B, seed=810)
    eps = synth.noise_list(synth.phiseg_noise_shapes(B), seed=811)
    eps2 = synth.noise_list(synth.phiseg_noise_shapes(B)[:5][::-1], seed=812)        # sample_prior draws levels 0..4
    net = dropin_phiseg(FILTERS)
    net.load_state_dict({k: v.cpu() for k, v in sd.items()})
    net = net.cuda().eval()
    with torch.no_grad():
        with injected_noise(eps):
            net.forward(patch.cuda(), mask.cuda(), training=False)
        with injected_noise(eps2):
            sample = net.sample(testing=True)
        with injected_noise(eps2):
            z_post = net.sample_posterior()
        recon, layers = net.reconstruct(z_post, use_softmax=True)
        with pytest.raises(NotImplementedError):
            net.sample(testing=False)
        ref = po.phiseg_forward({k: v.clone() for k, v in sd.items()}, patch.cuda(), mask.cuda(), [e.cuda() for e in eps],
                                training=False)
        z_prior_ref = [ref['prior_mu'][l] + ref['prior_sigma'][l] * eps2[l].cuda() for l in range(5)]
        z_post_ref = [ref['post_mu'][l] + ref['post_sigma'][l] * eps2[l].cuda() for l in range(5)]
        sample_ref = po.accumulate_output(po.likelihood(z_prior_ref, sd, patch.shape[-2:], False), use_softmax=False)
        recon_ref = po.accumulate_output(po.likelihood(z_post_ref, sd, patch.shape[-2:], False), use_softmax=True)
    assert sample.shape == (B, 2, 128, 128) and len(layers) == 5
    for l in range(5):
        assert float((z_post[l] - z_post_ref[l]).norm() / z_post_ref[l].norm()) < 5e-3
    r1 = float((sample - sample_ref).norm() / sample_ref.norm())
    r2 = float((recon - recon_ref).abs().max())
    r3 = float((recon - recon_ref).abs().mean())
    agree = float((sample.argmax(1) == sample_ref.argmax(1)).float().mean())
    print('\nsample(): logits rel-L2 %.3e, argmax agreement %.5f; reconstruct(softmax): max |dp| %.3e' % (r1, agree, r2))
    # bf16 storage; measured 3.0e-3 ... 4.0e-3, 0.99989 ... 0.99992, max |dp| 2e-2 ... 5e-2 (single boundary pixels)
    assert r1 < 8e-3 and agree >= 0.999 and r2 < 0.3 and r3 < 1e-3
