"""U-Net / Probabilistic U-Net: the CPU oracle (oracle/unet_oracle.py) against fixtures generated from the REAL
reference (oracle/make_golden.py::unet_cases) and the drop-in modules' key grammar / parameter counts."""
import os

import numpy as np
import pytest
import torch

from oracle import synth
from oracle import unet_oracle as uo
from tests.gpu_util import PKG  # noqa: F401


def _unet(filters):
    from models.unet import Unet
    return Unet(1, 2, list(filters))


def _probunet(filters, latent_dim=6, no_convs_fcomb=3):
    from models.probabilistic_unet import ProbabilisticUnet
    return ProbabilisticUnet(input_channels=1, num_classes=2, num_filters=list(filters), latent_dim=latent_dim,
                             no_convs_fcomb=no_convs_fcomb)


def test_unet_oracle_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, 'unet_probunet.npz'))
    filters = [int(v) for v in g['unet_filters']]
    sd = synth.synth_state_dict(_unet(filters).state_dict(), seed=2)
    params = {k: v.requires_grad_(True) for k, v in sd.items()}
    patch, labels, mask = synth.lidc_like_batch(3, seed=4)
    logits = uo.unet_forward(sd, patch, len(filters))
    loss = uo.unet_loss(logits, mask)
    np.testing.assert_allclose(logits.detach()[:, :, ::4, ::4].numpy(), g['unet_logits_ds4'], rtol=1e-4, atol=1e-5)
    assert float(loss) == pytest.approx(float(g['unet_loss']), rel=1e-6)
    loss.backward()
    for n, ref in zip(g['unet_grad_names'], g['unet_grad_norms']):
        assert float(params[str(n)].grad.norm()) == pytest.approx(float(ref), rel=1e-3, abs=1e-7), n


@pytest.mark.parametrize('training', [True, False])
def test_probunet_oracle_matches_reference_fixture(golden_dir, training):
    g = np.load(os.path.join(golden_dir, 'unet_probunet.npz'))
    key = 'train' if training else 'eval'
    filters = [int(v) for v in g['prob_filters']]
    sd = synth.synth_state_dict(_probunet(filters).state_dict(), seed=3)
    patch, labels, mask = synth.lidc_like_batch(3, seed=4)
    eps = synth.noise_list([(3, 6)], seed=8)[0]
    with torch.no_grad():
        o = uo.probunet_step(sd, patch, mask, eps, 7, 6, 3, training=training)
    assert float(o['loss']) == pytest.approx(float(g['prob_%s_loss' % key]), rel=1e-5)
    assert float(o['kl']) == pytest.approx(float(g['prob_%s_kl' % key]), rel=1e-4)
    assert float(o['reconstruction_loss']) == pytest.approx(float(g['prob_%s_rec' % key]), rel=1e-5)
    np.testing.assert_allclose(o['mu_q'].numpy(), g['prob_%s_mu_q' % key], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(o['sigma_p'].numpy(), g['prob_%s_sigma_p' % key], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(o['forward'][:, :, ::4, ::4].numpy(), g['prob_%s_forward_ds4' % key], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(o['reconstruction'][:, :, ::4, ::4].numpy(), g['prob_%s_reconstruction_ds4' % key],
                               rtol=1e-3, atol=1e-3)


def test_dropin_parameter_counts_and_key_sets(golden_dir):
    """SURVEY.md Appendix B / D."""
    u = _unet([32, 64, 128, 192])
    assert sum(p.numel() for p in u.parameters()) == 2260194 and len(u.state_dict()) == 44
    p = _probunet([32, 64, 128, 192, 192, 192, 192], latent_dim=6, no_convs_fcomb=3)
    assert sum(q.numel() for q in p.parameters()) == 17956988 and len(p.state_dict()) == 394
    g = np.load(os.path.join(golden_dir, 'unet_probunet.npz'))
    small = _probunet([int(v) for v in g['prob_filters']])
    names = set(n for n, _ in small.named_parameters())
    assert set(str(n) for n in g['prob_grad_names']) | set(str(n) for n in g['prob_nograd_names']) == names
    assert set(str(n) for n in g['unet_grad_names']) == set(n for n, _ in _unet(g['unet_filters']).named_parameters())
