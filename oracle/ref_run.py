"""Run the REAL reference modules with injected weights / noise (build container only; test infra)."""
import contextlib

import torch

from .ref_loader import load_reference


@contextlib.contextmanager
def injected_noise(eps_list):
    """Replace torch.randn_like by a pop from ``eps_list`` (reference models/phiseg.py:104 draws inside
    the module; SURVEY.md 8c 'noise injection')."""
    queue = list(eps_list)
    orig = torch.randn_like

    def fake(t, **kw):
        e = queue.pop(0)
        assert tuple(e.shape) == tuple(t.shape), (e.shape, t.shape)
        return e.to(t.device)

    torch.randn_like = fake
    try:
        yield queue
    finally:
        torch.randn_like = orig


def build_reference_phiseg(num_filters, image_size=(1, 128, 128), reversible=False, input_channels=1, num_classes=2):
    ns = load_reference()
    return ns.phiseg.PHISeg(input_channels=input_channels, num_classes=num_classes, num_filters=list(num_filters),
                            latent_levels=5, no_convs_fcomb=4, beta=10.0, image_size=image_size, reversible=reversible)
