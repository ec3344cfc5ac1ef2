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


# ------------------------------------------------------------------------------------------------ PHISeg3D
class _Utils3D:
    """Stands where ``utils`` does inside the reference's models/phiseg3D.py while a patched model runs: fix (iii) of
    SURVEY.md 8c -- an index volume [B,1,D,H,W] is one-hot encoded with num_classes labels (the shipped code asks for
    2 labels at :253 and its helper passes 4-D per-sample volumes through untouched, utils.py:296-298)."""

    def __init__(self, real_utils, num_classes):
        self._real = real_utils
        self._num_classes = num_classes

    def convert_batch_to_onehot(self, lblbatch, nlabels):
        return torch.cat([(lblbatch == k) for k in range(self._num_classes)], dim=1).long()

    def __getattr__(self, name):
        return getattr(self._real, name)


@contextlib.contextmanager
def phiseg3d_patches(net):
    """The three documented monkey-patches that make the UNMODIFIED reference PHISeg3D runnable:
    (i) is the caller's choice of a channel-consistent filter list; (ii) nearest upsampling of the logits to the full
    volume image_size[1:4] (the file passes a 2-element size for a 5-D tensor, :398); (iii) see _Utils3D."""
    ns = load_reference()
    mod = ns.phiseg3D
    orig_interp = torch.nn.functional.interpolate
    orig_utils = mod.utils
    full = list(net.image_size[1:4])

    def interp(x, size=None, **kw):
        if x.dim() == 5 and size is not None and len(size) == 2 and kw.get('mode') == 'nearest':
            size = full
        return orig_interp(x, size=size, **kw)

    torch.nn.functional.interpolate = interp
    mod.utils = _Utils3D(orig_utils, net.num_classes)
    try:
        yield
    finally:
        torch.nn.functional.interpolate = orig_interp
        mod.utils = orig_utils


def build_reference_phiseg3d(num_filters, image_size, latent_levels, input_channels=4, num_classes=3, reversible=False):
    ns = load_reference()
    return ns.phiseg3D.PHISeg3D(input_channels=input_channels, num_classes=num_classes, num_filters=list(num_filters),
                                latent_levels=latent_levels, no_convs_fcomb=4, beta=10.0, image_size=image_size,
                                reversible=reversible)
