"""state_dict templates (names, shapes, dtypes) built from the drop-in modules on CPU -- pure host logic.
tests/test_oracle_golden.py::test_oracle_matches_live_reference checks them against the real reference keys."""
from tests.gpu_util import PKG  # noqa: F401  (puts unet-zoo_b200/ on sys.path)


from b200 import build as _build


def dropin_phiseg(filters, reversible=False, image_size=(1, 128, 128), num_classes=2, input_channels=1):
    return _build.phiseg(filters, reversible, image_size, num_classes, input_channels)


def phiseg_state_template(filters, reversible=False):
    return dropin_phiseg(filters, reversible=reversible).state_dict()


def dropin_phiseg3d(filters, latent_levels, image_size, reversible=False, num_classes=3, input_channels=4):
    return _build.phiseg3d(filters, latent_levels, image_size, reversible, num_classes, input_channels)
