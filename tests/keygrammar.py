"""state_dict templates (names, shapes, dtypes) built from the drop-in modules on CPU -- pure host logic.
tests/test_oracle_golden.py::test_oracle_matches_live_reference checks them against the real reference keys."""
from tests.gpu_util import PKG  # noqa: F401  (puts unet-zoo_b200/ on sys.path)


def dropin_phiseg(filters, reversible=False, image_size=(1, 128, 128), num_classes=2, input_channels=1):
    from models.phiseg import PHISeg
    return PHISeg(input_channels=input_channels, num_classes=num_classes, num_filters=list(filters), latent_levels=5,
                  no_convs_fcomb=4, beta=10.0, image_size=image_size, reversible=reversible)


def phiseg_state_template(filters, reversible=False):
    return dropin_phiseg(filters, reversible=reversible).state_dict()


def dropin_phiseg3d(filters, latent_levels, image_size, reversible=False, num_classes=3, input_channels=4):
    from models.phiseg3D import PHISeg3D
    return PHISeg3D(input_channels=input_channels, num_classes=num_classes, num_filters=list(filters),
                    latent_levels=latent_levels, no_convs_fcomb=4, beta=10.0, image_size=image_size,
                    reversible=reversible)
