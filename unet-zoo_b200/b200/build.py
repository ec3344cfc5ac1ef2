"""Constructors of the drop-in models with the keyword set ``UNetModel.__init__`` uses (reference train_model.py:34-42)
for the configurations of BASELINE.json -- shared by bench.py, the tools and the tests."""


def phiseg(filters, reversible=False, image_size=(1, 128, 128), num_classes=2, input_channels=1):
    """models/experiments/phiseg_7_5_12.py / phiseg_rev_7_5_12.py"""
    from models.phiseg import PHISeg
    return PHISeg(input_channels=input_channels, num_classes=num_classes, num_filters=list(filters), latent_levels=5,
                  no_convs_fcomb=4, beta=10.0, image_size=image_size, reversible=reversible)


def phiseg3d(filters, latent_levels, image_size, reversible=False, num_classes=3, input_channels=4):
    """models/experiments/phiseg_brats.py under the fixed specification of SURVEY.md 8c"""
    from models.phiseg3D import PHISeg3D
    return PHISeg3D(input_channels=input_channels, num_classes=num_classes, num_filters=list(filters),
                    latent_levels=latent_levels, no_convs_fcomb=4, beta=10.0, image_size=image_size,
                    reversible=reversible)


def probunet(filters, latent_dim=6, num_classes=2, input_channels=1, no_convs_fcomb=3):
    """models/experiments/prob_unet.py (latent_dim 6 via the direct constructor)"""
    from models.probabilistic_unet import ProbabilisticUnet
    return ProbabilisticUnet(input_channels=input_channels, num_classes=num_classes, num_filters=list(filters),
                             latent_dim=latent_dim, no_convs_fcomb=no_convs_fcomb)


def unet(filters=(32, 64, 128, 192), num_classes=2, input_channels=1):
    """models/experiments/unet.py"""
    from models.unet import Unet
    return Unet(input_channels, num_classes, list(filters))
