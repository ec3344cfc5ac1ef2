"""PHiSeg 7/5, batch 12 -- the attributes of the reference's models/experiments/phiseg_7_5_12.py with the synthetic
LIDC-shaped data plug-in, short run lengths for smoke runs.  Usage:
    python unet-zoo_b200/launch.py --reference /path/to/UNet-Zoo unet-zoo_b200/experiments_b200/phiseg_7_5_12_synthetic.py local dummy
"""
from models.phiseg import PHISeg
from synthetic_data import synthetic_lidc
from utils import normalise_image

experiment_name = 'PHISeg_7_5_12_synthetic'
log_dir_name = 'lidc_synthetic'
data_loader = synthetic_lidc

filter_channels = [32, 64, 128, 192, 192, 192, 192]
latent_levels = 5
iterations = 41
n_classes = 2
num_labels_per_subject = 4
no_convs_fcomb = 4
beta = 10.0
use_reversible = False
exponential_weighting = True
input_channels = 1
epochs_to_train = 20
batch_size = 12
image_size = (1, 128, 128)
augmentation_options = {'do_flip_lr': True, 'do_flip_ud': True, 'do_rotations': True, 'do_scaleaug': True,
                        'nlabels': n_classes}
input_normalisation = normalise_image
validation_samples = 16
num_validation_images = 4
logging_frequency = 10
validation_frequency = 20
weight_decay = 10e-5
pretrained_model = None
model = PHISeg
