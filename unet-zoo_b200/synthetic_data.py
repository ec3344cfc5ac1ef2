"""Synthetic LIDC-shaped data plug-in for the reference's data contract (SURVEY.md 8b "Data plug-in contract"):
``exp_config.data_loader(sys_config=, exp_config=)`` -> object with ``train.next_batch(B)``, ``validation.images /
labels``, ``test.images / labels`` (reference data/lidc_data.py:10-53, data/batch_provider.py:43-67,126-137).
Lets the unmodified train_model.py run without the LIDC pickle; real-data loaders keep working through the reference's
own data package."""
import numpy as np


def _make_split(n, size, annotators, rs, empty_frac=0.25):
    images = np.clip(rs.standard_normal((n, size, size)) * 0.25, -0.5, 0.5).astype(np.float32)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    labels = np.zeros((n, size, size, annotators), np.uint8)
    for b in range(n):
        cy, cx = rs.uniform(0.3 * size, 0.7 * size, 2)
        ry, rx = rs.uniform(0.06 * size, 0.18 * size, 2)
        for m in range(annotators):
            if rs.uniform() < empty_frac:
                continue
            jy, jx = rs.normal(0, 0.02 * size, 2)
            sy, sx = rs.uniform(0.8, 1.25, 2)
            labels[b, :, :, m] = (((yy - cy - jy) / (ry * sy)) ** 2 + ((xx - cx - jx) / (rx * sx)) ** 2) <= 1.0
        images[b] += 0.2 * labels[b].mean(-1)
    return images, labels


class _Split:
    def __init__(self, images, labels, annotator_range, rs):
        self.images, self.labels = images, labels
        self.annotator_range = list(annotator_range)
        self.rs = rs

    def next_batch(self, batch_size):
        """(x [B,1,H,W] float, s [B,H,W]) with one random annotator per image (batch_provider.py:61-63,126-137)."""
        idx = self.rs.choice(self.images.shape[0], size=batch_size, replace=batch_size > self.images.shape[0])
        pick = self.rs.choice(self.annotator_range, size=batch_size)
        x = self.images[idx][:, None]
        s = np.stack([self.labels[i, :, :, a] for i, a in zip(idx, pick)])
        return x, s


class synthetic_lidc:
    def __init__(self, sys_config=None, exp_config=None, n_train=256, n_val=16, n_test=16, seed=0):
        size = exp_config.image_size[1]
        m = exp_config.num_labels_per_subject
        if not hasattr(exp_config, 'annotator_range'):
            exp_config.annotator_range = range(m)            # reference data/lidc_data.py:31-32
        rank = 0
        try:
            import os
            rank = int(os.environ.get('RANK', '0'))
        except Exception:
            pass
        rs = np.random.RandomState(seed + rank)
        self.train = _Split(*_make_split(n_train, size, m, rs), exp_config.annotator_range, rs)
        self.validation = _Split(*_make_split(n_val, size, m, rs), exp_config.annotator_range, rs)
        self.test = _Split(*_make_split(n_test, size, m, rs), exp_config.annotator_range, rs)
