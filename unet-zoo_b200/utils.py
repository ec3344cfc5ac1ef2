"""Drop-in for the hot-path part of the reference's ``utils.py`` (second boundary).

The unmodified caller does ``import utils`` (train_model.py:17) and calls
  utils.generalised_energy_distance(samples[N,H,W] int64 cuda, gt[M,H,W] float cuda, nlabels=, label_range=) -> float
  utils.variance_ncc_dist(probs[N,C,H,W] fp32 cuda, gt_onehot[M,C,H,W] int64)          -> np.float64 array, shape (1,)
  utils.convert_batch_to_onehot(lbl[M,1,H,W], nlabels)                                  -> int64 tensor
(train_model.py:198-205,398-406) plus makefolder / setup_logger (:588,592); the experiment files import
``normalise_image`` (e.g. models/experiments/phiseg_7_5_12.py:5); models use init_weights / l2_regularisation.

GED / NCC / one-hot run as CUDA kernels (one D2H of the final scalars); the small host helpers are restated here.
Host-side augmentation / NIfTI helpers of the reference (utils.py:12-67,250-268,350-370; out of scope)
are forwarded lazily to the reference's own file when it is importable next on sys.path.
"""
import importlib.util
import logging
import os
import sys

import numpy as np
import torch
import torch.nn as nn

from b200 import kern


# ------------------------------------------------------------------------------------------------ device metrics
def generalised_energy_distance(sample_arr, gt_arr, nlabels=1, **kwargs):
    """reference utils.py:148-200.  sample_arr [N,X,Y], gt_arr [M,X,Y]; returns a python float, bit-identical to the
    reference's nested loops (integer IoU counts, fp64 ratios, Python-order sums)."""
    label_range = list(kwargs.get('label_range', range(nlabels)))
    if len(label_range) != nlabels:
        # the reference divides the summed IoUs by ``nlabels`` whatever the length of ``label_range`` (utils.py:170); the
        # kernel divides by the number of labels it walks.  The two agree for every call site of the reference
        # (train_model.py:198-200,398-400: nlabels = n_classes - 1, label_range = range(1, n_classes)).
        raise NotImplementedError('generalised_energy_distance: len(label_range) = %d != nlabels = %d is not supported '
                                  'by the B200 kernel (the reference call sites always pass matching values)'
                                  % (len(label_range), nlabels))
    sample_arr = _as_cuda_labels(sample_arr)
    gt_arr = _as_cuda_labels(gt_arr).to(sample_arr.device)
    out = kern.ged(sample_arr, gt_arr, label_range)
    return float(out[0].item())


def variance_ncc_dist(sample_arr, gt_arr):
    """reference utils.py:202-247.  sample_arr [N,C,X,Y] probabilities, gt_arr [M,C,X,Y] one-hot.
    Returns numpy float64 array of shape (1,) like the reference (np.correlate output)."""
    if not torch.is_tensor(sample_arr):
        sample_arr = torch.as_tensor(np.asarray(sample_arr))
    if not torch.is_tensor(gt_arr):
        gt_arr = torch.as_tensor(np.asarray(gt_arr))
    sample_arr = _need_cuda(sample_arr.detach()).float()
    gt_arr = gt_arr.detach().to(sample_arr.device)
    if gt_arr.dtype not in (torch.int64, torch.float32, torch.uint8):
        gt_arr = gt_arr.float()
    return kern.variance_ncc(sample_arr, gt_arr).cpu().numpy()


def _need_cuda(t):
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise kern._lib.UnetZooLibError('UNet-Zoo B200 metrics need a CUDA device: there is no CPU fallback path')
        t = t.cuda()
    return t


def _as_cuda_labels(a):
    if not torch.is_tensor(a):
        a = torch.as_tensor(np.asarray(a))
    a = _need_cuda(a.detach())
    if a.dtype not in (torch.int64, torch.float32, torch.uint8):
        a = a.float() if a.is_floating_point() else a.long()
    return a


def ncc(a, v, zero_norm=True):
    """reference utils.py:130-145 (host numpy; kept for API completeness, the device path is variance_ncc_dist)."""
    a = np.asarray(a).flatten()
    v = np.asarray(v).flatten()
    if zero_norm:
        a = (a - np.mean(a)) / (np.std(a) * len(a))
        v = (v - np.mean(v)) / np.std(v)
    else:
        a = a / (np.std(a) * len(a))
        v = v / np.std(v)
    return np.correlate(a, v)


# ------------------------------------------------------------------------------------------------ one-hot
def convert_to_onehot(lblmap, nlabels):
    """numpy HW -> HWC one-hot (reference utils.py:279-285)."""
    lblmap = np.asarray(lblmap)
    return np.stack([(lblmap == k).astype(np.uint8) for k in range(nlabels)], axis=-1).astype(np.float64)


def convert_to_onehot_torch(lblmap, nlabels):
    """reference utils.py:289-299: CHW/1HW index map -> [nlabels,H,W] int64; 4-D (BraTS one-hot) passes through."""
    if len(lblmap.shape) == 3:
        flat = lblmap.reshape(lblmap.shape[-2], lblmap.shape[-1])
        return torch.stack([(flat == k) for k in range(nlabels)], dim=0).long()
    return lblmap.long()


def convert_batch_to_onehot(lblbatch, nlabels):
    """reference utils.py:303-311 without the per-image / per-label host loop: one vectorised comparison on whatever
    device the labels live on.  [B,1,H,W] -> int64 [B,nlabels,H,W]  (bit exact)."""
    if len(lblbatch.shape) == 5:           # BraTS volumes are already one-hot (utils.py:296-298)
        return lblbatch.long()
    flat = lblbatch.reshape(lblbatch.shape[0], 1, lblbatch.shape[-2], lblbatch.shape[-1])
    ks = torch.arange(nlabels, device=lblbatch.device, dtype=flat.dtype).view(1, nlabels, 1, 1)
    return (flat == ks).long()


# ------------------------------------------------------------------------------------------------ model-side helpers
def truncated_normal_(tensor, mean=0, std=1):
    """reference utils.py:69-75: first of four N(0,1) draws inside (-2,2), scaled."""
    draws = tensor.new_empty(tuple(tensor.shape) + (4,)).normal_()
    ok = (draws < 2) & (draws > -2)
    first = ok.max(-1, keepdim=True)[1]
    tensor.data.copy_(draws.gather(-1, first).squeeze(-1))
    tensor.data.mul_(std).add_(mean)


def init_weights(m):
    """reference utils.py:78-83: Kaiming-normal weights, truncated-normal(0, 1e-3) bias for conv layers."""
    if type(m) == nn.Conv2d or type(m) == nn.ConvTranspose2d:
        nn.init.kaiming_normal_(m.weight, mode='fan_in', nonlinearity='relu')
        truncated_normal_(m.bias, mean=0, std=0.001)


def init_weights_orthogonal_normal(m):
    """reference utils.py:86-90."""
    if type(m) == nn.Conv2d or type(m) == nn.ConvTranspose2d:
        nn.init.orthogonal_(m.weight)
        truncated_normal_(m.bias, mean=0, std=0.001)


def l2_regularisation(m):
    """reference utils.py:93-101: sum of the parameters' 2-norms (one multi-tensor norm instead of a launch each)."""
    params = list(m.parameters())
    if not params:
        return None
    return torch.stack(torch._foreach_norm(params, 2)).sum()


def normalise_image(image):
    """zero mean / unit std (reference utils.py:104-112)."""
    img = np.float32(np.array(image, copy=True))
    return np.divide(img - np.mean(img), np.std(img) + 1e-6)


def normalise_images(X):
    return np.stack([normalise_image(X[i, ...]) for i in range(X.shape[0])]).astype(np.float32)


def makefolder(folder):
    if not os.path.exists(folder):
        os.makedirs(folder)
        return True
    return False


def setup_logger(name, log_file, level=logging.INFO):
    handler = logging.FileHandler(log_file, mode='w')
    handler.setFormatter(logging.Formatter('%(asctime)s %(levelname)s %(message)s'))
    logger = logging.getLogger(name)
    logger.setLevel(level)
    logger.addHandler(handler)
    return logger


def convert_nhwc_to_nchw(tensor):
    return tensor.permute(0, 3, 1, 2)


def convert_nchw_to_nhwc(tensor):
    return tensor.permute(0, 2, 3, 1)


def show_tensor(tensor):
    raise NotImplementedError('matplotlib debugging helper of the reference (utils.py:250-268) is out of scope')


# ------------------------------------------------------------------------------------------------ out-of-scope helpers
_reference_utils = None


def _load_reference_utils():
    """Find the reference's own utils.py further down sys.path (the launcher keeps the reference checkout there) and
    load it under a private name, so cv2 augmentation / NIfTI helpers keep working for real-data runs."""
    global _reference_utils
    if _reference_utils is not None:
        return _reference_utils
    here = os.path.dirname(os.path.abspath(__file__))
    for p in sys.path:
        cand = os.path.join(p or '.', 'utils.py')
        if os.path.isfile(cand) and os.path.abspath(os.path.dirname(cand)) != here:
            spec = importlib.util.spec_from_file_location('_unetzoo_reference_utils', cand)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            _reference_utils = mod
            return mod
    return None


def __getattr__(name):
    if name.startswith('__'):
        raise AttributeError(name)
    ref = _load_reference_utils()
    if ref is not None and hasattr(ref, name):
        return getattr(ref, name)
    raise AttributeError("module 'utils' (UNet-Zoo B200 drop-in) has no attribute %r" % name)
