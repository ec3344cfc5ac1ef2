"""Helpers for the GPU parity tests: load the drop-in package (unet-zoo_b200/ is a sys.path root)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'unet-zoo_b200')
if PKG not in sys.path:
    sys.path.insert(0, PKG)


def kern():
    from b200 import kern as k
    return k


def bf16r(x):
    return x.to(torch.bfloat16).to(torch.float32)


def to_nhwc(x_nchw):
    """fp32 NCHW -> bf16 NHWC tensor (contiguous)"""
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def to_nchw(x_nhwc):
    return x_nhwc.float().permute(0, 3, 1, 2).contiguous()


def rel_err(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))
