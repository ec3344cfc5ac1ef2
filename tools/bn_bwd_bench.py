"""In-graph duration of the BatchNorm+ReLU backward per shape: the cluster kernel (one launch) against the reduce + apply
pair.   python tools/bn_bwd_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'unet-zoo_b200'))
import torch  # noqa: E402

from b200 import kern  # noqa: E402

SHAPES = [(48, 192), (192, 192), (768, 192), (768, 256), (3072, 192), (3072, 256), (12288, 192), (12288, 128), (49152, 64),
          (49152, 192), (196608, 32)]
REP = 20


def timed(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REP):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (10 * REP)


def main():
    dev = 'cuda'
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    for npix, c in SHAPES:
        y = torch.randn(1, 1, npix, c, device=dev).to(torch.bfloat16)
        d = torch.randn(1, 1, npix, c, device=dev).to(torch.bfloat16)
        scale, shift = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev) * 0.1
        gamma, mean, invstd = torch.ones(c, device=dev), torch.zeros(c, device=dev), torch.ones(c, device=dev)
        out = {}
        for fused in (True, False):
            if fused and not kern._lib.raw('uz_bn_bwd_fused_supported')(npix, c):
                continue
            kern._BN_BWD_FUSED = fused
            out[fused] = timed(lambda: kern.bn_relu_bwd_train(d, y, scale, shift, gamma, mean, invstd))
        print('npix %6d C %3d   cluster %s us   reduce+apply %6.2f us' %
              (npix, c, ('%6.2f' % out[True]) if True in out else '   n/a', out[False]), flush=True)


if __name__ == '__main__':
    main()
