"""Launches single instances of chosen kernels (after one warm-up each) for `ncu --set full` captures:
    ncu --set full --clock-control none --import-source on -o gpurun_out/probe python tools/ncu_probe.py --what wgrad_tiny,conv_tiny
Each probe prints a marker with the number of library launches it made so captures can be matched to shapes."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'unet-zoo_b200'))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from b200 import kern  # noqa: E402

B = 12


def act(h, c, n=B):
    return torch.randn(n, h, h, c, device='cuda').to(torch.bfloat16)


def conv(cin, cout, h, stats=True):
    x = act(h, cin)
    w = torch.randn(cout, cin, 3, 3, device='cuda') * 0.05
    wf, _ = kern.pack_conv_weight(w, need_dgrad=False)
    return lambda: kern.conv_fwd(x, wf, stats=stats)


def wgrad(cin, cout, h):
    x, dy = act(h, cin), act(h, cout)
    return lambda: kern.conv_wgrad(x, dy, 9, cin, cout)


def bn(c, h):
    y, d = act(h, c), act(h, c)
    g = torch.ones(c, device='cuda')
    b = torch.zeros(c, device='cuda')
    rm, rv = torch.zeros(c, device='cuda'), torch.ones(c, device='cuda')
    sums = torch.zeros(2 * c, device='cuda')
    sums[:c] = y.float().sum((0, 1, 2))
    sums[c:] = (y.float() ** 2).sum((0, 1, 2))

    def run():
        a, sc, sh, mean, invstd = kern.bn_apply_train(y, sums, B * h * h, g, b, rm, rv)
        kern.bn_relu_bwd_train(d, y, sc, sh, g, mean, invstd)
    return run


PROBES = {
    'wgrad_tiny': lambda: wgrad(192, 192, 4),
    'wgrad_small': lambda: wgrad(192, 192, 16),
    'wgrad_mid': lambda: wgrad(128, 128, 32),
    'conv_tiny': lambda: conv(192, 192, 2),
    'conv_small': lambda: conv(192, 192, 16),
    'conv_c32': lambda: conv(32, 32, 128),
    'conv_big': lambda: conv(128, 128, 128),
    'bn_c32': lambda: bn(32, 128),
    'bn_c128': lambda: bn(128, 128),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--what', default=','.join(PROBES))
    args = ap.parse_args()
    for name in args.what.split(','):
        fn = PROBES[name]()
        fn()                      # warm-up (also captured; the second instance is the warm one)
        torch.cuda.synchronize()
        fn()
        torch.cuda.synchronize()
        print('probe', name, 'done', flush=True)


if __name__ == '__main__':
    main()
