"""In-graph duration of conv + BatchNorm + ReLU (training) on the deepest levels: CUDA-core small-map kernel (default) vs
the tensor-core cluster kernel (UZ_CONV_SMALL=0).   python tools/conv_small_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'unet-zoo_b200'))
import torch  # noqa: E402
from b200 import kern  # noqa: E402

REP = 20
dev = 'cuda'
for (cin, cout, h, n) in ((192, 192, 2, 12), (192, 192, 4, 12), (256, 256, 4, 12), (64, 192, 2, 12)):
    x = torch.randn(n, h, h, cin, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, device=dev) * 0.05
    wf, wd = kern.pack_conv_weight(w, need_dgrad=True)
    bias, gamma, beta = torch.zeros(cout, device=dev), torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
    rm, rv = torch.zeros(cout, device=dev), torch.ones(cout, device=dev)
    for name, fn in (('conv+BN+ReLU', lambda: kern.conv_bn_act_fused(x, wf, bias, gamma, beta, rm, rv)),
                     ('conv (dgrad-like)', lambda: kern.conv_fwd(x, wf))):
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
        print('%d->%d @%dx%d x%d  %-18s %6.2f us' % (cin, cout, h, h, n, name, e0.elapsed_time(e1) * 1e3 / (10 * REP)), flush=True)
