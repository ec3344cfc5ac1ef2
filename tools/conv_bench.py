"""Per-shape timing of the tensor-core conv kernels (forward/dgrad kernel and wgrad kernel) on the PHiSeg-7/5 shape
census, CUDA events, L2 flushed between launches.  Usage (GPU box):
    python tools/conv_bench.py [--ncu-shape i]     # with --ncu-shape only that shape runs (for ncu -k filtering)
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'unet-zoo_b200'))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from b200 import kern  # noqa: E402

# (count, Cin, Cout, H, taps) at batch 12
CENSUS = [
    (1, 224, 128, 128, 9), (1, 128, 128, 128, 9), (1, 256, 192, 64, 9), (5, 192, 192, 32, 9), (1, 192, 192, 64, 9),
    (5, 32, 32, 128, 9), (5, 64, 64, 64, 9), (5, 128, 128, 32, 9), (4, 256, 256, 16, 9), (1, 320, 192, 32, 9),
    (6, 192, 192, 16, 9), (8, 192, 192, 8, 9), (1, 384, 192, 16, 9), (2, 32, 64, 64, 9), (2, 64, 128, 32, 9),
    (4, 256, 256, 8, 9), (2, 16, 32, 128, 9), (4, 192, 192, 4, 9), (4, 192, 192, 2, 9), (1, 16, 64, 32, 9),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=12)
    ap.add_argument('--reps', type=int, default=10)
    ap.add_argument('--ncu-shape', type=int, default=-1)
    ap.add_argument('--flags', type=int, default=0, help='uz_set_debug_flags (32 = generic kernel)')
    args = ap.parse_args()
    from b200 import _lib
    _lib.call('uz_set_debug_flags', args.flags)
    dev = 'cuda'
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot_f = tot_w = tot_flops = 0.0
    print('%-28s %10s %10s %10s %10s' % ('shape (Cin->Cout @H k)', 'fwd us', 'fwd TF/s', 'wgrad us', 'wgrad TF/s'))
    for idx, (cnt, cin, cout, h, taps) in enumerate(CENSUS):
        if args.ncu_shape >= 0 and idx != args.ncu_shape:
            continue
        ks = 3 if taps == 9 else 1
        x = torch.randn(args.batch, h, h, cin, device=dev).to(torch.bfloat16)
        dy = torch.randn(args.batch, h, h, cout, device=dev).to(torch.bfloat16)
        w = torch.randn(cout, cin, ks, ks, device=dev) * 0.05
        wf, _ = kern.pack_conv_weight(w, need_dgrad=False)
        flops = 2.0 * args.batch * h * h * cin * cout * taps
        tf, tw = [], []
        for r in range(args.reps if args.ncu_shape < 0 else 1):
            flush.zero_()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            torch.cuda.synchronize()
            e[0].record()
            kern.conv_fwd(x, wf, stats=True)
            e[1].record()
            flush.zero_()
            e[2].record()
            kern.conv_wgrad(x, dy, taps, cin, cout)
            e[3].record()
            torch.cuda.synchronize()
            tf.append(e[0].elapsed_time(e[1]) * 1e3)
            tw.append(e[2].elapsed_time(e[3]) * 1e3)
        f, wg = sorted(tf)[len(tf) // 2], sorted(tw)[len(tw) // 2]
        print('%dx %4d->%-4d @%-4d k%d      %10.1f %10.1f %10.1f %10.1f' %
              (cnt, cin, cout, h, ks, f, flops / f / 1e6, wg, flops / wg / 1e6))
        tot_f += cnt * f
        tot_w += cnt * wg
        tot_flops += cnt * flops
    print('census total: fwd %.0f us (%.1f TF/s), wgrad %.0f us (%.1f TF/s)' %
          (tot_f, tot_flops / tot_f / 1e6, tot_w, tot_flops / tot_w / 1e6))


if __name__ == '__main__':
    main()
