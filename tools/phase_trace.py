"""%globaltimer phase timestamps of CTA 0 of the small-shape weight-gradient kernel (profiling build `make prof`):
where do the ~16 us of a tiny launch go?   UNETZOO_PRECISION=prof python tools/phase_trace.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'unet-zoo_b200'))
sys.path.insert(0, ROOT)
os.environ.setdefault('UNETZOO_PRECISION', 'prof')
import torch  # noqa: E402
from b200 import _lib, kern  # noqa: E402

NAMES = ['entry', 'before tmem_alloc', 'after tmem_alloc', 'after prologue sync', 'first TMA tile landed (MMA warp)',
         'last tile landed', 'accumulators ready (epilogue)', 'epilogue stores issued', 'after final sync',
         'epilogue: first tmem_ld done', 'epilogue: first block staged', 'epilogue: after staging barrier',
         'epilogue: first block copied out']


def main():
    trace = torch.zeros(16, dtype=torch.int64, device='cuda')
    _lib.call('uz_set_trace_buffer', trace.data_ptr())
    for (cin, cout, h) in ((192, 192, 4), (192, 192, 2), (256, 256, 8), (64, 64, 4)):
        x = torch.randn(12, h, h, cin, device='cuda').to(torch.bfloat16)
        dy = torch.randn(12, h, h, cout, device='cuda').to(torch.bfloat16)
        for rep in range(3):
            trace.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            kern.conv_wgrad(x, dy, 9, cin, cout)
            e1.record()
            torch.cuda.synchronize()
        t = trace.cpu().tolist()
        print('wgrad %d->%d @%dx%d: events %.1f us (wgrad + reduce)' % (cin, cout, h, h, e0.elapsed_time(e1) * 1e3))
        for i, n in enumerate(NAMES):
            if t[i]:
                print('   +%7.2f us  %s' % ((t[i] - t[0]) / 1e3, n))


def conv_traces():
    trace = torch.zeros(16, dtype=torch.int64, device='cuda')
    _lib.call('uz_set_trace_buffer', trace.data_ptr())
    for (cin, cout, h) in ((192, 192, 2), (192, 192, 8), (256, 256, 4), (64, 64, 8)):
        x = torch.randn(12, h, h, cin, device='cuda').to(torch.bfloat16)
        w = torch.randn(cout, cin, 3, 3, device='cuda') * 0.05
        wf, _ = kern.pack_conv_weight(w, need_dgrad=False)
        for rep in range(3):
            trace.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            kern.conv_fwd(x, wf, stats=True)
            e1.record()
            torch.cuda.synchronize()
        t = trace.cpu().tolist()
        print('conv fwd %d->%d @%dx%d (generic kernel): events %.1f us' % (cin, cout, h, h, e0.elapsed_time(e1) * 1e3))
        for i, n in enumerate(NAMES):
            if t[i]:
                print('   +%7.2f us  %s' % ((t[i] - t[0]) / 1e3, n))


if __name__ == '__main__':
    conv_traces()
    main()
