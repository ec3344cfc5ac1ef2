"""Where does conv_tc_kernel's time go?  Times one big shape with parts of the kernel disabled (uz_set_debug_flags)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'unet-zoo_b200'))
import torch  # noqa: E402
from b200 import _lib, kern  # noqa: E402

dev = 'cuda'
for (cin, cout, h) in ((128, 128, 128), (192, 192, 64), (32, 32, 128)):
    x = torch.randn(12, h, h, cin, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, device=dev) * 0.05
    wf, _ = kern.pack_conv_weight(w, need_dgrad=False)
    out = kern.new_act(12, h, h, cout, dev)
    partial = torch.zeros((2, cout), dtype=torch.float32, device=dev)      # [2][Cout] statistics accumulators
    bias = torch.zeros(cout, device=dev)
    print('shape %d->%d @%d' % (cin, cout, h))
    for flags, name in ((0, 'v2 full'), (1, 'v2 no epilogue'), (2, 'v2 no MMA'), (3, 'v2 no MMA, no epilogue'),
                        (6, 'v2 B loads only'), (10, 'v2 A loads only'), (15, 'v2 barriers only'), (64, 'v2 full, no stats'),
                        (32, 'v1 full')):
        _lib.call('uz_set_debug_flags', flags & 63)
        stream = torch.cuda.Stream()
        reps = 10
        with torch.cuda.stream(stream):
            for r in range(2):
                _lib.call('uz_conv_fwd', x.data_ptr(), 12, h, h, cin, cin, wf.data_ptr(), cout, 9, out.data_ptr(), cout, None,
                          bias.data_ptr(), 0, None if flags == 64 else partial.data_ptr(), stream.cuda_stream)
            stream.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for r in range(reps):       # back-to-back launches: host latency hidden behind the queue
                _lib.call('uz_conv_fwd', x.data_ptr(), 12, h, h, cin, cin, wf.data_ptr(), cout, 9, out.data_ptr(), cout, None,
                          bias.data_ptr(), 0, None if flags == 64 else partial.data_ptr(), stream.cuda_stream)
            e1.record(stream)
            stream.synchronize()
        print('  %-24s %8.1f us' % (name, e0.elapsed_time(e1) * 1e3 / reps))
    _lib.call('uz_set_debug_flags', 0)
