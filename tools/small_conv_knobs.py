"""Where do the small-spatial conv launches spend their time?  20 identical launches captured in a CUDA graph (no host
latency), timed with parts of the kernel disabled (uz_set_debug_flags: 1 epilogue, 2 MMA, 4 / 8 activation / weight
loads)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'unet-zoo_b200'))
import torch  # noqa: E402
from b200 import _lib, kern  # noqa: E402

dev = 'cuda'
REPS = 20
for (cin, cout, h, n) in ((192, 192, 8, 12), (192, 192, 2, 12), (256, 256, 8, 12), (192, 192, 16, 12), (192, 192, 32, 12)):
    x = torch.randn(n, h, h, cin, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, device=dev) * 0.05
    wf, _ = kern.pack_conv_weight(w, need_dgrad=False)
    out = kern.new_act(n, h, h, cout, dev)
    stats = torch.zeros((2, cout), dtype=torch.float32, device=dev)
    bias = torch.zeros(cout, device=dev)
    print('shape %d->%d @%dx%d x%d  persistent=%d' % (cin, cout, h, h, n,
          _lib.raw('uz_conv_uses_persistent_kernel')(n, h, h, cin, cout, 9)))
    for flags, name in ((0, 'full'), (1, 'no epilogue'), (2, 'no MMA'), (3, 'no MMA, no epilogue'), (6, 'weight loads only'),
                        (10, 'activation loads only'), (15, 'barriers only')):
        _lib.call('uz_set_debug_flags', flags)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            def body():
                for _ in range(REPS):
                    _lib.call('uz_conv_fwd', x.data_ptr(), n, h, h, cin, cin, wf.data_ptr(), cout, 9, out.data_ptr(), cout,
                              None, bias.data_ptr(), 0, stats.data_ptr(), torch.cuda.current_stream().cuda_stream)
            body()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                body()
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
        print('  %-24s %6.2f us / launch' % (name, e0.elapsed_time(e1) * 1e3 / (5 * REPS)))
    _lib.call('uz_set_debug_flags', 0)
