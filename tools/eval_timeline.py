"""CUPTI kernel table of one GED-100 evaluation (EvalStep.run_host): python tools/eval_timeline.py"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'unet-zoo_b200'))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from b200 import train  # noqa: E402
from oracle import synth  # noqa: E402
from tests.keygrammar import dropin_phiseg  # noqa: E402

net = dropin_phiseg(bench.FILTERS)
net.load_state_dict(synth.synth_state_dict(net.state_dict(), seed=0))
net = net.cuda()
ev = train.EvalStep(net, 100, 2)
patch, labels, mask = synth.lidc_like_batch(1, seed=3)
img = patch[0, 0].contiguous().pin_memory()
lab = labels[0].contiguous().pin_memory()
for _ in range(3):
    ev.run_host(img, lab)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ev.run_host(img, lab)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.defaultdict(lambda: [0, 0.0])
t0 = min(e.time_range.start for e in evs)
t1 = max(e.time_range.end for e in evs)
for e in evs:
    name = e.name.replace('void ', '').replace('(anonymous namespace)::', '')
    name = ('torch:' + name[len('at::native::'):].split('<')[0]) if name.startswith('at::native::') else name.split('(')[0]
    agg[name][0] += 1
    agg[name][1] += e.time_range.end - e.time_range.start
print('span %.0f us, %d device activities, busy %.0f us' % (t1 - t0, len(evs), sum(v[1] for v in agg.values())))
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print('%5d %9.1f us  %s' % (c, v, k))
