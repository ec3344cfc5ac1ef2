"""One eager PHiSeg-7/5 B=12 training step (after warm-up) -- target for `ncu --metrics gpu__time_duration.sum`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'unet-zoo_b200'))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from b200 import train  # noqa: E402
from oracle import synth  # noqa: E402
from tests.keygrammar import dropin_phiseg  # noqa: E402

net = dropin_phiseg(bench.FILTERS)
net.load_state_dict(synth.synth_state_dict(net.state_dict(), seed=0))
net = net.cuda()
opt = train.make_adam(net, capturable=True)
step = train.TrainStep(net, opt, bench.BATCH, bench.IMAGE, use_graph=False)
patch, labels, mask = synth.lidc_like_batch(bench.BATCH, seed=1)
step.patch.copy_(patch)
step.mask.copy_(mask)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    step._body()
torch.cuda.synchronize()
print('done')
