"""cProfile of eager PHiSeg-7/5 training steps (the path the unmodified train_model.py takes): where does the HOST time go?"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'unet-zoo_b200'))
import torch  # noqa: E402

import bench  # noqa: E402
from b200 import train  # noqa: E402
from oracle import synth  # noqa: E402
from tests.keygrammar import dropin_phiseg  # noqa: E402

net = dropin_phiseg(bench.FILTERS)
net.load_state_dict(synth.synth_state_dict(net.state_dict(), seed=0))
net = net.cuda()
opt = train.make_adam(net)
st = train.TrainStep(net, opt, bench.BATCH, bench.IMAGE, use_graph=False)
patch, labels, mask = synth.lidc_like_batch(bench.BATCH, seed=1)
st.patch.copy_(patch)
st.mask.copy_(mask)
for _ in range(5):
    st._body()
torch.cuda.synchronize()
t0 = time.time()
for _ in range(10):
    st._body()
torch.cuda.synchronize()
print('eager step: %.2f ms' % ((time.time() - t0) * 100))
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    st._body()
torch.cuda.synchronize()
pr.disable()
ps = pstats.Stats(pr)
ps.sort_stats('tottime').print_stats(28)
