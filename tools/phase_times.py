"""Time the phases of the PHiSeg-7/5 B=12 training step as separately captured CUDA graphs (B200)."""
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

dev = torch.device('cuda')
net = dropin_phiseg(bench.FILTERS)
net.load_state_dict(synth.synth_state_dict(net.state_dict(), seed=0))
net = net.cuda().train()
opt = train.make_adam(net)
patch, labels, mask = synth.lidc_like_batch(bench.BATCH, seed=1)
patch, mask = patch.cuda(), mask.cuda()


def bodies():
    def fwd_only():
        with torch.no_grad():
            net.forward(patch, mask, training=True)

    def fwd_loss():
        with torch.no_grad():
            net.forward(patch, mask, training=True)
            net.loss(mask)

    def posterior_only():
        with torch.no_grad():
            net.posterior(patch, mask)

    def fwd_bwd():
        opt.zero_grad(set_to_none=True)
        net.forward(patch, mask, training=True)
        net.loss(mask).backward()

    def full():
        opt.zero_grad(set_to_none=True)
        net.forward(patch, mask, training=True)
        net.loss(mask).backward()
        opt.step()

    return [('posterior forward only', posterior_only), ('forward (no grad)', fwd_only), ('forward + loss (no grad)', fwd_loss),
            ('forward + loss + backward', fwd_bwd), ('full step', full)]


for name, fn in bodies():
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
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
    print('%-28s %.3f ms' % (name, e0.elapsed_time(e1) / 10))
