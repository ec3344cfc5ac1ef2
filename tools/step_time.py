"""Step time of the captured PHiSeg-7/5 training step (product multi-stream configuration and single stream) for A/B
runs of environment knobs (UZ_CARVEOUT, UNETZOO_FUSE_BN_BWD, ...).   python tools/step_time.py [--steps 30] [--model phiseg]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'unet-zoo_b200'))
import torch  # noqa: E402

import bench  # noqa: E402
from b200 import ops, train  # noqa: E402
import models.phiseg as mp  # noqa: E402
from oracle import synth  # noqa: E402
from tests.keygrammar import dropin_phiseg  # noqa: E402


def run(multi, steps, reversible=False):
    dev = torch.device('cuda', 0)
    net = dropin_phiseg(bench.FILTERS, reversible=reversible)
    net.load_state_dict(synth.synth_state_dict(net.state_dict(), seed=0))
    net = net.to(dev)
    mp._CONCURRENT = multi
    ops.set_concurrency(multi)
    st = train.TrainStep(net, train.make_adam(net), bench.BATCH, bench.IMAGE, use_graph=True, device=dev)
    b = bench.synthetic_batches(1, seed=1)
    st.patch.copy_(b[0][0])
    st.mask.copy_(b[0][1])
    st.prepare(warmup=2)
    for _ in range(5):
        st.step_device()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        st.step_device()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, st.launches_per_step, float(st.loss)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--model', default='phiseg')
    ap.add_argument('--tag', default='')
    ap.add_argument('--debug-flags', type=int, default=0, help='profiling knobs (needs UNETZOO_PRECISION=prof)')
    ap.add_argument('--multi-only', action='store_true')
    args = ap.parse_args()
    if args.debug_flags:
        from b200 import _lib
        _lib.call('uz_set_debug_flags', args.debug_flags)
    out = {'tag': args.tag, 'env': {k: v for k, v in os.environ.items() if k.startswith('UZ_') or k.startswith('UNETZOO_')}}
    for multi in ((True,) if args.multi_only else (True, False)):
        ms, launches, loss = run(multi, args.steps, reversible=args.model == 'revphiseg')
        out['multi_stream' if multi else 'single_stream'] = {'ms': round(ms, 3), 'images_per_s': round(bench.BATCH / ms * 1e3, 1),
                                                            'launches': launches, 'loss': loss}
    print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
