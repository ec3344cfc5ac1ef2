"""Warm, in-situ time of every kernel family of the PHiSeg-7/5 training step: the captured single-stream step is timed
with a family's launches elided (b200._lib.SKIP / uz_set_debug_flags) and the difference to the full step is that
family's busy time.  Elided variants compute garbage; only times are used.   python tools/family_times.py [--steps 20]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'unet-zoo_b200'))
import torch  # noqa: E402

import bench  # noqa: E402
from b200 import _lib, ops, train  # noqa: E402
import models.phiseg as mp  # noqa: E402
from oracle import synth  # noqa: E402
from tests.keygrammar import dropin_phiseg  # noqa: E402

FAMILIES = [
    ('conv persistent (fwd+dgrad)', 1024, ()),
    ('conv generic (fwd+dgrad)', 2048, ()),
    ('wgrad (mma + reduce)', 256, ()),
    ('wgrad reduce only', 4096, ()),
    ('bn forward (apply_train)', 0, ('uz_bn_apply_train',)),
    ('bn backward (reduce + apply)', 0, ('uz_bn_bwd_reduce_sums', 'uz_bn_bwd_apply_train')),
    ('slayer fwd', 0, ('uz_slayer_fwd',)),
    ('slayer bwd', 0, ('uz_slayer_bwd',)),
    ('heads fwd+bwd', 0, ('uz_head_fwd', 'uz_head_bwd')),
    ('kl + residual ce', 0, ('uz_kl_fwd', 'uz_kl_bwd', 'uz_residual_ce')),
    ('pool / upsample / copy', 0, ('uz_avgpool2_fwd', 'uz_avgpool2_bwd', 'uz_upsample2x_fwd', 'uz_upsample2x_bwd',
                                   'uz_copy_channels')),
    ('layout + input pack', 0, ('uz_nchw_to_nhwc', 'uz_nhwc_to_nchw', 'uz_input_pack')),
    ('weight pack', 0, ('uz_pack_conv_weights_batched',)),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--multi-stream', action='store_true')
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    net = dropin_phiseg(bench.FILTERS)
    net.load_state_dict(synth.synth_state_dict(net.state_dict(), seed=0))
    net = net.to(dev)
    if not args.multi_stream:
        mp._CONCURRENT = False
        ops.set_concurrency(False)
    batches = bench.synthetic_batches(1, seed=1)

    def timed(flag, skip, no_adam=False):
        _lib.call('uz_set_debug_flags', flag)
        _lib.SKIP = set(skip)
        try:
            opt = train.make_adam(net)
            if no_adam:
                opt.step = lambda *a, **k: None
            st = train.TrainStep(net, opt, bench.BATCH, bench.IMAGE, use_graph=True, device=dev)
            st.patch.copy_(batches[0][0])
            st.mask.copy_(batches[0][1])
            st.prepare(warmup=1)
            for _ in range(2):
                st.step_device()
            return bench.timed_region(lambda i: st.step_device(), args.steps, 1, dev) / args.steps
        finally:
            _lib.SKIP = set()
            _lib.call('uz_set_debug_flags', 0)

    full = timed(0, ())
    out = {'full_step_ms': full, 'families_ms': {}}
    for name, flag, skip in FAMILIES:
        out['families_ms'][name] = full - timed(flag, skip)
    out['families_ms']['fused Adam'] = full - timed(0, (), no_adam=True)
    out['families_ms']['unattributed (torch glue, launch gaps)'] = full - sum(
        v for k, v in out['families_ms'].items() if k != 'wgrad reduce only')
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
