"""Stock PyTorch / cuDNN on the same B200 (oracle/torch_cuda_arm.py) for every precision x launch mode, one JSON line.
    python tools/torch_cuda_bar.py [--steps 10] [--eval]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'unet-zoo_b200'))
import torch  # noqa: E402

import bench  # noqa: E402
from oracle import synth, torch_cuda_arm as arm  # noqa: E402
from tests.keygrammar import phiseg_state_template  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--eval', action='store_true')
    args = ap.parse_args()
    tmpl = phiseg_state_template(bench.FILTERS)
    out = {'train': {}, 'eval': {}}
    for prec in ('bf16', 'tf32', 'fp32'):
        for graph in (True, False):
            key = '%s_%s' % (prec, 'graph' if graph else 'eager')
            try:
                out['train'][key] = arm.train_images_per_s(tmpl, bench.BATCH, steps=args.steps, precision=prec, graph=graph)
            except Exception as exc:
                out['train'][key] = {'error': repr(exc)[:300]}
            print(key, out['train'][key], flush=True)
    if args.eval:
        patch, labels, _ = synth.lidc_like_batch(bench.BATCH, seed=1000)
        for prec in ('bf16', 'tf32'):
            key = '%s_graph' % prec
            try:
                out['eval'][key] = arm.eval_images_per_s(tmpl, bench.N_SAMPLES, labels[0], patch[0, 0], precision=prec)
            except Exception as exc:
                out['eval'][key] = {'error': repr(exc)[:300]}
            print('eval', key, out['eval'][key], flush=True)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
