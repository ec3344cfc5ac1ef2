"""Per-LAYER in-graph kernel times of the captured PHiSeg training step (single stream): the C-ABI calls of one step are
logged with their shapes while the step is captured, then the CUPTI kernel records of a graph replay are matched to them
in launch order.  Unlike tools/conv_bench.py (isolated, cold launches) these are the durations inside the real step.

    python tools/layer_times.py [--out gpurun_out/layer_times.json]"""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'unet-zoo_b200'))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from b200 import _lib, kern, ops, train  # noqa: E402
import models.phiseg as mp  # noqa: E402
from oracle import synth  # noqa: E402
from tests.keygrammar import dropin_phiseg  # noqa: E402

# C-ABI entry point -> kernel-name substrings it launches, in order
FAMILY = {
    'uz_conv_fwd': ['conv_tc'],
    'uz_conv_wgrad': ['wgrad_tc', 'wgrad_reduce'],
    'uz_bn_apply_train': ['bn_apply_train'],
    'uz_bn_bwd_reduce_sums': ['bn_bwd_reduce'],
    'uz_bn_bwd_apply_train': ['bn_bwd_apply_train'],
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'layer_times.json'))
    ap.add_argument('--model', default='phiseg', choices=['phiseg', 'revphiseg'])
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    net = dropin_phiseg(bench.FILTERS, reversible=args.model == 'revphiseg')
    net.load_state_dict(synth.synth_state_dict(net.state_dict(), seed=0))
    net = net.to(dev)
    mp._CONCURRENT = False
    ops.set_concurrency(False)
    st = train.TrainStep(net, train.make_adam(net), bench.BATCH, bench.IMAGE, use_graph=True, device=dev)
    b = bench.synthetic_batches(1, seed=1)
    st.patch.copy_(b[0][0])
    st.mask.copy_(b[0][1])

    log = []
    orig_call = _lib.call
    state = {'on': False}

    def logged_call(name, *a):
        if state['on'] and name in FAMILY:
            log.append((name, a))
        return orig_call(name, *a)

    # log during the capture only (prepare() = warm-up steps + capture)
    real_graph = torch.cuda.graph

    class logging_graph(real_graph):
        def __enter__(self):
            state['on'] = True
            return super().__enter__()

        def __exit__(self, *e):
            state['on'] = False
            return super().__exit__(*e)

    _lib.call = logged_call
    kern._lib.call = logged_call
    torch.cuda.graph = logging_graph
    st.prepare(warmup=2)
    torch.cuda.graph = real_graph
    for _ in range(5):
        st.step_device()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            st.step_device()
        torch.cuda.synchronize()
    evs = sorted(((e.time_range.start, e.time_range.end - e.time_range.start, e.name) for e in prof.events()
                  if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda t: t[0])
    n = len(evs) // 3
    evs = evs[-n:]
    span = evs[-1][0] + evs[-1][1] - evs[0][0]
    # match in order, per family substring
    queues = collections.defaultdict(collections.deque)
    for ts, dur, name in evs:
        for fam in ('wgrad_reduce', 'wgrad_tc', 'conv_tc', 'bn_bwd_apply_train', 'bn_bwd_reduce', 'bn_apply_train'):
            if fam in name:
                queues[fam].append((dur, name))
                break
    rows = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for name, a in log:
        if name == 'uz_conv_fwd':
            # (x, n, h, w, cin, ldx, w, cout, taps, y, ldy, scale, shift, relu, partial, stream)
            n_, h, w, cin, cout, taps = a[1], a[2], a[3], a[4], a[7], a[8]
            kind = 'fwd' if (a[14] is not None or a[11] is not None or a[12] is not None) else 'dgrad'
            key = ('conv %s' % kind, '%d->%d @%dx%d k%d' % (cin, cout, h, w, 3 if taps == 9 else 1))
            flops = 2.0 * n_ * h * w * cin * cout * taps
        elif name == 'uz_conv_wgrad':
            # (x, ldx, dy, lddy, n, h, w, cin, cout, taps, cin_l, cout_l, work, dw, stream)
            n_, h, w, cin, cout, taps = a[4], a[5], a[6], a[7], a[8], a[9]
            key = ('wgrad', '%d->%d @%dx%d k%d' % (cin, cout, h, w, 3 if taps == 9 else 1))
            flops = 2.0 * n_ * h * w * cin * cout * taps
        elif name == 'uz_bn_apply_train':
            key = ('bn_apply', 'C=%d npix=%d' % (a[18], a[17]))
            flops = 0.0
        elif name == 'uz_bn_bwd_reduce_sums':
            key = ('bn_bwd_reduce', 'C=%d npix=%d' % (a[8], a[7]))
            flops = 0.0
        else:
            key = ('bn_bwd_apply', 'C=%d npix=%d' % (a[17], a[16]))
            flops = 0.0
        for i, fam in enumerate(FAMILY[name]):
            if not queues[fam]:
                raise SystemExit('kernel record queue %s ran dry at %s' % (fam, key))
            dur, kname = queues[fam].popleft()
            k2 = key if i == 0 else (key[0] + '_reduce', key[1])
            rows[k2][0] += 1
            rows[k2][1] += dur
            rows[k2][2] = flops if i == 0 else 0.0
    left = {k: len(v) for k, v in queues.items() if v}
    out = []
    for (kind, shape), v in rows.items():
        cnt, us, fl = v[0], v[1], v[2]
        out.append({'kind': kind, 'shape': shape, 'launches': cnt, 'us': round(us, 1), 'us_each': round(us / cnt, 2),
                    'tflops': round(fl * cnt / (us * 1e-6) / 1e12, 1) if fl else None})
    out.sort(key=lambda r: -r['us'])
    tot = collections.defaultdict(float)
    for r in out:
        tot[r['kind']] += r['us']
    res = {'span_us': span, 'kernels': len(evs), 'unmatched': left, 'by_kind_us': dict(tot), 'rows': out}
    json.dump(res, open(args.out, 'w'), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != 'rows'}))
    for r in out:
        print('%-18s %-28s n=%3d  %8.1f us  each %7.2f  %s' % (r['kind'], r['shape'], r['launches'], r['us'], r['us_each'],
                                                               ('%.0f TF/s' % r['tflops']) if r['tflops'] else ''))


if __name__ == '__main__':
    main()
