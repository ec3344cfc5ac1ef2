"""Kernel timeline of graph-replayed training steps through torch.profiler (CUPTI): warm per-kernel durations and the idle
gaps between consecutive kernels.   python tools/timeline.py [--multi-stream] [--model phiseg]"""
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
from b200 import ops, train  # noqa: E402
import models.phiseg as mp  # noqa: E402
from oracle import synth  # noqa: E402
from tests.keygrammar import dropin_phiseg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--multi-stream', action='store_true')
    ap.add_argument('--debug-flags', type=int, default=0)
    ap.add_argument('--model', default='phiseg', choices=['phiseg', 'phiseg3d', 'revphiseg'])
    ap.add_argument('--list', default='', help='substring: print every launch of the matching kernels (grid, us)')
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'timeline.json'))
    ap.add_argument('--raw', default='', help='write every kernel of the last replay (start, duration, stream, grid) here')
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    from b200 import _lib
    if args.debug_flags:
        _lib.call('uz_set_debug_flags', args.debug_flags)
    if args.model == 'phiseg3d':
        from tests.keygrammar import dropin_phiseg3d
        batch_n, image = 1, (4, 128, 128, 128)
        net = dropin_phiseg3d([32, 64, 128], 3, image)
        batches = bench.synthetic_batches(1, seed=1, volume=128)
    else:
        batch_n, image = bench.BATCH, bench.IMAGE
        net = dropin_phiseg(bench.FILTERS, reversible=args.model == 'revphiseg')
        batches = bench.synthetic_batches(1, seed=1)
    net.load_state_dict(synth.synth_state_dict(net.state_dict(), seed=0))
    net = net.to(dev)
    if not args.multi_stream:
        mp._CONCURRENT = False
        ops.set_concurrency(False)
    st = train.TrainStep(net, train.make_adam(net), batch_n, image, use_graph=True, device=dev)
    st.patch.copy_(batches[0][0])
    st.mask.copy_(batches[0][1])
    st.prepare(warmup=2)
    for _ in range(5):
        st.step_device()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            st.step_device()
        torch.cuda.synchronize()
    if args.raw:
        trace = args.raw + '.trace.json'
        prof.export_chrome_trace(trace)
        tr = json.load(open(trace))
        rows = sorted((e['ts'], e['dur'], e['name'].replace('void ', '').replace('(anonymous namespace)::', '').split('(')[0][:60],
                       e.get('args', {}).get('stream'), e.get('args', {}).get('grid'), e.get('args', {}).get('block'),
                       e.get('args', {}).get('shared memory')) for e in tr['traceEvents'] if e.get('cat') == 'kernel')
        rows = rows[-(len(rows) // 3):]
        t0 = rows[0][0]
        json.dump([[round(r[0] - t0, 3)] + list(r[1:]) for r in rows], open(args.raw, 'w'))
        os.remove(trace)
    if args.list:
        trace = args.out + '.trace.json'
        prof.export_chrome_trace(trace)
        tr = json.load(open(trace))
        for pat in args.list.split(','):
            rows = [(e['ts'], e['dur'], e['name'], e.get('args', {})) for e in tr['traceEvents']
                    if e.get('cat') == 'kernel' and pat in e['name']]
            rows.sort()
            rows = rows[-(len(rows) // 3):]
            agg = collections.defaultdict(list)
            for ts, dur, name, a in rows:
                agg[(tuple(a.get('grid', [])), a.get('shared memory', 0))].append(dur)
            print('== %s: %d launches, %.1f us' % (pat, len(rows), sum(r[1] for r in rows)))
            for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:14]:
                print('grid %-16s smem %7s  n=%3d  total %8.1f us  avg %6.1f  min %6.1f max %6.1f' %
                      (k[0], k[1], len(v), sum(v), sum(v) / len(v), min(v), max(v)))
        os.remove(trace)
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
    print('cuda events', len(ks))
    if not ks:
        print('no CUPTI kernel records')
        return
    n = len(ks) // 3
    ks = ks[-n:]                       # last replay
    span = ks[-1][1] - ks[0][0]
    busy = collections.defaultdict(float)
    cnt = collections.Counter()
    for s, e, name in ks:
        short = name.replace('void ', '').replace('(anonymous namespace)::', '')
        if short.startswith('at::native::'):
            short = 'torch:' + short[len('at::native::'):].split('<')[0] + ('<fused adam>' if 'FusedOptimizer' in short else '')
        elif 'conv_tc2_kernel' in short or 'wgrad_tc_kernel' in short or 'head_' in short:
            short = short.split('(')[0]
        else:
            short = short.split('(')[0].split('<')[0]
        busy[short] += e - s
        cnt[short] += 1
    # union of busy intervals -> idle time
    cover, cur_e = 0.0, ks[0][0]
    gaps = []
    for s, e, _ in ks:
        if s > cur_e:
            gaps.append(s - cur_e)
        cover += max(0.0, e - max(s, cur_e))
        cur_e = max(cur_e, e)
    out = {'kernels': len(ks), 'span_us': span, 'covered_us': cover, 'idle_us': span - cover,
           'median_gap_us': sorted(gaps)[len(gaps) // 2] if gaps else 0.0, 'gaps': len(gaps),
           'sum_kernel_us': sum(busy.values()),
           'by_kernel': sorted(((k, cnt[k], round(v, 1)) for k, v in busy.items()), key=lambda t: -t[2])}
    json.dump(out, open(args.out, 'w'), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != 'by_kernel'}))
    for k, c, v in out['by_kernel'][:45]:
        print('%6d %10.1f us  %s' % (c, v, k))


if __name__ == '__main__':
    main()
