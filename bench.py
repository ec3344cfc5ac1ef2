#!/usr/bin/env python
"""Benchmark of the B200 hot path: PHiSeg-7/5 LIDC-128^2 training images/s (BASELINE.json metric), with
GED-100 evaluation images/s, the tensor-core roofline of the conv kernels and the reference's CPU path beside it.

    python bench.py --gpus N --steps K --warmup W            # one process per GPU (torchrun for N > 1)
    python bench.py --impl reference ...                      # the reference algorithm on the host cores (oracle port)

One "step" = forward(training=True) + loss + backward + Adam.step on one synthetic LIDC-shaped batch of 12 images
per GPU (reference train_model.py:101-122, models/experiments/phiseg_7_5_12.py).  `value` is timed with the batch
resident in HBM (CUDA-graph replay of the whole step); `e2e` goes through the public TrainStep.step_host call with
pinned host buffers: H2D of the batch and D2H of the loss inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'unet-zoo_b200')
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

FILTERS = [32, 64, 128, 192, 192, 192, 192]       # models/experiments/phiseg_7_5_12.py:13
BATCH = 12                                         # per GPU (phiseg_7_5_12.py:30)
IMAGE = (1, 128, 128)
N_SAMPLES = 100                                    # GED-100 (BASELINE.json configs[3])
ANNOTATORS = 4
METRIC = 'PHiSeg-7/5 LIDC-128^2 train images/s'


def conv_forward_flops_per_image(net, hw=128):
    """Algorithmic conv FLOPs of one training forward per image: 2*Cout*H*W*Cin*k^2 per conv, counted on the
    modules forward(training=True) runs (posterior + prior + likelihood; SURVEY.md 8d: 33.465 GFLOP for PHiSeg-7/5)."""
    import torch.nn as nn
    total = 0
    res = {}

    def level_of(name):
        # resolution of each conv follows from the module path
        parts = name.split('.')
        if parts[1] == 'contracting_path':
            return hw >> int(parts[2])
        if parts[1] == 'upsampling_path':                       # index i-1 used at latent level 4-i
            i = int(parts[2]) + 1
            return hw >> (4 - i + 2)
        if parts[1] == 'sample_z_path':
            return hw >> (4 - int(parts[2]) + 2)
        if parts[1] == 'likelihood_ups_path':
            return hw >> (4 - int(parts[2]) + 2)
        if parts[1] == 'likelihood_post_ups_path':
            lvl = 4 - int(parts[2])
            base = hw >> (lvl + 2)
            return base * (2 if parts[3] == '1' else 4)
        if parts[1] == 'likelihood_post_c_path':
            return hw >> int(parts[2])
        if parts[1] == 's_layer':
            return hw >> (4 - int(parts[2]))
        raise KeyError(name)

    for name, m in net.named_modules():
        if isinstance(m, nn.Conv2d):
            if '.upsampling_path.4.' in name:                   # constructed but never called (phiseg.py:199)
                continue
            r = level_of(name)
            total += 2 * m.out_channels * r * r * m.in_channels * m.kernel_size[0] * m.kernel_size[1]
            res[name] = r
    return total


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def __enter__(self):
        # one long-lived nvidia-smi in loop mode (100 ms period) -- spawning one per sample is too slow for short regions
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is None:
            return
        try:
            self.proc.terminate()
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ''
        for line in (out or '').strip().splitlines():
            self.rows.append([c.strip() for c in line.split(',')])

    def summary(self):
        rows = [r for r in self.rows if len(r) > 8 and r[1].replace('.', '').isdigit()]
        # samples under load: power draw above the idle floor (first sample is taken before the load starts)
        pw = [float(r[3]) if r[3].replace('.', '').isdigit() else 0.0 for r in rows]
        if pw:
            thr = min(pw) + 0.3 * (max(pw) - min(pw))
            loaded = [r for r, p in zip(rows, pw) if p >= thr] or rows
        else:
            loaded = rows
        self.rows = loaded
        sm = [float(r[1]) for r in loaded]
        mx = [float(r[2]) for r in loaded if r[2].replace('.', '').isdigit()]
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            if len(r) > 8:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'power_w_max': max(pw) if pw else None}


def synthetic_batches(n_batches, seed, volume=None):
    from oracle import synth
    out = []
    if volume is not None:               # BraTS-shaped volumes [1,4,S,S,S] + index labels
        for i in range(n_batches):
            vol, lab = synth.brats_like_batch(1, size=volume, seed=seed + i)
            out.append((vol.pin_memory(), lab.pin_memory(), None))
        return out
    for i in range(n_batches):
        patch, labels, mask = synth.lidc_like_batch(BATCH, seed=seed + i)
        out.append((patch.pin_memory(), mask.pin_memory(), labels))
    return out


# ---------------------------------------------------------------------------------------------- reference (CPU) arm
def cpu_train_steps(steps, warmup, threads):
    """The reference algorithm (oracle port, fp32, stock torch CPU ops + Adam) on the host cores: one step = one
    B=12 training step of the same PHiSeg-7/5 configuration."""
    from oracle import phiseg_oracle as po
    from oracle import synth
    from tests.keygrammar import phiseg_state_template
    torch.set_num_threads(threads)
    sd = synth.synth_state_dict(phiseg_state_template(FILTERS), seed=0)
    params = [v.requires_grad_(True) for k, v in sd.items() if v.dtype == torch.float32 and 'running_' not in k]
    opt = torch.optim.Adam(params, lr=1e-3, weight_decay=1e-5)
    patch, labels, mask = synth.lidc_like_batch(BATCH, seed=100)
    times = []
    for it in range(warmup + steps):
        eps = [torch.randn(s) for s in synth.phiseg_noise_shapes(BATCH)]
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        out = po.phiseg_forward(sd, patch, mask, eps, training=True)
        loss = po.elbo(out, mask)['total']
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return times


def run_reference_arm(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    times = cpu_train_steps(args.steps, max(args.warmup, 1), threads)
    ms = 1000.0 * float(np.mean(times))
    val = BATCH / (ms / 1000.0)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'images/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': max(args.warmup, 1), 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'PHiSeg-7/5 train step, LIDC-shaped 1x128x128, batch 12 (one bounded sample = one step)',
                   'filters': FILTERS},
        'cpu_baseline': {'value': val, 'unit': 'images/s', 'cores': threads, 'kind': 'port',
                         'sample': '%d timed B=12 training steps of the oracle port (torch CPU fp32, Adam)' % args.steps},
        'e2e': {'value': val, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- B200 arm
def timed_region(fn, steps, world, device):
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def kernel_family_time(make_step, steps, world, device, lib):
    """In-situ GPU time of the tensor-core conv kernels inside the step: the SAME captured step is timed with CUDA
    events (graph replay, like the headline number) once complete, once with uz_conv_fwd launches elided and once with
    uz_conv_wgrad launches elided (uz_set_debug_flags 128 / 256); the differences are the kernel families' times.
    Values computed by the elided variants are garbage, so the caller restores the weights afterwards."""
    out = {}
    # kernel busy time, not exposed time: the multi-stream overlap of the product path is switched off for this
    # measurement so that a family's contribution to the step equals the sum of its launch durations
    import models.phiseg as _mp
    from b200 import ops as _ops
    saved_flags = (_mp._CONCURRENT, _ops._AUX_ENABLED)
    _mp._CONCURRENT = False
    _ops.set_concurrency(False)
    for name, flag in (('all', 0), ('without conv_tc (fwd+dgrad)', 128), ('without wgrad_tc', 256)):
        lib.call('uz_set_debug_flags', flag)
        try:
            st = make_step()
            st.prepare(warmup=1)
            for _ in range(2):
                st.step_device()
            out[name] = timed_region(lambda i: st.step_device(), steps, world, device) / steps
        finally:
            lib.call('uz_set_debug_flags', 0)
    _mp._CONCURRENT = saved_flags[0]
    _ops.set_concurrency(saved_flags[1])
    return out


def cupti_kernel_table(step, replays=3):
    """Per-kernel busy time of one captured step from CUPTI activity records (torch.profiler): name -> (launches, us).
    Complements the CUDA-event numbers: same step, device timestamps per kernel.  Returns None when CUPTI is unavailable."""
    try:
        import collections
        from torch.profiler import ProfilerActivity, profile
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(replays):
                step.step_device()
            torch.cuda.synchronize()
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        if not evs:
            return None
        agg = collections.defaultdict(lambda: [0, 0.0])
        for e in evs:
            name = e.name.replace('void ', '').replace('(anonymous namespace)::', '')
            if name.startswith('at::native::'):
                name = 'torch:' + name[len('at::native::'):].split('<')[0]
            elif 'conv_tc' in name or 'wgrad_tc' in name:
                name = name.split('(')[0]
            else:
                name = name.split('(')[0].split('<')[0]
            agg[name][0] += 1
            agg[name][1] += e.time_range.end - e.time_range.start
        rows = sorted(((k, v[0] // replays, round(v[1] / replays, 1)) for k, v in agg.items()), key=lambda t: -t[2])
        return [{'kernel': k, 'launches': c, 'us': u} for k, c, u in rows[:16]]
    except Exception as exc:        # profiling is side information, never fatal
        return [{'error': repr(exc)}]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-graph', action='store_true', help='time eager steps instead of CUDA-graph replay')
    ap.add_argument('--skip-eval', action='store_true')
    ap.add_argument('--skip-cpu', action='store_true')
    ap.add_argument('--debug-flags', type=int, default=0, help='uz_set_debug_flags for A/B measurements')
    ap.add_argument('--volume', type=int, default=128, help='edge of the cubic volume for --model phiseg3d')
    ap.add_argument('--model', default='phiseg', choices=['phiseg', 'revphiseg', 'probunet', 'unet', 'phiseg3d', 'revphiseg3d'],
                    help='phiseg = the headline workload; the others are reported as side information')
    args = ap.parse_args()

    from b200 import dp as dpmod
    if args.impl == 'reference':
        rank = int(os.environ.get('RANK', '0'))
        run_reference_arm(args, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the '
                         'CPU arm)')
    rank, world, local = dpmod.init_from_env('nccl')
    device = torch.device('cuda', local)
    from b200 import kern, train, _lib
    from tests.keygrammar import dropin_phiseg
    from oracle import synth

    if args.debug_flags:
        _lib.call('uz_set_debug_flags', args.debug_flags)
    torch.manual_seed(1234 + rank)
    batch_n, image = BATCH, IMAGE
    if args.model in ('phiseg3d', 'revphiseg3d'):
        from tests.keygrammar import dropin_phiseg3d
        batch_n, image = 1, (4, args.volume, args.volume, args.volume)     # 4 x 128^3, B = 1 per GPU
        net = dropin_phiseg3d([32, 64, 128], 3, image, reversible=args.model == 'revphiseg3d')
    elif args.model == 'phiseg':
        net = dropin_phiseg(FILTERS)
    elif args.model == 'revphiseg':
        net = dropin_phiseg(FILTERS, reversible=True)
    elif args.model == 'probunet':
        from models.probabilistic_unet import ProbabilisticUnet
        net = ProbabilisticUnet(input_channels=1, num_classes=2, num_filters=FILTERS, latent_dim=6, no_convs_fcomb=3)
    else:
        from models.unet import Unet
        net = Unet(1, 2, [32, 64, 128, 192])
    net.load_state_dict(synth.synth_state_dict(net.state_dict(), seed=0))       # same weights on every rank
    net = net.to(device)
    opt = train.make_adam(net, capturable=True)
    dp = dpmod.GradientAllReduce(net.parameters()) if world > 1 else None
    step = train.TrainStep(net, opt, batch_n, image, use_graph=not args.no_graph, dp=dp, device=device)
    batches = synthetic_batches(4, seed=1000 * (rank + 1), volume=args.volume if args.model.endswith('3d') else None)
    step.patch.copy_(batches[0][0])
    step.mask.copy_(batches[0][1])
    step.prepare(warmup=3)
    W = max(args.warmup, 3)
    for _ in range(W):
        step.step_device()
    torch.cuda.synchronize()

    # ---- device-resident throughput (value)
    l0 = _lib.raw('uz_launch_count')()
    clk = ClockSampler(local)
    clk.__enter__()
    time.sleep(0.25)                     # let the sampler start; the GPU stays busy from here to the end of the e2e loop
    for _ in range(3):
        step.step_device()
    ms_total = timed_region(lambda i: step.step_device(), args.steps, world, device)
    eager_launches = _lib.raw('uz_launch_count')() - l0
    gpu_launches = step.launches_per_step * args.steps if step.graph is not None else eager_launches
    ms_step = ms_total / args.steps
    value = world * batch_n / (ms_step / 1000.0)

    # ---- end to end through the public API (pinned host batch in, loss float out)
    losses = []

    def e2e_fn(i):
        pb, mb, _ = batches[i % len(batches)]
        losses.append(step.step_host(pb, mb))

    for i in range(2):
        e2e_fn(i)
    ms_e2e = timed_region(e2e_fn, args.steps, world, device) / args.steps
    for _ in range(max(0, int(1200.0 / max(ms_step, 1e-3)) - 2 * args.steps)):
        step.step_device()               # keep the load up for >= ~1.2 s so the 100 ms clock sampler sees it
    torch.cuda.synchronize()
    clk.__exit__()
    e2e = {'value': world * batch_n / (ms_e2e / 1000.0), 'unit': 'images/s',
           'h2d_bytes_per_step': int(batches[0][0].numel() * 4 + batches[0][1].numel() * 4), 'd2h_bytes_per_step': 4,
           'ms_per_step': ms_e2e, 'api': 'b200.train.TrainStep.step_host (CUDA-graph replay)' if step.graph is not None
           else 'b200.train.TrainStep.step_host (eager)'}

    # ---- roofline of the tensor-core conv kernels: events around every launch of instrumented eager steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak_tf = float(peaks.get('bf16_tflops_sustained', 1400.0))
    peak_src = 'MEASURED_PEAKS.json bf16_tflops_sustained (measured)' if peaks else 'fallback 1.4 PFLOP/s sustained'
    side_models = {'revphiseg': 18.318e9, 'probunet': 13.598e9, 'unet': 6.958e9,      # SURVEY.md 8d, forward GFLOP / image
                   'phiseg3d': 8.146e12 * (args.volume / 128.0) ** 3,                 # per 128^3 volume
                   'revphiseg3d': 2.248e12 * (args.volume / 128.0) ** 3}
    fwd_flops = conv_forward_flops_per_image(net, IMAGE[1]) if args.model == 'phiseg' else side_models[args.model]
    roofline = None      # filled after the evaluation block (the measurement overwrites the weights)

    # ---- GED-100 evaluation throughput (N=100 samples of one image, 4 annotators), samples sharded over ranks
    eval_block = None
    if not args.skip_eval and args.model == 'phiseg':
        ev = train.EvalStep(net, N_SAMPLES, 2, shard=(rank, world))
        labels = batches[0][2]
        img = batches[0][0][0, 0].contiguous().pin_memory()
        lab = labels[0].contiguous().pin_memory()
        for _ in range(2):
            ged, ncc = ev.run_host(img, lab)
        k_eval = max(3, min(args.steps, 10))
        t_ms = timed_region(lambda i: ev.run_host(img, lab), k_eval, world, device) / k_eval
        eval_block = {'metric': 'PHiSeg GED-100 eval images/s (100 samples, 4 annotators, GED + NCC)',
                      'value': 1000.0 / t_ms, 'unit': 'images/s', 'ms_per_image': t_ms, 'ged': ged, 'ncc': ncc,
                      'samples_per_rank': ev.counts or [N_SAMPLES],
                      'path': 'EvalStep.run_host: H2D image+labels, forward(training=False, replicate=n) on this '
                              'rank\'s share of the 100 copies (encoders once, latent sampling + likelihood per copy), '
                              'accumulate_output(softmax), all-gather of the class probabilities (N>1), argmax, GED, '
                              'NCC, D2H of two scalars'}
        net.train()

    # ---- roofline of the tensor-core conv kernels, in situ (differential graph replays)
    if args.model != 'phiseg':
        if rank == 0:
            what = ('%s [32,64,128] L=3, 4x%d^3 volumes/s' % (args.model, args.volume) if args.model.endswith('3d')
                    else '%s LIDC-128^2 train images/s' % args.model)
            print(json.dumps({'metric': what + ' (side information)', 'value': value,
                              'unit': 'images/s', 'n_gpus': world, 'ms_per_step': ms_step, 'e2e': e2e,
                              'gpu_launches': int(gpu_launches), 'peak_mem_gb': torch.cuda.max_memory_allocated() / 2 ** 30,
                              'algorithmic_tflops': 3.0 * fwd_flops * batch_n * world / (ms_step / 1000.0) / 1e12}))
        sys.stdout.flush()
        if world > 1:
            torch.cuda.synchronize()
            os._exit(0)
        return
    saved = {k: v.clone() for k, v in net.state_dict().items()}
    fam = kernel_family_time(lambda: train.TrainStep(net, train.make_adam(net), BATCH, IMAGE, use_graph=True, dp=None,
                                                     device=device), max(5, args.steps // 2), 1, device, _lib)
    net.load_state_dict(saved)
    # single-GPU only: under data parallelism a replay contains the NCCL all-reduces, which every rank would have to join
    by_kernel = cupti_kernel_table(step) if world == 1 else None
    t_all = fam['all']
    t_conv = max(t_all - fam['without conv_tc (fwd+dgrad)'], 1e-6)
    t_wgrad = max(t_all - fam['without wgrad_tc'], 1e-6)
    tc_ms = t_conv + t_wgrad
    train_flops = 3.0 * fwd_flops * BATCH
    achieved = train_flops / (tc_ms / 1000.0) / 1e12
    roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved / peak_tf,
                'traffic': None, 'kernel': 'conv_tc2_kernel / conv_tc_kernel (fwd + dgrad) + wgrad_tc_kernel',
                'algorithmic_flops_per_step': train_flops, 'kernel_ms_per_step': tc_ms,
                'ms_per_kernel_family': {'conv_tc (fwd+dgrad)': t_conv, 'wgrad_tc (+reduce)': t_wgrad},
                'step_ms': fam, 'peak_source': peak_src, 'share_of_step': tc_ms / t_all,
                'by_kernel_cupti_multistream_step': by_kernel,
                'how': 'CUDA-event time of the captured step, issued on ONE stream (multi-stream overlap off), minus the '
                       'same step with that kernel family elided (uz_set_debug_flags 128 / 256); no gradient '
                       'all-reduce; algorithmic FLOPs = 3 x forward conv FLOPs (SURVEY.md 8d).  The headline value '
                       'uses the overlapped multi-stream step.',
                'ncu_example': {'kernel': 'conv_tc2_kernel<64>, 128->128 @128^2, batch 12', 'us': 67.7,
                                'dram_bytes': 56.1e6, 'algorithmic_bytes': 100.9e6, 'tma_l2_to_sm_bytes': 396.4e6,
                                'tensor_pipe_active_frac': 0.42,
                                'source': 'profiles/r01_conv_v2_ncu_full_128to128_at128.md'}}

    # ---- the reference's CPU path beside it (rank 0, N = 1 only): bounded sample
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        threads = os.cpu_count() or 1
        times = cpu_train_steps(3, 1, threads)
        cpu_baseline = {'value': BATCH / float(np.mean(times)), 'unit': 'images/s', 'cores': threads, 'kind': 'port',
                        'sample': '3 timed + 1 warm-up B=12 PHiSeg-7/5 training steps of the oracle port '
                                  '(oracle/phiseg_oracle.py, torch CPU fp32 + Adam)'}

    if rank == 0:
        clocks = clk.summary()
        act_mb = 15.0e6 * BATCH * 2 * 2 / 1e6      # ~15 M conv-output elements per image, y and a, bf16
        line = {
            'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps, 'warmup': W,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
            'data': 'synthetic',
            'config': {'workload': 'PHiSeg-7/5 training step (forward+loss+backward+fused Adam), LIDC-shaped 1x128x128, '
                                   '4 annotators, batch %d per GPU' % BATCH,
                       'filters': FILTERS, 'global_batch': BATCH * world, 'parallelism': 'dp%d' % world,
                       'cuda_graph': step.graph is not None,
                       'l2': 'no flush needed: a step streams ~%.0f MB of activations (> 126 MB L2)' % act_mb,
                       'loss_last': losses[-1] if losses else None},
            'e2e': e2e, 'gpu_launches': int(gpu_launches), 'clocks': clocks, 'roofline': roofline,
        }
        if cpu_baseline is not None:
            line['cpu_baseline'] = cpu_baseline
        if eval_block is not None:
            line['eval_ged100'] = eval_block
        print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier()
        # captured graphs still reference the NCCL communicator: tearing the process group down can block, and there is
        # nothing left to flush -- leave without running destructors
        os._exit(0)


if __name__ == '__main__':
    main()
