#!/usr/bin/env python
"""Benchmark of the B200 hot path: PHiSeg-7/5 LIDC-128^2 training images/s (BASELINE.json metric), with GED-100
evaluation images/s, the tensor-core / HBM rooflines of the kernels, and beside it the stock-PyTorch/cuDNN path on the
same GPU and the reference's CPU path.

    python bench.py --gpus N --steps K --warmup W            # one process per GPU (torchrun for N > 1)
    python bench.py --impl reference ...                      # the reference algorithm on the host cores (oracle port)
    python bench.py --impl torch-cuda ...                     # the reference algorithm as stock torch ops on cuda:0

One "step" = forward(training=True) + loss + backward + Adam.step on one synthetic LIDC-shaped batch of 12 images
per GPU (reference train_model.py:101-122, models/experiments/phiseg_7_5_12.py).  `value` is timed with the batch
resident in HBM (CUDA-graph replay of the whole step); `e2e` goes through the public TrainStep.step_host call with
pinned host buffers: H2D of the batch and D2H of the loss inside the timed region.  `extra` holds the other BASELINE
configurations (RevPHiSeg, ProbUNet, U-Net, PHISeg3D 4x128^3) measured with the same harness at the same N.
"""
import argparse
import collections
import json
import os
import subprocess
import sys
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
# forward conv GFLOP per image / volume; training = 3 x forward
FWD_FLOPS = {'revphiseg': 18.318e9, 'probunet': 13.598e9, 'unet': 6.958e9, 'phiseg3d': 8.146e12, 'revphiseg3d': 2.248e12}
TENSOR_KERNELS = ('conv_tc2_kernel', 'conv_tc_kernel', 'wgrad_tc2_kernel', 'wgrad_tc_kernel', 'wgrad_reduce_kernel',
                  'wgrad_reduce_batched_kernel')


def conv_forward_flops_per_image(net, hw=128):
    """Algorithmic conv FLOPs of one training forward per image: 2*Cout*H*W*Cin*k^2 per conv, counted on the
    modules forward(training=True) runs (posterior + prior + likelihood; SURVEY.md 8d: 33.465 GFLOP for PHiSeg-7/5)."""
    import torch.nn as nn
    total = 0

    def level_of(name):
        parts = name.split('.')
        if parts[1] == 'contracting_path':
            return hw >> int(parts[2])
        if parts[1] == 'upsampling_path':                       # index i-1 used at latent level 4-i
            return hw >> (4 - (int(parts[2]) + 1) + 2)
        if parts[1] in ('sample_z_path', 'likelihood_ups_path'):
            return hw >> (4 - int(parts[2]) + 2)
        if parts[1] == 'likelihood_post_ups_path':
            base = hw >> ((4 - int(parts[2])) + 2)
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
        pw = [float(r[3]) if r[3].replace('.', '').isdigit() else 0.0 for r in rows]
        if pw:
            thr = min(pw) + 0.3 * (max(pw) - min(pw))
            loaded = [r for r, p in zip(rows, pw) if p >= thr] or rows
        else:
            loaded = rows
        sm = [float(r[1]) for r in loaded]
        mx = [float(r[2]) for r in loaded if r[2].replace('.', '').isdigit()]
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in loaded:
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'power_w_max': max(pw) if pw else None}


def synthetic_batches(n_batches, seed, volume=None, batch=BATCH):
    from b200 import synth
    out = []
    if volume is not None:               # BraTS-shaped volumes [1,4,S,S,S] + index labels
        for i in range(n_batches):
            vol, lab = synth.brats_like_batch(1, size=volume, seed=seed + i)
            out.append((vol.pin_memory(), lab.pin_memory(), None))
        return out
    for i in range(n_batches):
        patch, labels, mask = synth.lidc_like_batch(batch, seed=seed + i)
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


def cpu_eval_once(threads, n_samples):
    """the reference's N-sample evaluation of one image on the host cores (oracle port: network + GED + NCC)"""
    from oracle import metrics_oracle as mo
    from oracle import phiseg_oracle as po
    from oracle import synth
    from tests.keygrammar import phiseg_state_template
    torch.set_num_threads(threads)
    sd = synth.synth_state_dict(phiseg_state_template(FILTERS), seed=0)
    patch, labels, _ = synth.lidc_like_batch(1, seed=1000)
    masks = labels[0].permute(2, 0, 1).float()
    t0 = time.perf_counter()
    with torch.no_grad():
        eps = [torch.randn(s) for s in synth.phiseg_noise_shapes(n_samples)]
        out = po.phiseg_forward(sd, patch.repeat(n_samples, 1, 1, 1), masks[0][None, None].repeat(n_samples, 1, 1, 1), eps,
                                training=False)
        probs = po.accumulate_output(out['s'], use_softmax=True)
    t1 = time.perf_counter()
    pr = probs.argmax(1).numpy()
    mo.generalised_energy_distance(pr, masks.numpy(), 1, range(1, 2))
    mo.variance_ncc_dist(probs.numpy(), mo.convert_batch_to_onehot(masks.numpy()[:, None], 2))
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


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
        'config': workload_config(args.gpus),
        'cpu_baseline': {'value': val, 'unit': 'images/s', 'cores': threads, 'kind': 'port',
                         'sample': '%d timed B=12 training steps of the oracle port (torch CPU fp32, Adam)' % args.steps},
        'e2e': {'value': val, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def workload_config(world):
    """identical in both arms (the driver compares them)"""
    return {'workload': 'PHiSeg-7/5 training step (forward+loss+backward+Adam), LIDC-shaped 1x128x128, 4 annotators, '
                        'batch %d per GPU' % BATCH,
            'filters': FILTERS, 'global_batch': BATCH * world, 'parallelism': 'dp%d' % world}


def torch_cuda_block(steps, with_eval=True):
    """stock PyTorch / cuDNN on this GPU (oracle/torch_cuda_arm.py): the bar SURVEY.md 8(d) names"""
    from oracle import synth, torch_cuda_arm as arm
    from tests.keygrammar import phiseg_state_template
    tmpl = phiseg_state_template(FILTERS)
    out = {'what': 'reference algorithm as stock torch ops on this GPU (oracle restatement; no kernel of this repo)',
           'train': {}, 'eval': {}}
    for prec, graph in (('bf16', True), ('tf32', True), ('tf32', False)):
        key = '%s_%s' % (prec, 'graph' if graph else 'eager')
        try:
            r = arm.train_images_per_s(tmpl, BATCH, steps=steps, precision=prec, graph=graph)
            out['train'][key] = {'images_per_s': round(r['images_per_s'], 1), 'ms_per_step': round(r['ms_per_step'], 3)}
        except Exception as exc:
            out['train'][key] = {'error': repr(exc)[:200]}
    ok = [v['images_per_s'] for v in out['train'].values() if 'images_per_s' in v]
    out['train']['best_images_per_s'] = max(ok) if ok else None
    if with_eval:
        patch, labels, _ = synth.lidc_like_batch(1, seed=1000)
        try:
            r = arm.eval_images_per_s(tmpl, N_SAMPLES, labels[0], patch[0, 0], precision='bf16')
            out['eval']['bf16_graph'] = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()}
        except Exception as exc:
            out['eval']['bf16_graph'] = {'error': repr(exc)[:200]}
    torch.cuda.empty_cache()
    return out


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


def _short(name):
    name = name.replace('void ', '').replace('(anonymous namespace)::', '')
    if name.startswith('at::native::'):
        return 'torch:' + name[len('at::native::'):].split('<')[0]
    if 'nccl' in name.lower():
        return 'nccl:' + name.split('(')[0][:40]
    return name.split('(')[0].split('<')[0]


def cupti_kernel_table(step, replays=3):
    """Per-kernel busy time of one captured step from CUPTI activity records (torch.profiler): device timestamps of
    every kernel inside the graph replay.  -> (rows sorted by time, span_us) or (None, None)."""
    try:
        from torch.profiler import ProfilerActivity, profile
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(replays):
                step.step_device()
            torch.cuda.synchronize()
        evs = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
                      if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda t: t[0])
        if not evs:
            return None, None
        n = len(evs) // replays
        evs = evs[-n:]
        agg = collections.defaultdict(lambda: [0, 0.0])
        for s, e, name in evs:
            k = _short(name)
            agg[k][0] += 1
            agg[k][1] += e - s
        rows = sorted(({'kernel': k, 'launches': v[0], 'us': round(v[1], 1)} for k, v in agg.items()),
                      key=lambda r: -r['us'])
        return rows, evs[-1][1] - evs[0][0]
    except Exception as exc:        # profiling is side information, never fatal
        return [{'error': repr(exc)}], None


# algorithmic HBM bytes of the memory-bound families, from the arguments of their C-ABI calls (elements x bytes/element)
_HBM_BYTES = {
    'uz_bn_apply_train': ('bn_apply_train_kernel', lambda a: a[17] * a[18] * 4),          # read y, write a (bf16)
    'uz_bn_bwd_reduce_sums': ('bn_bwd_reduce_kernel', lambda a: a[7] * a[8] * 4),         # read dout, y
    'uz_bn_bwd_apply_train': ('bn_bwd_apply_train_kernel', lambda a: a[16] * a[17] * 6),  # read dout, y, write dy
    'uz_bn_bwd_fused': ('bn_bwd_cluster_kernel', lambda a: a[15] * a[16] * 6),            # the same in one launch
}


def single_stream_profile(net, device, hbm_peak):
    """The same training step captured on ONE stream (multi-stream overlap off: a kernel's busy time is then its own)
    and replayed under CUPTI: busy time of the tensor-core families and GB/s of the BatchNorm passes."""
    import models.phiseg as _mp
    from b200 import _lib, ops as _ops, train
    saved = (_mp._CONCURRENT, _ops._AUX_ENABLED, _lib.raw('uz_get_pdl')())
    _mp._CONCURRENT = False
    _ops.set_concurrency(False)
    _lib.call('uz_set_pdl', 0)      # programmatic dependent launch off: an early-launched kernel's record includes its wait
    bytes_by_kernel = collections.defaultdict(float)
    orig = _lib.call

    def logged(name, *a):
        if name in _HBM_BYTES and torch.cuda.is_current_stream_capturing():
            k, fn = _HBM_BYTES[name]
            bytes_by_kernel[k] += float(fn(a))
        return orig(name, *a)

    try:
        _lib.call = logged
        st = train.TrainStep(net, train.make_adam(net), BATCH, IMAGE, use_graph=True, dp=None, device=device)
        st.prepare(warmup=1)
        _lib.call = orig
        for _ in range(3):
            st.step_device()
        rows, span = cupti_kernel_table(st)
    finally:
        _lib.call = orig
        _mp._CONCURRENT = saved[0]
        _ops.set_concurrency(saved[1])
        _lib.call('uz_set_pdl', saved[2])
    del st
    if not rows or 'error' in rows[0]:
        return None
    tensor_us = sum(r['us'] for r in rows if any(r['kernel'].startswith(t) for t in TENSOR_KERNELS))
    hbm = {}
    for r in rows:
        if r['kernel'] in bytes_by_kernel:
            gbs = bytes_by_kernel[r['kernel']] / (r['us'] * 1e-6) / 1e9
            hbm[r['kernel']] = {'launches': r['launches'], 'us': r['us'], 'algorithmic_mb': round(bytes_by_kernel[r['kernel']] / 1e6, 1),
                                'gb_per_s': round(gbs, 1), 'frac_of_hbm_peak': round(gbs / hbm_peak, 3)}
    return {'rows': rows[:18], 'span_us': span, 'tensor_us': tensor_us, 'hbm': hbm,
            'launches': sum(r['launches'] for r in rows)}


def build_model(name, volume):
    from b200 import build
    if name == 'phiseg':
        return build.phiseg(FILTERS), BATCH, IMAGE
    if name == 'revphiseg':
        return build.phiseg(FILTERS, reversible=True), BATCH, IMAGE
    if name == 'probunet':
        return build.probunet(FILTERS, latent_dim=6), BATCH, IMAGE
    if name == 'unet':
        return build.unet([32, 64, 128, 192]), BATCH, IMAGE
    if name in ('phiseg3d', 'revphiseg3d'):
        image = (4, volume, volume, volume)
        return build.phiseg3d([32, 64, 128], 3, image, reversible=name == 'revphiseg3d'), 1, image
    raise KeyError(name)


def measure_training(name, args, rank, world, device, keep=False):
    """capture the training step of ``name`` (data parallel over ``world`` ranks), time K replays device-resident and K
    steps through TrainStep.step_host; -> dict (+ the step object when ``keep``)"""
    from b200 import _lib, dp as dpmod, synth, train
    net, batch_n, image = build_model(name, args.volume)
    net.load_state_dict(synth.synth_state_dict(net.state_dict(), seed=0))       # same weights on every rank
    net = net.to(device)
    opt = train.make_adam(net)
    dp = dpmod.GradientAllReduce(net.parameters()) if world > 1 else None
    step = train.TrainStep(net, opt, batch_n, image, use_graph=not args.no_graph, dp=dp, device=device)
    vol = args.volume if name.endswith('3d') else None
    batches = synthetic_batches(4, seed=1000 * (rank + 1), volume=vol)
    step.patch.copy_(batches[0][0])
    step.mask.copy_(batches[0][1])
    step.prepare(warmup=3)
    W = max(args.warmup, 3)
    for _ in range(W):
        step.step_device()
    torch.cuda.synchronize()
    l0 = _lib.raw('uz_launch_count')()
    ms_total = timed_region(lambda i: step.step_device(), args.steps, world, device)
    eager_launches = _lib.raw('uz_launch_count')() - l0
    launches = step.launches_per_step * args.steps if step.graph is not None else eager_launches
    ms_step = ms_total / args.steps
    losses = []

    def e2e_fn(i):
        pb, mb, _ = batches[i % len(batches)]
        losses.append(step.step_host(pb, mb))

    for i in range(2):
        e2e_fn(i)
    ms_e2e = timed_region(e2e_fn, args.steps, world, device) / args.steps
    unit = 'volumes/s' if name.endswith('3d') else 'images/s'
    fwd = FWD_FLOPS.get(name)
    if fwd is not None and name.endswith('3d'):
        fwd *= (args.volume / 128.0) ** 3
    out = {'value': world * batch_n / (ms_step / 1000.0), 'unit': unit, 'ms_per_step': ms_step,
           'e2e': {'value': world * batch_n / (ms_e2e / 1000.0), 'unit': unit,
                   'h2d_bytes_per_step': int(batches[0][0].numel() * 4 + batches[0][1].numel() * 4),
                   'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e,
                   'api': 'b200.train.TrainStep.step_host (%s)' % ('CUDA-graph replay' if step.graph is not None else 'eager')},
           'gpu_launches': int(launches), 'launches_per_step': int(step.launches_per_step or 0),
           'loss_last': losses[-1] if losses else None, 'batch_per_gpu': batch_n,
           'peak_mem_gb': round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
    if fwd is not None:
        out['algorithmic_tflops'] = 3.0 * fwd * batch_n * world / (ms_step / 1000.0) / 1e12
    if keep:
        return out, step, net, batches
    if dp is not None:
        dp.remove()
    del step, net, opt
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference', 'torch-cuda'])
    ap.add_argument('--no-graph', action='store_true', help='time eager steps instead of CUDA-graph replay')
    ap.add_argument('--skip-eval', action='store_true')
    ap.add_argument('--skip-cpu', action='store_true')
    ap.add_argument('--skip-extra', action='store_true', help='only the headline workload')
    ap.add_argument('--skip-torch', action='store_true', help='no stock-PyTorch-on-this-GPU baseline')
    ap.add_argument('--volume', type=int, default=128, help='edge of the cubic volume of the PHISeg3D configuration')
    ap.add_argument('--model', default='phiseg', choices=['phiseg', 'revphiseg', 'probunet', 'unet', 'phiseg3d', 'revphiseg3d'],
                    help='phiseg = the headline workload (with everything else in the line); any other name times only '
                         'that configuration and prints it as side information')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    if args.impl == 'reference':
        run_reference_arm(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the '
                         'CPU arm)')
    if args.impl == 'torch-cuda':
        if rank == 0:
            torch.cuda.set_device(0)
            print(json.dumps({'impl': 'torch-cuda', 'metric': METRIC, 'unit': 'images/s',
                              'torch_cuda_baseline': torch_cuda_block(args.steps)}))
        return

    from b200 import dp as dpmod
    rank, world, local = dpmod.init_from_env('nccl')
    device = torch.device('cuda', local)
    from b200 import _lib, synth, train
    torch.manual_seed(1234 + rank)

    if args.model != 'phiseg':
        r = measure_training(args.model, args, rank, world, device)
        if rank == 0:
            r['metric'] = '%s train %s (side information)' % (args.model, r['unit'])
            r['n_gpus'] = world
            print(json.dumps(r))
        sys.stdout.flush()
        if world > 1:
            torch.cuda.synchronize()
            os._exit(0)
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak_tf = float(peaks.get('bf16_tflops_sustained', 1400.0))
    peak_hbm = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'MEASURED_PEAKS.json (measured)' if peaks else 'fallback 1.4 PFLOP/s sustained, 6.65 TB/s'

    # ---- GED-100 evaluation FIRST, on the fixed seed-0 weights (before training mutates them): `world` images per call,
    # the 100 samples of every image sharded over the ranks, metrics of image i on rank i % world
    eval_block = None
    if not args.skip_eval:
        net_e, _, _ = build_model('phiseg', args.volume)
        net_e.load_state_dict(synth.synth_state_dict(net_e.state_dict(), seed=0))
        net_e = net_e.to(device)
        ev = train.EvalStep(net_e, N_SAMPLES, 2, shard=(rank, world), images_per_step=world)
        patch, labels, _ = synth.lidc_like_batch(world, seed=1000)          # same images on every rank
        img = patch[:, 0].contiguous().pin_memory()
        lab = labels.contiguous().pin_memory()
        for _ in range(2):
            res = ev.run_host(img, lab)
        k_eval = max(3, min(args.steps, 10))
        t_ms = timed_region(lambda i: ev.run_host(img, lab), k_eval, world, device) / k_eval
        # the same with the weight packing / BatchNorm folds hoisted out of the per-image call (a validation pass runs many
        # images on unchanged parameters, train_model.py:150-222): reported beside the per-call number, not instead of it
        ev_s = train.EvalStep(net_e, N_SAMPLES, 2, shard=(rank, world), images_per_step=world, static_weights=True)
        for _ in range(2):
            ev_s.run_host(img, lab)
        t_ms_static = timed_region(lambda i: ev_s.run_host(img, lab), k_eval, world, device) / k_eval
        del ev_s
        fg = None
        if rank == 0:
            # foreground fraction of the argmax masks of image 0 (a degenerate all-background set would skip the popcounts)
            nl, words = 1, (IMAGE[1] * IMAGE[2] + 31) // 32
            nloc = ev.n_local
            cnt = ev.flat[world * nloc * nl * words:world * nloc * nl * (words + 1)].view(world, nloc, nl)[0]
            fg = float(cnt.float().mean().item()) / (IMAGE[1] * IMAGE[2])
        eval_block = {'metric': 'PHiSeg GED-100 eval images/s (100 samples, 4 annotators, GED + NCC + Dice)',
                      'value': world * 1000.0 / t_ms, 'unit': 'images/s', 'ms_per_call': t_ms, 'images_per_call': world,
                      'value_static_weights': world * 1000.0 / t_ms_static, 'ms_per_call_static_weights': t_ms_static,
                      'ged': float(res[0, 0]), 'ncc': float(res[0, 1]), 'dice': [float(v) for v in res[0, 2:]],
                      'foreground_fraction_of_samples': fg, 'samples_per_rank': ev.counts or [N_SAMPLES],
                      'launches_per_call': int(getattr(ev, 'launches_per_step', 0)),
                      'weights': 'fixed synthetic seed-0 weights (evaluated before the training loop touches a model)',
                      'path': 'EvalStep.run_host: H2D images+labels, forward(training=False, replicate=n) on this rank\'s '
                              'share of the 100 copies of every image (encoders once per image), fused tail kernel over '
                              'the low-resolution level logits (accumulate + softmax + argmax -> bit-packed masks, per-pixel '
                              'sum p / sum log p), one all-gather of the masks + one all-reduce of the sums (N>1), GED / NCC / '
                              'Dice of image i on rank i % N, D2H of the scalars',
                      'algorithmic_bytes_per_image': int(sum(N_SAMPLES * 2 * (128 >> l) ** 2 * 4 for l in range(5)) + 4 * 128 * 128)}
        del ev, net_e
        torch.cuda.empty_cache()

    # ---- headline: PHiSeg-7/5 training step
    clk = ClockSampler(local)
    clk.__enter__()
    time.sleep(0.25)
    head, step, net, batches = measure_training('phiseg', args, rank, world, device, keep=True)
    for _ in range(max(0, int(1200.0 / max(head['ms_per_step'], 1e-3)) - 2 * args.steps)):
        step.step_device()               # keep the load up for >= ~1.2 s so the 100 ms clock sampler sees it
    torch.cuda.synchronize()
    clk.__exit__()
    fwd_flops = conv_forward_flops_per_image(net, IMAGE[1])
    train_flops = 3.0 * fwd_flops * BATCH

    # ---- rooflines (N = 1: a replay under data parallelism contains collectives every rank would have to join)
    roofline = None
    if world == 1:
        multi_rows, multi_span = cupti_kernel_table(step)
        prof = single_stream_profile(net, device, peak_hbm)
        if prof is not None:
            tc_ms = prof['tensor_us'] / 1000.0
            achieved = train_flops / (tc_ms / 1000.0) / 1e12
            traffic = None
            tsrc = os.path.join(ROOT, 'profiles', 'r02_ncu_summary.json')
            ncu = None
            if os.path.isfile(tsrc):
                try:
                    ncu = json.load(open(tsrc))
                    traffic = ncu.get('dominant_kernel', {}).get('dram_bytes_per_launch')
                except Exception:
                    ncu = None
            roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s',
                        'frac': achieved / peak_tf, 'traffic': traffic,
                        'kernel': 'conv_tc2_kernel / conv_tc_kernel (fwd + dgrad, the latter incl. the cluster-fused BatchNorm epilogue) + '
                                  'wgrad_tc2_kernel / wgrad_tc_kernel (+ wgrad_reduce_batched_kernel)',
                        'algorithmic_flops_per_step': train_flops, 'kernel_ms_per_step': tc_ms,
                        'peak_source': peak_src, 'single_stream_step_us': prof['span_us'],
                        'share_of_step': (prof['tensor_us'] / prof['span_us']) if prof['span_us'] else None,
                        'by_kernel_cupti_single_stream_step': prof['rows'],
                        'by_kernel_cupti_multistream_step': (multi_rows or [])[:12],
                        'hbm': dict(prof['hbm'], peak_gb_per_s=peak_hbm,
                                    note='BatchNorm passes: algorithmic bytes (elements x bytes per element from the call '
                                         'arguments) / CUPTI busy time of the family in the single-stream step'),
                        'how': 'algorithmic FLOPs (3 x forward conv FLOPs, SURVEY.md 8d) / summed CUPTI device durations of '
                               'the tensor-core kernels in a graph replay of the step captured on ONE stream (so a kernel\'s '
                               'time is its own); the headline value uses the overlapped multi-stream step',
                        'ncu': ncu}

    # ---- baselines (rank 0, N = 1 only)
    cpu_baseline = torch_cuda = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        threads = os.cpu_count() or 1
        times = cpu_train_steps(3, 1, threads)
        cpu_baseline = {'value': BATCH / float(np.mean(times)), 'unit': 'images/s', 'cores': threads, 'kind': 'port',
                        'sample': '3 timed + 1 warm-up B=12 PHiSeg-7/5 training steps of the oracle port '
                                  '(oracle/phiseg_oracle.py, torch CPU fp32 + Adam); the reference arm (--impl reference) '
                                  'times more steps of the same port and is the baseline the driver compares with'}
        if eval_block is not None:
            tn, tm = cpu_eval_once(threads, N_SAMPLES)
            eval_block['cpu_baseline'] = {'value': 1.0 / (tn + tm), 'unit': 'images/s', 'cores': threads, 'kind': 'port',
                                          'network_s': tn, 'metrics_s': tm,
                                          'sample': 'one image, 100 samples: oracle port forward(training=False) + numpy GED / NCC'}
    if rank == 0 and world == 1 and not args.skip_torch:
        del step
        torch.cuda.empty_cache()
        torch_cuda = torch_cuda_block(max(5, args.steps // 2), with_eval=not args.skip_eval)
        best = torch_cuda['train'].get('best_images_per_s')
        if best:
            torch_cuda['train']['b200_over_best'] = head['value'] / best
        ev_t = torch_cuda.get('eval', {}).get('bf16_graph', {})
        if eval_block is not None and 'images_per_s' in ev_t:
            torch_cuda['eval']['b200_over_torch'] = eval_block['value'] / ev_t['images_per_s']

    # ---- the other BASELINE configurations with the same harness
    extra = {}
    if not args.skip_extra:
        try:
            del step
        except NameError:
            pass
        del net
        torch.cuda.empty_cache()
        for name in ('revphiseg', 'probunet', 'unet', 'phiseg3d'):
            try:
                r = measure_training(name, args, rank, world, device)
                if 'algorithmic_tflops' in r:
                    r['tensor_frac_whole_step'] = r['algorithmic_tflops'] / world / peak_tf
                extra[name] = r
            except Exception as exc:
                extra[name] = {'error': repr(exc)[:300]}

    if rank == 0:
        clocks = clk.summary()
        act_mb = 15.0e6 * BATCH * 2 * 2 / 1e6      # ~15 M conv-output elements per image, y and a, bf16
        line = {
            'metric': METRIC, 'value': head['value'], 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': head['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': workload_config(world),
            'config_detail': {'cuda_graph': not args.no_graph, 'loss_last': head['loss_last'],
                              'l2': 'no flush needed: a step streams ~%.0f MB of activations (> 126 MB L2)' % act_mb},
            'e2e': head['e2e'], 'gpu_launches': head['gpu_launches'], 'launches_per_step': head['launches_per_step'],
            'clocks': clocks, 'roofline': roofline,
        }
        if cpu_baseline is not None:
            line['cpu_baseline'] = cpu_baseline
        if torch_cuda is not None:
            line['torch_cuda_baseline'] = torch_cuda
        if eval_block is not None:
            line['eval_ged100'] = eval_block
        if extra:
            line['extra'] = extra
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
