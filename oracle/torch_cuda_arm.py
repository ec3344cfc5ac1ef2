"""Stock-PyTorch-on-the-same-GPU baseline arm -- TEST / MEASUREMENT INFRASTRUCTURE, never on the product path.

SURVEY.md 8(d) last row: "also time stock-PyTorch-on-B200 (the reference modules on cuda) -- that is the number the
kernels must beat".  The reference itself (/root/reference) does not exist on the GPU box, so this arm runs the oracle
restatement (oracle/phiseg_oracle.py: the same ATen / cuDNN calls the reference's nn.Conv2d / nn.BatchNorm2d /
F.interpolate / CrossEntropyLoss make, reference torchlayers.py:7-29, models/phiseg.py:414-537) on ``cuda``:

  precision  'tf32'  torch defaults on this image (cudnn.allow_tf32 = True): what `python train_model.py` would run
             'fp32'  TF32 off (strict fp32 FFMA convolutions)
             'bf16'  torch.autocast(bfloat16) + channels_last weights / inputs (the fastest stock configuration)
  graph      False   eager launches like the reference's loop (train_model.py:100-134)
             True    the whole step (noise draw, forward, ELBO, backward, fused capturable Adam) captured in ONE CUDA
                     graph and replayed -- the strongest stock baseline, same launch mechanism as b200.train.TrainStep

No kernel of libunetzoo_b200.so is involved.  bench.py prints these numbers as ``torch_cuda_baseline`` beside its own.
"""
import contextlib
import time

import numpy as np
import torch

from . import metrics_oracle as mo
from . import phiseg_oracle as po
from . import synth


@contextlib.contextmanager
def _precision(precision):
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, po.NATIVE_BN)
    po.NATIVE_BN = True
    torch.backends.cudnn.allow_tf32 = precision != 'fp32'
    torch.backends.cuda.matmul.allow_tf32 = precision != 'fp32'
    try:
        if precision == 'bf16':
            with torch.autocast('cuda', dtype=torch.bfloat16):
                yield
        else:
            yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, po.NATIVE_BN = saved


def _state(template, device, channels_last):
    sd = {k: v.to(device) for k, v in synth.synth_state_dict(template, seed=0).items()}
    if channels_last:
        for k, v in sd.items():
            if v.dim() == 4:
                sd[k] = v.contiguous(memory_format=torch.channels_last)
    return sd


def _time_steps(fn, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def train_images_per_s(template, batch, steps=10, warmup=3, precision='bf16', graph=True, device=None):
    """PHiSeg-7/5 training step (forward + ELBO + backward + Adam) with stock torch ops on ``device``.
    Returns dict(images_per_s, ms_per_step, loss)."""
    device = device or torch.device('cuda', torch.cuda.current_device())
    cl = precision == 'bf16'
    sd = _state(template, device, cl)
    params = [v.requires_grad_(True) for k, v in sd.items() if v.dtype == torch.float32 and 'running_' not in k]
    opt = torch.optim.Adam(params, lr=1e-3, weight_decay=1e-5, capturable=graph, fused=True)
    patch, _, mask = synth.lidc_like_batch(batch, seed=100)
    patch, mask = patch.to(device), mask.to(device)
    if cl:
        patch = patch.contiguous(memory_format=torch.channels_last)
    shapes = synth.phiseg_noise_shapes(batch)
    loss_out = torch.zeros((), device=device)

    def body():
        eps = [torch.randn(s, device=device) for s in shapes]
        opt.zero_grad(set_to_none=True)
        with _precision(precision):
            out = po.phiseg_forward(sd, patch, mask, eps, training=True)
            loss = po.elbo(out, mask)['total']
        loss.backward()
        opt.step()
        loss_out.copy_(loss.detach())

    if graph:
        s = torch.cuda.Stream(device=device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        run = g.replay
    else:
        run = body
    for _ in range(warmup):
        run()
    ms = _time_steps(run, steps)
    return {'images_per_s': batch / (ms / 1000.0), 'ms_per_step': ms, 'loss': float(loss_out)}


def eval_images_per_s(template, n_samples, labels, image, reps=3, precision='bf16', graph=True, device=None):
    """GED-N evaluation of ONE image like train_model.py:177-205: N copies -> forward(training=False) ->
    accumulate_output(softmax) -> argmax on the GPU with stock torch ops; GED + NCC with the reference's host algorithm
    (numpy restatement, oracle/metrics_oracle.py).  Returns dict(images_per_s, network_ms, metrics_ms, ged, ncc)."""
    device = device or torch.device('cuda', torch.cuda.current_device())
    cl = precision == 'bf16'
    sd = _state(template, device, cl)
    patch = image[None, None].repeat(n_samples, 1, 1, 1).to(device)
    masks = labels.permute(2, 0, 1).float()                      # [M,H,W]
    mask = masks[0][None, None].repeat(n_samples, 1, 1, 1).to(device)
    if cl:
        patch = patch.contiguous(memory_format=torch.channels_last)
    shapes = synth.phiseg_noise_shapes(n_samples)
    probs = torch.zeros((n_samples, 2) + tuple(image.shape), device=device)
    pred = torch.zeros((n_samples,) + tuple(image.shape), dtype=torch.int64, device=device)

    @torch.no_grad()
    def body():
        eps = [torch.randn(s, device=device) for s in shapes]
        with _precision(precision):
            out = po.phiseg_forward(sd, patch, mask, eps, training=False)
        p = po.accumulate_output([t.float() for t in out['s']], use_softmax=True)
        probs.copy_(p)
        pred.copy_(torch.argmax(p, dim=1))

    if graph:
        s = torch.cuda.Stream(device=device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        run = g.replay
    else:
        run = body
    run()
    net_ms = _time_steps(run, reps)
    t0 = time.perf_counter()
    pr, pb = pred.cpu().numpy(), probs.cpu().numpy()
    gt = masks.numpy()
    ged = mo.generalised_energy_distance(pr, gt, 1, range(1, 2))
    ncc = float(mo.variance_ncc_dist(pb, mo.convert_batch_to_onehot(gt[:, None], 2))[0])
    met_ms = 1000.0 * (time.perf_counter() - t0)
    return {'images_per_s': 1000.0 / (net_ms + met_ms), 'network_ms': net_ms, 'metrics_ms': met_ms, 'ged': ged,
            'ncc': ncc, 'foreground_frac': float(np.mean(pr != 0))}
