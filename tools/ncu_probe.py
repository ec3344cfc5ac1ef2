"""Launches single instances of chosen kernels (after one warm-up each) for `ncu --set full` captures:
    ncu --set full --clock-control none --import-source on -o gpurun_out/probe python tools/ncu_probe.py --what wgrad_tiny,conv_tiny
Each probe prints a marker with the number of library launches it made so captures can be matched to shapes."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'unet-zoo_b200'))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from b200 import kern  # noqa: E402

B = 12


def act(h, c, n=B):
    return torch.randn(n, h, h, c, device='cuda').to(torch.bfloat16)


def conv(cin, cout, h, stats=True):
    x = act(h, cin)
    w = torch.randn(cout, cin, 3, 3, device='cuda') * 0.05
    wf, _ = kern.pack_conv_weight(w, need_dgrad=False)
    return lambda: kern.conv_fwd(x, wf, stats=stats)


def wgrad(cin, cout, h):
    x, dy = act(h, cin), act(h, cout)
    return lambda: kern.conv_wgrad(x, dy, 9, cin, cout)


def bn(c, h):
    y, d = act(h, c), act(h, c)
    g = torch.ones(c, device='cuda')
    b = torch.zeros(c, device='cuda')
    rm, rv = torch.zeros(c, device='cuda'), torch.ones(c, device='cuda')
    sums = torch.zeros(2 * c, device='cuda')
    sums[:c] = y.float().sum((0, 1, 2))
    sums[c:] = (y.float() ** 2).sum((0, 1, 2))

    def run():
        a, sc, sh, mean, invstd = kern.bn_apply_train(y, sums, B * h * h, g, b, rm, rv)
        kern.bn_relu_bwd_train(d, y, sc, sh, g, mean, invstd)
    return run


def eval_tail():
    """the fused evaluation tail at GED-100 size: 100 samples, 2 classes, 128^2, low-resolution level logits"""
    n, C, H = 100, 2, 128
    levels = [torch.randn(n, C, H >> l, H >> l, device='cuda') for l in range(5)]
    factors = [1 << l for l in range(5)]
    gts = (torch.rand(4, H, H, device='cuda') < 0.05).to(torch.uint8)

    def run():
        bits, cnts, sums = kern.eval_sample_stats(levels, factors, n, 1, C, (H, H), [1])
        kern.ged_from_bits(bits[0], cnts[0], gts, [1], H * H)
        kern.ncc_dice_from_sums(sums[0], gts, n, dice_annotator=0)
    return run


def heads():
    """latent head, logits layer and residual cross-entropy at batch 12 (forward + backward kernels)"""
    feat = act(32, 192)
    wmu, wsig = torch.randn(2, 192, device='cuda') * 0.05, torch.randn(2, 192, device='cuda') * 0.05
    bmu, bsig = torch.zeros(2, device='cuda'), torch.zeros(2, device='cuda')
    eps = torch.randn(B, 2, 32, 32, device='cuda')
    f128 = act(128, 128)
    ws, bs = torch.randn(2, 128, device='cuda') * 0.05, torch.zeros(2, device='cuda')
    target = (torch.rand(B, 1, 128, 128, device='cuda') < 0.05).float()

    def run():
        mu, sigma, z = kern.head_fwd(feat, wmu, bmu, wsig, bsig, eps)
        kern.head_bwd(feat, wmu, wsig, eps, sigma, torch.ones_like(mu), torch.ones_like(mu), torch.ones_like(mu))
        s0 = kern.slayer_fwd(f128, ws, bs, 1)
        kern.slayer_bwd(torch.ones_like(s0), f128, ws, 1)
        s_list = [s0] + [torch.randn_like(s0) for _ in range(4)]
        kern.residual_ce(s_list, target, need_grad=True, upstream=torch.ones(1, device='cuda'))
        kern.kl_fwd(mu, sigma, mu * 0.9, sigma * 1.1, 4.0)
    return run


def memops():
    x64, x128 = act(64, 64), act(128, 32)

    def run():
        kern.avgpool2_fwd(x128)
        kern.upsample2x_fwd(x64, True)
        kern.upsample2x_bwd(act(128, 64), True)
    return run


def fused_small(h):
    """conv + BatchNorm + ReLU in one cluster launch (uz_conv_bn_act_fused), 192 -> 192 at h x h x 12"""
    x = act(h, 192)
    w = torch.randn(192, 192, 3, 3, device='cuda') * 0.05
    wf, _ = kern.pack_conv_weight(w, need_dgrad=False)
    z, o = torch.zeros(192, device='cuda'), torch.ones(192, device='cuda')
    rm, rv = torch.zeros(192, device='cuda'), torch.ones(192, device='cuda')
    return lambda: kern.conv_bn_act_fused(x, wf, z, o, z, rm, rv)


def bn_cluster(c, h):
    """BatchNorm + ReLU backward in one cluster launch (uz_bn_bwd_fused)"""
    y, d = act(h, c), act(h, c)
    g, b = torch.ones(c, device='cuda'), torch.zeros(c, device='cuda')
    return lambda: kern.bn_relu_bwd_train(d, y, g, b, g, b, g)


def optimizer():
    """fused Adam + bf16 weight packing (uz_adam_pack_step) on 16 layers of 192 -> 192 x 3 x 3, gradients from split-K
    slabs reduced by the batched kernel (uz_wgrad_reduce_batched) for half of them"""
    from b200.optim import FusedAdam
    ws = [torch.nn.Parameter(torch.randn(192, 192, 3, 3, device='cuda') * 0.05) for _ in range(16)]
    pk = kern.WeightPacker(ws)
    opt = FusedAdam(ws, lr=1e-3, weight_decay=1e-5)
    opt.attach_packer(pk)
    x, dy = act(16, 192), act(16, 192)

    def run():
        for k, w in enumerate(ws):
            w.grad = kern.conv_wgrad(x, dy, 9, 192, 192, defer=True).view(192, 192, 3, 3)
        kern.wgrad_reducer.flush()
        opt.step()
    return run


def kl_hier():
    lv = []
    for l in range(5):
        r = 4 << l
        mk = lambda pos: (torch.rand(B, 2, r, r, device='cuda') + 0.1) if pos else torch.randn(B, 2, r, r, device='cuda')
        lv.append((mk(False), mk(True), mk(False), mk(True)))
    up = torch.ones(1, device='cuda')

    def run():
        kern.kl_hierarchy_fwd(lv, [4.0 ** i for i in range(5)], 1.0)
        kern.kl_hierarchy_bwd(lv, [4.0 ** i for i in range(5)], 1.0, up)
    return run


PROBES = {
    'fused_8x8': lambda: fused_small(8),
    'fused_2x2': lambda: fused_small(2),
    'bn_cluster_16': lambda: bn_cluster(192, 16),
    'bn_cluster_4': lambda: bn_cluster(192, 4),
    'optimizer': optimizer,
    'kl_hier': kl_hier,
    'eval_tail': eval_tail,
    'heads': heads,
    'memops': memops,
    'wgrad_tiny': lambda: wgrad(192, 192, 4),
    'wgrad_small': lambda: wgrad(192, 192, 16),
    'wgrad_mid': lambda: wgrad(128, 128, 32),
    'conv_tiny': lambda: conv(192, 192, 2),
    'conv_small': lambda: conv(192, 192, 16),
    'conv_c32': lambda: conv(32, 32, 128),
    'conv_big': lambda: conv(128, 128, 128),
    'bn_c32': lambda: bn(32, 128),
    'bn_c128': lambda: bn(128, 128),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--what', default=','.join(PROBES))
    args = ap.parse_args()
    for name in args.what.split(','):
        fn = PROBES[name]()
        fn()                      # warm-up (also captured; the second instance is the warm one)
        torch.cuda.synchronize()
        fn()
        torch.cuda.synchronize()
        print('probe', name, 'done', flush=True)


if __name__ == '__main__':
    main()
