"""Per-kernel parity of the CUDA path (through the C ABI) against plain torch fp32 ops / the CPU oracle on the same
(bf16-rounded where the kernel stores bf16) inputs.  Tolerances are written at each assert:
  * tensor-core convs: fp32 accumulation of bf16 products, output stored bf16  -> |err| <= 2^-8 |ref| + 1e-2*rms
  * fp32 heads / losses: 1e-4 relative (north-star: loss terms within 1e-3)
  * integer work (one-hot, argmax, IoU counts, GED): bit exact
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.gpu_util import bf16r, kern, rel_err, to_nchw, to_nhwc

pytestmark = pytest.mark.gpu
DEV = 'cuda'
# the torch reference must be true fp32: no TF32 in cuDNN / cuBLAS
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def _assert_bf16_close(got, ref, what=''):
    rms = float(ref.pow(2).mean().sqrt())
    err = (got - ref).abs()
    tol = ref.abs() * 2 ** -7 + 1e-2 * rms
    bad = int((err > tol).sum())
    assert bad == 0, '%s: %d / %d outside tolerance, max err %g (rms %g)' % (what, bad, err.numel(), float(err.max()), rms)


CONV_SHAPES = [
    # N, H, W, Cin, Cout, taps
    (2, 16, 16, 64, 64, 9),
    (3, 32, 32, 128, 192, 9),
    (12, 8, 8, 192, 192, 9),      # tile spans 2 images
    (12, 2, 2, 192, 192, 9),      # tile spans 32 images > batch: TMA zero-fills the overhang
    (5, 4, 4, 256, 256, 9),
    (2, 128, 128, 16, 32, 9),     # KC = 16 (32-byte swizzle): padded 1/3-channel inputs
    (2, 64, 64, 32, 64, 9),       # KC = 32 (64-byte swizzle)
    (2, 32, 32, 224, 128, 9),     # 7 x 32-channel K blocks
    (1, 16, 16, 320, 192, 9),
    (2, 16, 16, 64, 256, 9),
    (2, 32, 32, 224, 128, 1),     # 1x1
    (2, 32, 32, 16, 16, 9),       # dgrad shape towards a padded 2-channel z
    (2, 32, 32, 64, 48, 9),       # odd multiple of 16 output channels: 16-column epilogue of the persistent kernel
    (1, 16, 16, 32, 16, 9),
]


@pytest.mark.parametrize('N,H,W,Cin,Cout,taps', CONV_SHAPES)
def test_conv_fwd_matches_torch(N, H, W, Cin, Cout, taps):
    k = kern()
    ks = 3 if taps == 9 else 1
    x = bf16r(_rand(N, Cin, H, W, seed=1))
    w = bf16r(_rand(Cout, Cin, ks, ks, seed=2, scale=(2.0 / (Cin * taps)) ** 0.5))
    wf, wd = k.pack_conv_weight(w)
    y, partial = k.conv_fwd(to_nhwc(x), wf, stats=True)
    ref = F.conv2d(x, w, padding=ks // 2)
    _assert_bf16_close(to_nchw(y), ref, 'conv')
    # statistics of the stored values
    yq = to_nchw(y)
    s = partial.sum(0)
    torch.testing.assert_close(s[0], yq.sum((0, 2, 3)), rtol=1e-4, atol=1e-2)
    torch.testing.assert_close(s[1], yq.pow(2).sum((0, 2, 3)), rtol=1e-4, atol=1e-2)


def test_conv_fwd_affine_relu_and_channel_slices():
    k = kern()
    N, H, W, Cin, Cout = 2, 32, 32, 64, 128
    x = bf16r(_rand(N, Cin, H, W, seed=3))
    w = bf16r(_rand(Cout, Cin, 3, 3, seed=4, scale=0.05))
    scale = _rand(Cout, seed=5).abs() + 0.5
    shift = _rand(Cout, seed=6)
    wf, _ = k.pack_conv_weight(w, need_dgrad=False)
    # input lives in channels [32, 96) of a 128-wide buffer; output goes to channels [64, 192) of a 256-wide one
    xbuf = torch.zeros((N, H, W, 128), dtype=torch.bfloat16, device=DEV)
    xbuf[..., 32:96] = to_nhwc(x)
    ybuf = torch.full((N, H, W, 256), 7.0, dtype=torch.bfloat16, device=DEV)
    k.conv_fwd(xbuf[..., 32:96], wf, out=ybuf[..., 64:192], scale=scale, shift=shift, relu=True)
    ref = F.relu(F.conv2d(x, w, padding=1) * scale[None, :, None, None] + shift[None, :, None, None])
    _assert_bf16_close(to_nchw(ybuf[..., 64:192].contiguous()), ref, 'conv affine')
    assert float((ybuf[..., :64].float() - 7).abs().max()) == 0 and float((ybuf[..., 192:].float() - 7).abs().max()) == 0


@pytest.mark.parametrize('N,H,W,Cin,Cout', [(2, 32, 32, 64, 128), (12, 8, 8, 192, 192), (2, 64, 64, 16, 32),
                                            (1, 16, 16, 320, 192)])
def test_conv_dgrad_matches_autograd(N, H, W, Cin, Cout):
    k = kern()
    x = bf16r(_rand(N, Cin, H, W, seed=1)).requires_grad_(True)
    w = bf16r(_rand(Cout, Cin, 3, 3, seed=2, scale=(2.0 / (Cin * 9)) ** 0.5))
    dy = bf16r(_rand(N, Cout, H, W, seed=3))
    F.conv2d(x, w, padding=1).backward(dy)
    _, wd = k.pack_conv_weight(w)
    dx, _ = k.conv_fwd(to_nhwc(dy), wd)
    _assert_bf16_close(to_nchw(dx), x.grad, 'dgrad')


@pytest.mark.parametrize('N,H,W,Cin,Cout,taps', [
    (2, 32, 32, 64, 128, 9), (12, 8, 8, 192, 192, 9), (2, 64, 64, 16, 32, 9), (1, 16, 16, 320, 192, 9),
    (3, 16, 16, 256, 256, 9), (2, 32, 32, 224, 128, 9), (12, 2, 2, 192, 192, 9), (2, 128, 128, 32, 32, 9),
    (2, 16, 16, 384, 192, 9), (2, 32, 32, 224, 128, 1), (4, 32, 32, 16, 64, 9)])
def test_conv_wgrad_matches_autograd(N, H, W, Cin, Cout, taps):
    k = kern()
    ks = 3 if taps == 9 else 1
    x = bf16r(_rand(N, Cin, H, W, seed=1))
    w = _rand(Cout, Cin, ks, ks, seed=2).requires_grad_(True)
    dy = bf16r(_rand(N, Cout, H, W, seed=3))
    F.conv2d(x, w, padding=ks // 2).backward(dy)
    dw = k.conv_wgrad(to_nhwc(x), to_nhwc(dy), taps, Cin, Cout).reshape(Cout, Cin, ks, ks)
    ref = w.grad
    assert rel_err(dw, ref) < 2e-3, rel_err(dw, ref)
    torch.testing.assert_close(dw, ref, rtol=2e-2, atol=2e-3 * float(ref.abs().max()))


def test_conv_wgrad_padded_channels():
    """logical 3 -> 32 first layer: stored input has 16 channels, gradient is returned for the 3 real ones."""
    k = kern()
    N, H, W = 2, 32, 32
    x = bf16r(_rand(N, 3, H, W, seed=1))
    w = _rand(32, 3, 3, 3, seed=2).requires_grad_(True)
    dy = bf16r(_rand(N, 32, H, W, seed=3))
    F.conv2d(x, w, padding=1).backward(dy)
    xp = k.nchw_to_nhwc(x, 16)
    dw = k.conv_wgrad(xp, to_nhwc(dy), 9, 3, 32).reshape(32, 3, 3, 3)
    assert rel_err(dw, w.grad) < 2e-3


def test_pack_weight_layouts():
    """packed taps are dx-major (t = kw*3 + kh); the dgrad copy is tap-flipped and channel-transposed."""
    k = kern()
    w = _rand(5, 3, 3, 3, seed=1)
    wf, wd = k.pack_conv_weight(w)
    assert wf.shape == (9, 16, 16) and wd.shape == (9, 16, 16)
    ref_f = torch.zeros(9, 16, 16, device=DEV)
    ref_f[:, :5, :3] = bf16r(w).permute(3, 2, 0, 1).reshape(9, 5, 3)
    assert torch.equal(wf.float(), ref_f)
    ref_d = torch.zeros(9, 16, 16, device=DEV)
    ref_d[:, :3, :5] = ref_f[:, :5, :3].flip(0).permute(0, 2, 1)
    assert torch.equal(wd.float(), ref_d)


def test_conv_persistent_kernel_matches_generic_kernel_at_baseline_size():
    """BASELINE-size layer (batch 12, 128 -> 128 @ 128^2): the persistent 16x16-tile kernel and the generic 128-pixel
    kernel (uz_set_debug_flags(32)) must agree to bf16 rounding, and both match torch on a sub-batch."""
    from b200 import _lib
    k = kern()
    N, H, C = 12, 128, 128
    x = bf16r(_rand(N, C, H, H, seed=1))
    w = bf16r(_rand(C, C, 3, 3, seed=2, scale=(2.0 / (C * 9)) ** 0.5))
    wf, _ = k.pack_conv_weight(w, need_dgrad=False)
    xn = to_nhwc(x)
    y2, p2 = k.conv_fwd(xn, wf, stats=True)
    _lib.call('uz_set_debug_flags', 32)
    try:
        y1, p1 = k.conv_fwd(xn, wf, stats=True)
    finally:
        _lib.call('uz_set_debug_flags', 0)
    assert p2.shape == p1.shape == (1, 2, C)          # [2][Cout] accumulators
    d = (y1.float() - y2.float()).abs()
    assert float(d.max()) <= 2 ** -7 * float(y1.float().abs().max())
    torch.testing.assert_close(p1.sum(0), p2.sum(0), rtol=1e-3, atol=1e-1)
    ref = F.conv2d(x[:2], w, padding=1)
    _assert_bf16_close(to_nchw(y2[:2].contiguous()), ref, 'persistent conv')


@pytest.mark.parametrize('N,H,W,C', [(12, 16, 16, 64), (3, 32, 32, 192), (12, 2, 2, 192), (2, 64, 64, 32)])
def test_batchnorm_train_forward_backward(N, H, W, C):
    """conv-epilogue statistics -> uz_bn_finalize -> uz_affine_act, and the two-pass backward, against
    F.batch_norm(training=True, eps=1e-3, momentum=0.01) + ReLU autograd on the same stored pre-activations."""
    k = kern()
    x = bf16r(_rand(N, 16, H, W, seed=1))
    w = bf16r(_rand(C, 16, 3, 3, seed=2, scale=0.2))
    gamma = (1 + 0.1 * _rand(C, seed=3)).requires_grad_(True)
    beta = (0.1 * _rand(C, seed=4)).requires_grad_(True)
    rm, rv = 0.1 * _rand(C, seed=5), _rand(C, seed=6).abs() + 0.5
    rm_ref, rv_ref = rm.clone(), rv.clone()
    wf, _ = k.pack_conv_weight(w, need_dgrad=False)
    y, partial = k.conv_fwd(to_nhwc(x), wf, stats=True)
    scale, shift, mean, invstd = k.bn_finalize(partial, N * H * W, gamma.detach(), beta.detach(), rm, rv)
    a = k.affine_act(y, scale, shift, relu=True)
    yq = to_nchw(y).requires_grad_(True)
    ref = F.relu(F.batch_norm(yq, rm_ref, rv_ref, gamma, beta, True, 0.01, 1e-3))
    torch.testing.assert_close(mean, yq.detach().mean((0, 2, 3)), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rm, rm_ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rv, rv_ref, rtol=1e-4, atol=1e-6)
    _assert_bf16_close(to_nchw(a), ref.detach(), 'bn+relu')
    da = bf16r(_rand(N, C, H, W, seed=7))
    ref.backward(da)
    # fused single-launch forward (finalize + normalise + ReLU from the accumulators) must agree with the split path
    rm2, rv2 = 0.1 * _rand(C, seed=5), _rand(C, seed=6).abs() + 0.5
    a2, scale2, shift2, mean2, invstd2 = k.bn_apply_train(y, partial, N * H * W, gamma.detach(), beta.detach(), rm2, rv2)
    assert torch.equal(a2, a) or float((a2.float() - a.float()).abs().max()) <= 2 ** -7 * float(a.float().abs().max())
    torch.testing.assert_close(scale2, scale, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(mean2, mean, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(rm2, rm, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rv2, rv, rtol=1e-4, atol=1e-6)
    dy2, dgamma2, dbeta2 = k.bn_relu_bwd_train(to_nhwc(da), y, scale, shift, gamma.detach(), mean, invstd)
    assert rel_err(to_nchw(dy2), yq.grad) < 1e-2
    torch.testing.assert_close(dgamma2, gamma.grad, rtol=2e-3, atol=2e-3 * float(gamma.grad.abs().max()))
    torch.testing.assert_close(dbeta2, beta.grad, rtol=2e-3, atol=2e-3 * float(beta.grad.abs().max()))
    dy, dgamma, dbeta = k.bn_relu_bwd(to_nhwc(da), y, scale, shift, gamma.detach(), mean, invstd)
    # the ReLU mask is taken from fp32 a = y*scale+shift in both paths; differences are bf16 storage of dy only
    assert rel_err(to_nchw(dy), yq.grad) < 1e-2
    torch.testing.assert_close(dgamma, gamma.grad, rtol=2e-3, atol=2e-3 * float(gamma.grad.abs().max()))
    torch.testing.assert_close(dbeta, beta.grad, rtol=2e-3, atol=2e-3 * float(beta.grad.abs().max()))


def test_bn_eval_fold():
    k = kern()
    C = 64
    bias, gamma, beta = _rand(C, seed=1), 1 + 0.1 * _rand(C, seed=2), _rand(C, seed=3)
    rm, rv = _rand(C, seed=4), _rand(C, seed=5).abs() + 0.5
    scale, shift = k.bn_eval_fold(bias, gamma, beta, rm, rv)
    y = _rand(2, C, 4, 4, seed=6)
    ref = F.batch_norm(y + bias[None, :, None, None], rm, rv, gamma, beta, False, 0.01, 1e-3)
    got = y * scale[None, :, None, None] + shift[None, :, None, None]
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)


def test_avgpool_forward_backward():
    k = kern()
    x = bf16r(_rand(3, 32, 16, 24, seed=1)).requires_grad_(True)
    ref = F.avg_pool2d(x, 2, 2, 0, ceil_mode=True)
    got = k.avgpool2_fwd(to_nhwc(x))
    _assert_bf16_close(to_nchw(got), ref.detach(), 'pool')
    g = bf16r(_rand(*ref.shape, seed=2))
    ref.backward(g)
    dx = k.avgpool2_bwd(to_nhwc(g))
    _assert_bf16_close(to_nchw(dx), x.grad, 'pool bwd')


@pytest.mark.parametrize('align', [True, False])
@pytest.mark.parametrize('h,w', [(2, 2), (4, 4), (16, 16), (5, 7), (64, 64)])
def test_bilinear_x2_forward_backward(align, h, w):
    k = kern()
    x = bf16r(_rand(2, 16, h, w, seed=1)).requires_grad_(True)
    ref = F.interpolate(x, mode='bilinear', scale_factor=2, align_corners=align)
    got = k.upsample2x_fwd(to_nhwc(x), align_corners=align)
    _assert_bf16_close(to_nchw(got), ref.detach(), 'up2')
    g = bf16r(_rand(*ref.shape, seed=2))
    ref.backward(g)
    dx = k.upsample2x_bwd(to_nhwc(g), align_corners=align)
    _assert_bf16_close(to_nchw(dx), x.grad, 'up2 bwd')


def test_bilinear_known_answer():
    """2x2 ramp -> first row [0,1/3,2/3,1] (align_corners=True), [0,.25,.75,1] (False)."""
    k = kern()
    x = torch.zeros(1, 2, 2, 16, dtype=torch.bfloat16, device=DEV)
    x[0, :, :, 0] = torch.tensor([[0., 1.], [2., 3.]])
    a = k.upsample2x_fwd(x, align_corners=True)[0, 0, :, 0].float().cpu().numpy()
    b = k.upsample2x_fwd(x, align_corners=False)[0, 0, :, 0].float().cpu().numpy()
    np.testing.assert_allclose(a, [0, 1 / 3, 2 / 3, 1], atol=4e-3)
    np.testing.assert_allclose(b, [0, 0.25, 0.75, 1], atol=1e-6)


def test_input_pack_onehot_bit_exact():
    from oracle import phiseg_oracle as po
    k = kern()
    patch = _rand(3, 1, 16, 16, seed=1)
    mask = (torch.rand(3, 1, 16, 16, device=DEV) > 0.6).float()
    got = k.input_pack(patch, mask, nlabels=2, cp=16)
    ref = torch.cat([patch, po.onehot_minus_half(mask, 2)], 1)
    assert torch.equal(to_nchw(got)[:, :3], bf16r(ref))
    assert float(to_nchw(got)[:, 3:].abs().max()) == 0
    got2 = k.input_pack(patch, None, cp=16)
    assert torch.equal(to_nchw(got2)[:, :1], bf16r(patch)) and float(to_nchw(got2)[:, 1:].abs().max()) == 0


def test_layout_roundtrip_and_copy_channels():
    k = kern()
    x = _rand(2, 2, 8, 8, seed=1)
    nh = k.nchw_to_nhwc(x, 16)
    assert nh.shape == (2, 8, 8, 16)
    assert torch.equal(k.nhwc_to_nchw(nh, 2), bf16r(x))
    buf = torch.zeros(2, 8, 8, 48, dtype=torch.bfloat16, device=DEV)
    k.copy_channels(nh, buf[..., 16:32])
    k.copy_channels(nh, buf[..., 16:32], accumulate=True)
    assert torch.equal(buf[..., 16:32].float(), 2 * nh.float()) and float(buf[..., :16].float().abs().max()) == 0


@pytest.mark.parametrize('B,r,C', [(12, 32, 192), (12, 2, 192), (4, 16, 256)])
def test_head_forward_backward(B, r, C):
    k = kern()
    feat = bf16r(_rand(B, C, r, r, seed=1)).requires_grad_(True)
    wmu = (_rand(2, C, seed=2) * 0.05).requires_grad_(True)
    wsg = (_rand(2, C, seed=3) * 0.05).requires_grad_(True)
    bmu = _rand(2, seed=4).requires_grad_(True)
    bsg = _rand(2, seed=5).requires_grad_(True)
    eps = _rand(B, 2, r, r, seed=6)
    mu_r = F.conv2d(feat, wmu[:, :, None, None], bmu)
    sg_r = F.softplus(F.conv2d(feat, wsg[:, :, None, None], bsg))
    z_r = mu_r + sg_r * eps
    mu, sg, z = k.head_fwd(to_nhwc(feat.detach()), wmu.detach(), bmu.detach(), wsg.detach(), bsg.detach(), eps)
    for a, b in ((mu, mu_r), (sg, sg_r), (z, z_r)):
        torch.testing.assert_close(a, b.detach(), rtol=1e-4, atol=1e-4)
    gmu, gsg, gz = _rand(B, 2, r, r, seed=7), _rand(B, 2, r, r, seed=8), _rand(B, 2, r, r, seed=9)
    (mu_r * gmu + sg_r * gsg + z_r * gz).sum().backward()
    dfeat, dw, db = k.head_bwd(to_nhwc(feat.detach()), wmu.detach(), wsg.detach(), eps, sg, gmu, gsg, gz)
    assert rel_err(to_nchw(dfeat), feat.grad) < 5e-3
    torch.testing.assert_close(dw[:2], wmu.grad, rtol=1e-3, atol=1e-3 * float(wmu.grad.abs().max()))
    torch.testing.assert_close(dw[2:], wsg.grad, rtol=1e-3, atol=1e-3 * float(wsg.grad.abs().max()))
    torch.testing.assert_close(db[:2], bmu.grad, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(db[2:], bsg.grad, rtol=1e-3, atol=1e-3)


def test_kl_known_answer_and_gradients():
    """KL-1 of SURVEY.md Appendix B (sigma1*sigma0 quirk), then random tensors vs the oracle with autograd."""
    from oracle import phiseg_oracle as po
    k = kern()
    t = lambda v: torch.tensor(v, device=DEV).reshape(1, 1, 1, 2)
    out = k.kl_fwd(t([0.5, -1.0]), t([0.8, 1.5]), t([0.0, 0.25]), t([1.2, 0.7]), 1.0)
    assert float(out) == pytest.approx(1.1006804704666138, rel=1e-6)
    B, r = 12, 16
    mu0, mu1 = _rand(B, 2, r, r, seed=1).requires_grad_(True), _rand(B, 2, r, r, seed=2).requires_grad_(True)
    s0 = (_rand(B, 2, r, r, seed=3).abs() + 0.1).requires_grad_(True)
    s1 = (_rand(B, 2, r, r, seed=4).abs() + 0.1).requires_grad_(True)
    ref = 16 * po.kl_two_gauss(mu0, s0, mu1, s1)
    got = k.kl_fwd(mu0.detach(), s0.detach(), mu1.detach(), s1.detach(), 16.0)
    assert float(got) == pytest.approx(float(ref), rel=1e-5)
    (ref * 0.7).backward()
    up = torch.tensor([0.7], device=DEV)
    g = k.kl_bwd(mu0.detach(), s0.detach(), mu1.detach(), s1.detach(), 16.0, up)
    for a, b in zip(g, (mu0.grad, s0.grad, mu1.grad, s1.grad)):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5 * float(b.abs().max()))


@pytest.mark.parametrize('ncls,factor,C', [(2, 1, 128), (2, 16, 192), (3, 4, 192), (12, 1, 192), (16, 2, 256)])
def test_slayer_forward_backward(ncls, factor, C):
    k = kern()
    B, h = 3, 8
    feat = bf16r(_rand(B, C, h, h, seed=1)).requires_grad_(True)
    w = (_rand(ncls, C, seed=2) * 0.1).requires_grad_(True)
    b = _rand(ncls, seed=3).requires_grad_(True)
    ref = F.interpolate(F.conv2d(feat, w[:, :, None, None], b), size=[h * factor, h * factor], mode='nearest')
    got = k.slayer_fwd(to_nhwc(feat.detach()), w.detach(), b.detach(), factor)
    torch.testing.assert_close(got, ref.detach(), rtol=1e-4, atol=1e-4)
    g = _rand(*ref.shape, seed=4)
    ref.backward(g)
    dfeat, dw, db = k.slayer_bwd(g, to_nhwc(feat.detach()), w.detach(), factor)
    assert rel_err(to_nchw(dfeat), feat.grad) < 5e-3
    torch.testing.assert_close(dw, w.grad, rtol=1e-3, atol=1e-3 * float(w.grad.abs().max()))
    torch.testing.assert_close(db, b.grad, rtol=1e-3, atol=1e-3 * float(b.grad.abs().max()))


@pytest.mark.parametrize('ncls', [2, 3])
def test_residual_ce_and_accumulate(ncls):
    from oracle import phiseg_oracle as po
    k = kern()
    B, H = 4, 32
    s = [(_rand(B, ncls, H, H, seed=10 + l)).requires_grad_(True) for l in range(5)]
    target = torch.randint(0, ncls, (B, 1, H, H), device=DEV).float()
    acc, ref_levels = None, {}
    for lvl in reversed(range(5)):
        acc = s[lvl] if acc is None else acc + s[lvl]
        ref_levels[lvl] = po.multinoulli(acc, target)
    ce, grads = k.residual_ce([t.detach() for t in s], target)
    for lvl in range(5):
        assert float(ce[lvl]) == pytest.approx(float(ref_levels[lvl]), rel=1e-5)
    sum(ref_levels.values()).backward()
    for lvl in range(5):
        torch.testing.assert_close(grads[lvl], s[lvl].grad, rtol=1e-4, atol=1e-6)
    # accumulate_output: in place into the last entry, optional softmax (quirk Q2)
    lst = [t.detach().clone() for t in s]
    ref_sum = po.accumulate_output(lst, use_softmax=True)
    out = k.accumulate_output(lst, True, lst[-1])
    assert out.data_ptr() == lst[-1].data_ptr()
    torch.testing.assert_close(out, ref_sum, rtol=1e-5, atol=1e-6)


def test_ged_bit_exact_and_ncc(golden_dir):
    from oracle import metrics_oracle as mo
    k = kern()
    g = np.load(os.path.join(golden_dir, 'metrics.npz'))
    for tag, C in (('bin', 2), ('tri', 3)):
        samples = torch.from_numpy(g[tag + '_samples'].astype(np.int64)).to(DEV)
        gts = torch.from_numpy(g[tag + '_gts'].astype(np.float32)).to(DEV)
        out = k.ged(samples, gts, list(range(1, C))).cpu().numpy()
        assert out[0] == float(g[tag + '_ged'])          # bit exact vs the reference's python loops
        sy, ss, yy = mo.ged_pair_sums(g[tag + '_samples'], g[tag + '_gts'], C - 1, range(1, C))
        assert (out[1], out[2], out[3]) == (sy, ss, yy)
        logits = torch.from_numpy(g[tag + '_logits']).to(DEV)
        probs = torch.softmax(logits, 1)
        onehot = torch.from_numpy(g[tag + '_onehot'].astype(np.int64)).to(DEV)
        ncc = k.variance_ncc(probs, onehot).cpu().numpy()
        np.testing.assert_allclose(ncc, g[tag + '_ncc'], atol=1e-6)      # north-star: within 1e-4 absolute
        am = k.argmax_classes(logits)
        assert torch.equal(am.long(), logits.argmax(1))


def test_ged_known_answers_and_empty_sets():
    k = kern()
    s = torch.zeros(3, 2, 4, dtype=torch.int64, device=DEV)
    s[0, 0, :2] = 1
    s[1, 0, :1] = 1
    y = torch.zeros(2, 2, 4, device=DEV)
    y[0, 0, :3] = 1
    assert float(k.ged(s, y, [1])[0]) == pytest.approx(5 / 18, abs=1e-15)
    assert float(k.ged(s, s, [1])[0]) == 0.0
    z2 = torch.zeros(2, 2, 4, dtype=torch.int64, device=DEV)
    z3 = torch.zeros(3, 2, 4, device=DEV)
    assert float(k.ged(z2, z3, [1])[0]) == 0.0
    s4 = torch.tensor([[[1, 2, 2, 0]], [[1, 1, 0, 0]]], device=DEV)
    y4 = torch.tensor([[[1., 2., 0., 0.]]], device=DEV)
    assert float(k.ged(s4, y4, [1, 2])[0]) == pytest.approx(0.625, abs=1e-15)


def test_ncc_identity_reference_unit_test():
    """reference test/test_scores.py:31-50 (NCC(gt, gt) == 1): the input-independent form is N == M == 1."""
    k = kern()
    p = torch.softmax(_rand(1, 2, 64, 64, seed=1), 1)
    out = k.variance_ncc(p, p)
    assert float(out[0]) == pytest.approx(1.0, abs=1e-6)


def test_ged_full_size_properties():
    """BASELINE size (N=100 samples, M=4 annotators, 128x128): GED(s, s) == 0, symmetry in the pair sums, and
    equality with the oracle on a sub-sample."""
    from oracle import metrics_oracle as mo
    k = kern()
    g = torch.Generator(device='cpu').manual_seed(0)
    samples = (torch.rand(100, 128, 128, generator=g) < 0.1).long().to(DEV)
    samples[:7] = 0
    gts = (torch.rand(4, 128, 128, generator=g) < 0.1).float().to(DEV)
    gts[3] = 0
    out = k.ged(samples, gts, [1]).cpu().numpy()
    self_out = k.ged(samples, samples.float(), [1]).cpu().numpy()
    assert self_out[0] == 0.0
    sub = mo.generalised_energy_distance(samples[:9].cpu().numpy(), gts.cpu().numpy(), 1, range(1, 2))
    assert float(k.ged(samples[:9], gts, [1])[0]) == sub
    assert np.isfinite(out).all()


def test_fused_adam_matches_torch_adam():
    """b200.optim.FusedAdam (one launch for all parameters) against stock torch.optim.Adam, three steps, odd sizes."""
    from b200.optim import FusedAdam
    g = torch.Generator(device='cpu').manual_seed(5)
    shapes = [(192, 192, 3, 3), (7,), (33, 5, 3, 3), (4097,), (1,), (64, 16, 1, 1)]
    ref_p = [torch.nn.Parameter(torch.randn(*s, generator=g).to(DEV)) for s in shapes]
    my_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    ref = torch.optim.Adam(ref_p, lr=1e-3, weight_decay=1e-5)
    mine = FusedAdam(my_p, lr=1e-3, weight_decay=1e-5)
    for it in range(3):
        grads = [torch.randn(*s, generator=g).to(DEV) * (0.1 + it) for s in shapes]
        for p, q, gr in zip(ref_p, my_p, grads):
            p.grad = gr.clone()
            q.grad = gr.clone() if it != 1 or p.numel() != 7 else None      # a parameter without gradient is skipped
            if q.grad is None:
                p.grad = None
        ref.step()
        mine.step()
    for p, q in zip(ref_p, my_p):
        torch.testing.assert_close(q, p, rtol=2e-6, atol=1e-7)
    for p, q in zip(ref_p, my_p):
        torch.testing.assert_close(mine.state[q]['exp_avg_sq'], ref.state[p]['exp_avg_sq'], rtol=1e-5, atol=1e-12)


def test_kl_hierarchy_matches_per_level_kernels():
    """uz_kl_hierarchy_fwd / _bwd (all latent levels in one launch, models/phiseg.py:463-472) against the per-level
    kernels: level values bit-identical (same per-block fp64 partial sums), total = fp32 sum in the reference's order
    (level L-1 first), gradients bit-identical for an upstream of 1 and scaled for any other."""
    k = kern()
    B, L = 12, 5
    lw = [4.0 ** i for i in range(L)]
    levels = []
    for l in range(L):
        r = 4 << l
        mk = lambda s, pos: ((_rand(B, 2, r, r, seed=s).abs() + 0.1) if pos else _rand(B, 2, r, r, seed=s)).contiguous()
        levels.append((mk(10 * l, False), mk(10 * l + 1, True), mk(10 * l + 2, False), mk(10 * l + 3, True)))
    total, per = k.kl_hierarchy_fwd(levels, lw, 0.5)
    ref = [k.kl_fwd(*levels[l], lw[l]) for l in range(L)]
    for l in range(L):
        assert torch.equal(per[l:l + 1], ref[l]), l
    want = torch.zeros((), device=DEV)
    for l in reversed(range(L)):
        want = want + 0.5 * ref[l][0] if l != L - 1 else 0.5 * ref[l][0]
    torch.testing.assert_close(total[0], want, rtol=1e-6, atol=0)
    up = torch.full((1,), 0.75, device=DEV)
    grads = k.kl_hierarchy_bwd(levels, lw, 0.5, up)
    for l in range(L):
        g_ref = k.kl_bwd(*levels[l], lw[l], torch.full((1,), 0.75 * 0.5, device=DEV))
        for a, b in zip(grads[4 * l:4 * l + 4], g_ref):
            torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-30)


def test_fused_adam_with_weight_packing_matches_separate_passes():
    """uz_adam_pack_step: Adam on the conv weights with the bf16 forward / dgrad copies rewritten in the same pass, against
    the plain fused Adam followed by the packing kernel -- parameters, moments and packed copies bit for bit -- for 3x3,
    1x1 and 3x3x3 layers with odd channel counts, next to unpacked tensors in the same optimizer."""
    from b200.optim import FusedAdam
    k = kern()
    g = torch.Generator(device='cpu').manual_seed(9)
    shapes = [(192, 192, 3, 3), (32, 1, 3, 3), (48, 33, 3, 3), (64, 16, 1, 1), (32, 16, 3, 3, 3), (7,), (4097,)]
    a_p = [torch.nn.Parameter(torch.randn(*s, generator=g).to(DEV)) for s in shapes]
    b_p = [torch.nn.Parameter(p.detach().clone()) for p in a_p]
    pk_a = k.WeightPacker([p for p in a_p if p.dim() >= 4])
    pk_b = k.WeightPacker([p for p in b_p if p.dim() >= 4])
    opt_a = FusedAdam(a_p, lr=1e-3, weight_decay=1e-5)
    opt_a.attach_packer(pk_a)
    opt_b = FusedAdam(b_p, lr=1e-3, weight_decay=1e-5)
    assert pk_a.external and not pk_b.external
    for it in range(3):
        grads = [torch.randn(*s, generator=g).to(DEV) * (0.1 + it) for s in shapes]
        for p, q, gr in zip(a_p, b_p, grads):
            p.grad, q.grad = gr.clone(), gr.clone()
        opt_a.step()                     # update + re-pack
        pk_a.refresh()                   # no-op: the optimizer keeps the copies current
        opt_b.step()
        pk_b.refresh()                   # separate packing pass
        torch.cuda.synchronize()
        for p, q in zip(a_p, b_p):
            assert torch.equal(p, q)
            assert torch.equal(opt_a.state[p]['exp_avg'], opt_b.state[q]['exp_avg'])
            assert torch.equal(opt_a.state[p]['exp_avg_sq'], opt_b.state[q]['exp_avg_sq'])
            assert torch.equal(opt_a.state[p]['step'], opt_b.state[q]['step'])
            if p.dim() >= 4:
                (fa, da), (fb, db) = pk_a.lookup(p), pk_b.lookup(q)
                assert torch.equal(fa, fb) and torch.equal(da, db), tuple(p.shape)
