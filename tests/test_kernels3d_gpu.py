"""Per-kernel parity of the volume (NDHWC) kernels behind models/phiseg3D.py against plain torch fp32 ops on the same
bf16-rounded inputs, through the C ABI.  Tolerances as in tests/test_kernels_gpu.py: tensor-core convs
|err| <= 2^-7 |ref| + 1e-2 rms (bf16 output storage), fp32 reductions 1e-4 relative."""
import pytest
import torch
import torch.nn.functional as F

from tests.gpu_util import bf16r, kern, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def to_ndhwc(x):
    return x.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16)


def to_ncdhw(x):
    return x.float().permute(0, 4, 1, 2, 3).contiguous()


def _assert_bf16_close(got, ref, what=''):
    rms = float(ref.pow(2).mean().sqrt())
    err = (got - ref).abs()
    tol = ref.abs() * 2 ** -7 + 1e-2 * rms
    bad = int((err > tol).sum())
    assert bad == 0, '%s: %d / %d outside tolerance, max err %g (rms %g)' % (what, bad, err.numel(), float(err.max()), rms)


CONV3D_SHAPES = [
    # N, D, H, W, Cin, Cout
    (1, 16, 16, 16, 32, 32),
    (2, 8, 32, 32, 32, 64),       # KC = 32
    (1, 6, 16, 48, 64, 64),       # KC = 64, non-cubic
    (2, 8, 8, 8, 64, 96),         # partial tiles (8 < 16): TMA clips the stores, statistics mask the overhang
    (1, 5, 12, 20, 96, 32),       # odd sizes, 3 x 32-channel K blocks
    (1, 4, 16, 16, 128, 128),
    (2, 4, 4, 4, 32, 32),
    (1, 8, 16, 16, 16, 16),       # 16-channel halves of a reversible block: 16-column epilogue, 32-byte store boxes
    (1, 4, 16, 32, 64, 48),       # odd multiple of 16 output channels
    (1, 4, 16, 16, 16, 64),       # dgrad towards a padded 2-channel latent has Cout = 16 (here the forward direction)
]


@pytest.mark.parametrize('N,D,H,W,Cin,Cout', CONV3D_SHAPES)
def test_conv3d_fwd_dgrad_wgrad_match_torch(N, D, H, W, Cin, Cout):
    k = kern()
    x = bf16r(_rand(N, Cin, D, H, W, seed=1))
    w = bf16r(_rand(Cout, Cin, 3, 3, 3, seed=2, scale=(2.0 / (Cin * 27)) ** 0.5))
    wf, wd = k.pack_conv_weight(w)
    assert wf.shape == (27, Cout, Cin) and wd.shape == (27, Cin, Cout)
    y, sums = k.conv_fwd(to_ndhwc(x), wf, stats=True)
    ref = F.conv3d(x, w, padding=1)
    _assert_bf16_close(to_ncdhw(y), ref, 'conv3d fwd')
    yq = to_ncdhw(y)
    torch.testing.assert_close(sums[0, 0], yq.sum((0, 2, 3, 4)), rtol=1e-4, atol=2e-2)
    torch.testing.assert_close(sums[0, 1], yq.pow(2).sum((0, 2, 3, 4)), rtol=1e-4, atol=2e-2)
    # input gradient = the same kernel with the flipped / transposed weights
    dy = bf16r(_rand(N, Cout, D, H, W, seed=3))
    dx, _ = k.conv_fwd(to_ndhwc(dy), wd)
    ref_dx = torch.nn.grad.conv3d_input(x.shape, w, dy, padding=1)
    _assert_bf16_close(to_ncdhw(dx), ref_dx, 'conv3d dgrad')
    # weight gradient (fp32 out)
    dw = k.conv_wgrad(to_ndhwc(x), to_ndhwc(dy), 27, Cin, Cout).view(Cout, Cin, 3, 3, 3)
    ref_dw = torch.nn.grad.conv3d_weight(x, w.shape, dy, padding=1)
    assert rel_err(dw, ref_dw) < 2e-3
    torch.testing.assert_close(dw, ref_dw, rtol=2e-2, atol=2e-3 * float(ref_dw.abs().max()))


def test_conv3d_affine_relu_and_channel_slices():
    k = kern()
    N, D, H, W, Cin, Cout = 1, 8, 16, 16, 32, 64
    x = bf16r(_rand(N, Cin, D, H, W, seed=4))
    w = bf16r(_rand(Cout, Cin, 3, 3, 3, seed=5, scale=0.05))
    scale = _rand(Cout, seed=6).abs() + 0.5
    shift = _rand(Cout, seed=7)
    wf, _ = k.pack_conv_weight(w, need_dgrad=False)
    xbuf = torch.zeros((N, D, H, W, 96), dtype=torch.bfloat16, device=DEV)
    xbuf[..., 32:64] = to_ndhwc(x)
    ybuf = torch.full((N, D, H, W, 128), 7.0, dtype=torch.bfloat16, device=DEV)
    k.conv_fwd(xbuf[..., 32:64], wf, out=ybuf[..., 64:128], scale=scale, shift=shift, relu=True)
    ref = F.relu(F.conv3d(x, w, padding=1) * scale[None, :, None, None, None] + shift[None, :, None, None, None])
    _assert_bf16_close(to_ncdhw(ybuf[..., 64:128]), ref, 'conv3d affine')
    assert float((ybuf[..., :64].float() - 7.0).abs().max()) == 0.0      # neighbours untouched


def test_conv3d_pointwise():
    k = kern()
    N, D, H, W, Cin, Cout = 2, 4, 8, 8, 64, 32
    x = bf16r(_rand(N, Cin, D, H, W, seed=8))
    w = bf16r(_rand(Cout, Cin, 1, 1, 1, seed=9, scale=0.1))
    wf, wd = k.pack_conv_weight(w)
    y, _ = k.conv_fwd(to_ndhwc(x), wf)
    _assert_bf16_close(to_ncdhw(y), F.conv3d(x, w), 'conv3d 1x1x1')
    dy = bf16r(_rand(N, Cout, D, H, W, seed=10))
    dw = k.conv_wgrad(to_ndhwc(x), to_ndhwc(dy), 1, Cin, Cout).view(Cout, Cin, 1, 1, 1)
    assert rel_err(dw, torch.nn.grad.conv3d_weight(x, w.shape, dy)) < 2e-3


@pytest.mark.parametrize('N,D,H,W,C', [(2, 8, 8, 8, 32), (1, 4, 12, 20, 64)])
def test_avgpool3_and_trilinear_match_torch(N, D, H, W, C):
    k = kern()
    x = bf16r(_rand(N, C, D, H, W, seed=11))
    p = k.avgpool2_fwd(to_ndhwc(x))
    ref_p = F.avg_pool3d(x, 2, 2, 0, ceil_mode=True)
    torch.testing.assert_close(to_ncdhw(p), ref_p, rtol=2 ** -7, atol=1e-3)
    g = bf16r(_rand(*ref_p.shape, seed=12))
    dx = k.avgpool2_bwd(to_ndhwc(g))
    xr = x.clone().requires_grad_(True)
    F.avg_pool3d(xr, 2, 2, 0, ceil_mode=True).backward(g)
    torch.testing.assert_close(to_ncdhw(dx), xr.grad, rtol=2 ** -7, atol=1e-3)
    # trilinear x2 align_corners=True, forward into a channel slice of a wider buffer, and its transpose
    out = torch.zeros((N, 2 * D, 2 * H, 2 * W, C + 32), dtype=torch.bfloat16, device=DEV)
    k.upsample2x_fwd(to_ndhwc(x), True, out=out[..., 32:])
    ref_u = F.interpolate(x, mode='trilinear', scale_factor=2, align_corners=True)
    torch.testing.assert_close(to_ncdhw(out[..., 32:]), ref_u, rtol=2 ** -7, atol=4e-3)
    assert float(out[..., :32].float().abs().max()) == 0.0
    gu = bf16r(_rand(*ref_u.shape, seed=13))
    du = k.upsample2x_bwd(to_ndhwc(gu), True)
    xr = x.clone().requires_grad_(True)
    F.interpolate(xr, mode='trilinear', scale_factor=2, align_corners=True).backward(gu)
    torch.testing.assert_close(to_ncdhw(du), xr.grad, rtol=2 ** -6, atol=2e-2)


@pytest.mark.parametrize('factor', [1, 2, 4])
def test_slayer3d_matches_torch(factor):
    k = kern()
    N, d, C, ncls = 2, 4, 64, 3
    feat = bf16r(_rand(N, C, d, d, d, seed=14))
    w = _rand(ncls, C, seed=15, scale=0.2)
    b = _rand(ncls, seed=16)
    out = k.slayer_fwd(to_ndhwc(feat), w, b, factor)
    s_in = F.conv3d(feat, w.view(ncls, C, 1, 1, 1), b)
    ref = F.interpolate(s_in, size=[d * factor] * 3, mode='nearest')
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)
    g = _rand(*ref.shape, seed=17)
    dfeat, dw, db = k.slayer_bwd(g, to_ndhwc(feat), w, factor)
    fr = feat.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    F.interpolate(F.conv3d(fr, wr.view(ncls, C, 1, 1, 1), br), size=[d * factor] * 3, mode='nearest').backward(g)
    torch.testing.assert_close(dw, wr.grad, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(db, br.grad, rtol=1e-4, atol=1e-3)
    _assert_bf16_close(to_ncdhw(dfeat), fr.grad, 'slayer3d dfeat')


def test_flat_kernels_on_volumes():
    """layout conversion, input packing, BatchNorm apply on 5-D activations"""
    k = kern()
    x = _rand(2, 5, 4, 6, 8, seed=18)
    a = k.nchw_to_nhwc(x, 32)
    assert a.shape == (2, 4, 6, 8, 32)
    torch.testing.assert_close(to_ncdhw(a)[:, :5], bf16r(x), rtol=0, atol=0)
    assert float(a[..., 5:].float().abs().max()) == 0.0
    torch.testing.assert_close(k.nhwc_to_nchw(a, 5), bf16r(x), rtol=0, atol=0)
    vol = _rand(2, 4, 4, 6, 8, seed=19)
    lab = torch.randint(0, 3, (2, 1, 4, 6, 8), device=DEV).float()
    p = k.input_pack(vol, lab, nlabels=3, cp=32)
    ref = torch.cat([vol] + [(lab == c).float() - 0.5 for c in range(3)], dim=1)
    torch.testing.assert_close(to_ncdhw(p)[:, :7], bf16r(ref), rtol=0, atol=0)
    assert float(p[..., 7:].float().abs().max()) == 0.0
