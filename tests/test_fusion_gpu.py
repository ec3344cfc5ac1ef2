"""Round-2 kernels through the C ABI: the extended conv epilogue (uz_conv_fwd_ex: fused BatchNorm/ReLU-backward
reduction, residual add, per-CTA statistics rows), deterministic training mode, and the model-level equivalence of the
fused and unfused backward paths.  Tolerances at each assert."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.gpu_util import bf16r, kern, rel_err, to_nchw, to_nhwc

pytestmark = pytest.mark.gpu
DEV = 'cuda'
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


# (N, H, W, Cin, Cout): persistent kernel (H, W multiples of 16), generic kernel (small maps), 16-wide epilogue
EX_SHAPES = [(3, 32, 32, 64, 128), (12, 16, 16, 192, 192), (12, 8, 8, 192, 192), (12, 2, 2, 64, 192), (2, 32, 32, 64, 48),
             (2, 128, 128, 32, 32)]


@pytest.mark.parametrize('N,H,W,Cin,Cout', EX_SHAPES)
def test_conv_epilogue_fused_bn_backward_sums(N, H, W, Cin, Cout):
    """dgrad with the producer layer's (y, scale, shift): the stored gradient must equal the plain result times the ReLU
    mask BIT FOR BIT, and the accumulated (sum g, sum g*y) must match an fp32 reduction of those stored values to 1e-3
    (fp32 atomics across CTAs: order-dependent rounding only)."""
    k = kern()
    dy = to_nhwc(_rand(N, Cin, H, W, seed=1))
    w = bf16r(_rand(Cin, Cout, 3, 3, seed=2, scale=(2.0 / (Cin * 9)) ** 0.5))      # OIHW of the forward layer Cout -> Cin
    _, wd = k.pack_conv_weight(w, need_dgrad=True)                                 # dgrad packing [taps][Cout][Cin]
    y_prev = to_nhwc(_rand(N, Cout, H, W, seed=3))
    scale, shift = 1 + 0.2 * _rand(Cout, seed=4), 0.3 * _rand(Cout, seed=5)
    plain, _ = k.conv_fwd(dy, wd)
    k.zero_arena.reset(dy.device)
    fused, sums = k.conv_fwd(dy, wd, bn_prev=(y_prev, scale, shift, True))
    mask = (y_prev.float() * scale + shift) > 0
    want = torch.where(mask, plain.float(), torch.zeros_like(plain.float()))
    assert torch.equal(fused.float(), want)
    s0 = want.sum((0, 1, 2))
    s1 = (want * y_prev.float()).sum((0, 1, 2))
    ref_scale = float(want.abs().sum((0, 1, 2)).max())
    torch.testing.assert_close(sums[0], s0, rtol=1e-3, atol=1e-5 * ref_scale)
    torch.testing.assert_close(sums[1], s1, rtol=1e-3, atol=1e-5 * ref_scale * float(y_prev.float().abs().max()))
    # and those sums drive the apply pass to the same dy as the two-launch path
    gamma = 1 + 0.1 * _rand(Cout, seed=6)
    mean, invstd = 0.1 * _rand(Cout, seed=7), 1 + 0.1 * _rand(Cout, seed=8).abs()
    a, da_, db_ = k.bn_relu_bwd_train(fused, y_prev, scale, shift, gamma, mean, invstd, relu=True, sums=sums)
    b, dg, dbt = k.bn_relu_bwd_train(plain, y_prev, scale, shift, gamma, mean, invstd, relu=True)
    assert rel_err(a.float(), b.float()) < 2e-3
    torch.testing.assert_close(da_, dg, rtol=2e-3, atol=2e-3 * float(dg.abs().max()))
    torch.testing.assert_close(db_, dbt, rtol=2e-3, atol=2e-3 * float(dbt.abs().max()))


@pytest.mark.parametrize('N,H,W,Cin,Cout', [(3, 32, 32, 64, 64), (12, 8, 8, 96, 96)])
@pytest.mark.parametrize('sign', [1, -1])
def test_conv_epilogue_residual(N, H, W, Cin, Cout, sign):
    """out = residual + sign * conv(x): one rounding of the fp32 sum (the unfused path rounds the conv first)."""
    k = kern()
    x = bf16r(_rand(N, Cin, H, W, seed=1))
    w = bf16r(_rand(Cout, Cin, 3, 3, seed=2, scale=(2.0 / (Cin * 9)) ** 0.5))
    res = bf16r(_rand(N, Cout, H, W, seed=3))
    wf, _ = k.pack_conv_weight(w, need_dgrad=False)
    out, _ = k.conv_fwd(to_nhwc(x), wf, residual=to_nhwc(res), res_sign=sign)
    ref = res + sign * F.conv2d(x, w, padding=1)
    err = (to_nchw(out) - ref).abs()
    assert float((err - (ref.abs() * 2 ** -8 + 2e-3 * float(ref.abs().mean()))).max()) <= 0


@pytest.mark.parametrize('N,H,W,Cin,Cout', [(12, 16, 16, 64, 64), (12, 8, 8, 192, 192), (2, 64, 64, 32, 64)])
def test_deterministic_statistics_rows(N, H, W, Cin, Cout):
    """per-CTA statistics rows + fixed-order finalize: bit-identical across repeated launches, equal to the atomic
    accumulators within fp32 rounding."""
    k = kern()
    x = to_nhwc(_rand(N, Cin, H, W, seed=1))
    w = bf16r(_rand(Cout, Cin, 3, 3, seed=2, scale=0.1))
    wf, _ = k.pack_conv_weight(w, need_dgrad=False)
    k.zero_arena.reset(x.device)
    _, acc = k.conv_fwd(x, wf, stats=True)
    prev = k.set_deterministic(True)
    try:
        outs = []
        for _ in range(3):
            y, rows = k.conv_fwd(x, wf, stats=True)
            g, b = torch.ones(Cout, device=DEV), torch.zeros(Cout, device=DEV)
            outs.append((rows.clone(),) + tuple(t.clone() for t in k.bn_finalize(rows, N * H * W, g, b)))
    finally:
        k.set_deterministic(prev)
    assert outs[0][0].dim() == 3 and outs[0][0].shape[1:] == (2, Cout)
    for o in outs[1:]:
        for a, b in zip(o, outs[0]):
            assert torch.equal(a, b)
    torch.testing.assert_close(outs[0][0].sum(0), acc[0], rtol=1e-4, atol=1e-2)


def _phiseg_step(det, fuse, seed=0, filters=(16, 32, 32, 32, 32, 32, 32), B=4):
    from b200 import ops, train
    from oracle import synth
    from oracle.ref_run import injected_noise
    from tests.keygrammar import dropin_phiseg
    k = kern()
    net = dropin_phiseg(list(filters))
    net.load_state_dict(synth.synth_state_dict(net.state_dict(), seed=1))
    net = net.cuda().train()
    patch, labels, mask = synth.lidc_like_batch(B, seed=3)
    eps = synth.noise_list(synth.phiseg_noise_shapes(B), seed=5)
    pd, pf = k.set_deterministic(det), ops.set_fuse_bn_backward(fuse)
    try:
        with injected_noise(eps):
            net.forward(patch.cuda(), mask.cuda(), training=True)
            loss = net.loss(mask.cuda())
        loss.backward()
        torch.cuda.synchronize()
    finally:
        k.set_deterministic(pd)
        ops.set_fuse_bn_backward(pf)
    grads = {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}
    stats = {n: b.detach().clone() for n, b in net.named_buffers()}
    return float(loss), grads, stats


def test_deterministic_training_step_is_bit_reproducible():
    """UNETZOO_DETERMINISTIC: two identical steps give the same loss, gradients and BatchNorm running statistics bit
    for bit (the atomic path differed by 1.6e-3 run to run)."""
    l1, g1, s1 = _phiseg_step(True, False)
    l2, g2, s2 = _phiseg_step(True, False)
    assert l1 == l2
    for n in g1:
        assert torch.equal(g1[n], g2[n]), n
    for n in s1:
        assert torch.equal(s1[n], s2[n]), n


def test_fused_bn_backward_matches_unfused_step():
    """model level: the step with the BatchNorm-backward reductions fused into the dgrad epilogues (opt-in) vs the default
    step.  The random-weight fixture is chaotic (two runs of the SAME code end up 17 % apart in their gradients because
    the fp32 atomics of the forward statistics land in a different order), so the yardstick is that noise floor: the fused
    path must not be further from an unfused run than two unfused runs are from each other (x1.5).  The kernel itself is
    checked bit-exactly in test_conv_epilogue_fused_bn_backward_sums."""
    from b200 import _lib

    def med(a, b):
        v = [float((a[n] - b[n]).norm()) / float(b[n].norm()) for n in b if float(b[n].norm()) > 1e-6]
        return float(np.median(v))

    k = kern()
    prev = k._BN_BWD_FUSED          # the launch-count comparison is between the dgrad-epilogue sums and the reduce + apply pair
    k._BN_BWD_FUSED = False
    try:
        n0 = _lib.raw('uz_launch_count')()
        l_f, g_f, _ = _phiseg_step(False, True)
        n1 = _lib.raw('uz_launch_count')()
        l_u, g_u, _ = _phiseg_step(False, False)
        n2 = _lib.raw('uz_launch_count')()
        l_v, g_v, _ = _phiseg_step(False, False)
    finally:
        k._BN_BWD_FUSED = prev
    assert (n1 - n0) < (n2 - n1) - 20, 'fusion must remove reduction launches (%d vs %d)' % (n1 - n0, n2 - n1)
    floor_loss = abs(l_u - l_v) / abs(l_v)
    assert abs(l_f - l_u) / abs(l_u) <= max(3 * floor_loss, 2e-3)
    floor, fused = med(g_u, g_v), med(g_f, g_u)
    print('\nmedian gradient rel-L2: unfused vs unfused %.3e, fused vs unfused %.3e' % (floor, fused))
    assert fused <= 1.5 * floor + 1e-3


def _torch_reversible_reference(seq, x0, g0):
    """fp32 autograd of the same ReversibleSequence with stock torch ops (plain forward: gradients of the additive
    coupling equal those of the inverse-recompute backward)"""
    import copy
    x = x0.clone().requires_grad_(True)
    params = {}
    cur = x
    for bi, block in enumerate(seq.sequence.reversible_blocks):
        x1, x2 = torch.chunk(cur, 2, dim=1)

        def unit(m, inp, tag):
            conv, bn = m.convolution[0], m.convolution[1]
            w = conv.weight.detach().to(torch.bfloat16).float().requires_grad_(True)     # kernels round the weights to bf16
            g, b = bn.weight.detach().clone().requires_grad_(True), bn.bias.detach().clone().requires_grad_(True)
            params['%d.%s.w' % (bi, tag)], params['%d.%s.g' % (bi, tag)], params['%d.%s.b' % (bi, tag)] = w, g, b
            y = F.conv2d(inp, w, conv.bias.detach(), padding=1)
            return F.relu(F.batch_norm(y, None, None, g, b, True, 0.01, 1e-3))

        y1 = x1 + unit(block.f_block[0], x2, 'f')
        y2 = x2 + unit(block.g_block[0], y1, 'g')
        cur = torch.cat([y1, y2], dim=1)
    cur.backward(g0)
    return cur.detach(), x.grad.detach(), params


@pytest.mark.parametrize('shape', [(3, 32, 32, 64), (12, 8, 8, 128)])
def test_fused_reversible_sequence_matches_nested_autograd_path(shape):
    """ReversibleSequence forward + inverse-recompute backward: the fused kernels (coupling add inside the BatchNorm pass,
    coupling inverse inside the BatchNorm-backward reduction, gradient add in the dgrad epilogue) and the autograd-nested
    path (separate add launches) against fp32 torch autograd of the same stack.  The fused path rounds once less per
    coupling, so it must be at least as close to fp32 as the nested one (x1.5 + bf16 floor); running statistics (two
    momentum updates, quirk Q7) equal to fp32 rounding between the two paths."""
    import torchlayers
    n, h, w, c = shape
    torch.manual_seed(0)
    seq_a = torchlayers.ReversibleSequence(c, c, reversible_depth=2).cuda().train()
    seq_b = torchlayers.ReversibleSequence(c, c, reversible_depth=2).cuda().train()
    seq_b.load_state_dict(seq_a.state_dict())
    x0 = bf16r(_rand(n, c, h, w, seed=3))
    g0 = bf16r(_rand(n, c, h, w, seed=4))
    y_ref, dx_ref, p_ref = _torch_reversible_reference(seq_a, x0, g0)
    outs = []
    for seq, fused in ((seq_a, True), (seq_b, False)):
        prev = torchlayers.set_fused_reversible(fused)
        try:
            x = x0.clone().requires_grad_(True)
            y = seq(x)
            y.backward(g0)
            torch.cuda.synchronize()
        finally:
            torchlayers.set_fused_reversible(prev)
        outs.append((y.detach(), x.grad.detach(), {k: p.grad.detach() for k, p in seq.named_parameters()},
                     {k: b.detach().clone() for k, b in seq.named_buffers()}))
    (ya, dxa, ga, ba), (yb, dxb, gb, bb) = outs
    ey = (rel_err(ya, y_ref), rel_err(yb, y_ref))
    ed = (rel_err(dxa, dx_ref), rel_err(dxb, dx_ref))
    print('\nreversible stack %s: y rel-L2 vs fp32 fused %.3e nested %.3e; dx fused %.3e nested %.3e' % (shape, ey[0], ey[1], ed[0], ed[1]))
    assert ey[0] < 1.5 * ey[1] + 2e-3
    assert ed[0] < 1.5 * ed[1] + 1e-2
    wf = ga['sequence.reversible_blocks.0.f_block.0.convolution.0.weight']
    wn = gb['sequence.reversible_blocks.0.f_block.0.convolution.0.weight']
    assert rel_err(wf, p_ref['0.f.w'].grad) < 1.5 * rel_err(wn, p_ref['0.f.w'].grad) + 1e-2
    gf = ga['sequence.reversible_blocks.1.g_block.0.convolution.1.weight']
    gn = gb['sequence.reversible_blocks.1.g_block.0.convolution.1.weight']
    assert rel_err(gf, p_ref['1.g.g'].grad) < 1.5 * rel_err(gn, p_ref['1.g.g'].grad) + 1e-2
    for k in bb:
        if k.endswith('num_batches_tracked'):
            assert int(ba[k]) == int(bb[k]) == 2
        else:
            torch.testing.assert_close(ba[k], bb[k], rtol=2e-3, atol=1e-5)


# (N, H, W, C, channel-slice stride): K = 1 / 2 / 8 pixel slices per cluster, staged and re-read plans, odd pixel counts
BN_CLUSTER_SHAPES = [(12, 2, 2, 192, 192), (12, 4, 4, 192, 256), (5, 4, 4, 48, 48), (12, 8, 8, 256, 256),
                     (12, 16, 16, 64, 64), (12, 32, 32, 192, 192), (12, 64, 64, 64, 96), (3, 20, 12, 32, 32)]


@pytest.mark.parametrize('N,H,W,C,ld', BN_CLUSTER_SHAPES)
def test_bn_backward_cluster_kernel(N, H, W, C, ld):
    """uz_bn_bwd_fused (one launch on thread-block clusters, partial sums through distributed shared memory) against
    autograd through F.batch_norm(training=True) + ReLU in fp32 on the same stored values, against the two-launch path,
    and bit-reproducible run to run (fixed summation order)."""
    k = kern()
    assert k._lib.raw('uz_bn_bwd_fused_supported')(N * H * W, C) == 1
    ybuf = to_nhwc(bf16r(_rand(N, ld, H, W, seed=11)))               # y is a channel slice of a wider buffer
    y = ybuf[..., :C]
    gamma = (1 + 0.1 * _rand(C, seed=3)).requires_grad_(True)
    beta = (0.1 * _rand(C, seed=4)).requires_grad_(True)
    yq = to_nchw(y.contiguous()).requires_grad_(True)
    ref = F.relu(F.batch_norm(yq, None, None, gamma, beta, True, 0.01, 1e-3))
    da = bf16r(_rand(N, C, H, W, seed=7))
    ref.backward(da)
    mean = yq.detach().mean((0, 2, 3))
    var = yq.detach().var((0, 2, 3), unbiased=False)
    invstd = (var + 1e-3).rsqrt()
    scale = gamma.detach() * invstd
    shift = beta.detach() - mean * scale
    dout = to_nhwc(da)
    prev, prev_max = k._BN_BWD_FUSED, k._BN_BWD_FUSED_MAX_PIX
    try:
        k._BN_BWD_FUSED, k._BN_BWD_FUSED_MAX_PIX = True, 1 << 30      # the kernel's whole range, not only the routed sizes
        dy, dgamma, dbeta = k.bn_relu_bwd_train(dout, y, scale, shift, gamma.detach(), mean, invstd)
        dy_b, dgamma_b, dbeta_b = k.bn_relu_bwd_train(dout, y, scale, shift, gamma.detach(), mean, invstd)
        k._BN_BWD_FUSED = False
        dy_u, dgamma_u, dbeta_u = k.bn_relu_bwd_train(dout, y, scale, shift, gamma.detach(), mean, invstd)
    finally:
        k._BN_BWD_FUSED, k._BN_BWD_FUSED_MAX_PIX = prev, prev_max
    assert torch.equal(dy, dy_b) and torch.equal(dgamma, dgamma_b) and torch.equal(dbeta, dbeta_b)   # deterministic
    assert rel_err(to_nchw(dy), yq.grad) < 6e-3                       # bf16 storage of dy: 2^-9 rms
    torch.testing.assert_close(dgamma, gamma.grad, rtol=1e-4, atol=1e-4 * float(gamma.grad.abs().max()))
    torch.testing.assert_close(dbeta, beta.grad, rtol=1e-4, atol=1e-4 * float(beta.grad.abs().max()))
    # the two-launch path sums with atomics in arrival order: equal up to fp32 summation order
    torch.testing.assert_close(dgamma, dgamma_u, rtol=1e-4, atol=1e-4 * float(dgamma_u.abs().max()))
    assert float((dy.float() - dy_u.float()).abs().max()) <= 2 ** -7 * float(dy_u.float().abs().max())


@pytest.mark.parametrize('accumulate', [True, False])
def test_deferred_batched_wgrad_reduction(accumulate):
    """uz_conv_wgrad_partial + ONE uz_wgrad_reduce_batched launch for a mix of layers (persistent and generic plans, 1x1,
    padded logical channels, a 3x3x3 volume layer) against the per-layer uz_conv_wgrad.  accumulate: the split-K CTAs add
    into one slab through L2 (bulk reduce-add stores) -- equal up to fp32 summation order; slab mode (what the
    deterministic mode uses): splits summed strictly in order, bit-reproducible run to run."""
    k = kern()
    cases = [  # N, H, W, Cin, Cout, taps, Cin_logical, Cout_logical
        (12, 16, 16, 192, 192, 9, 192, 192), (12, 2, 2, 192, 192, 9, 192, 192), (3, 32, 32, 64, 128, 9, 64, 128),
        (2, 64, 64, 32, 32, 9, 32, 32), (2, 32, 32, 224, 128, 1, 224, 128), (2, 32, 32, 16, 32, 9, 3, 32),
        (12, 8, 8, 80, 64, 9, 66, 64), (4, 128, 128, 32, 32, 9, 32, 32)]
    prev_acc, prev_det = k._WGRAD_ACCUMULATE, k.set_deterministic(not accumulate)
    k._WGRAD_ACCUMULATE = accumulate
    k.wgrad_reducer.flush()
    try:
        ref, got, keep = [], [], []
        for i, (N, H, W, Cin, Cout, taps, cil, col) in enumerate(cases):
            x = to_nhwc(bf16r(_rand(N, Cin, H, W, seed=20 + i)))
            dy = to_nhwc(bf16r(_rand(N, Cout, H, W, seed=40 + i)))
            ref.append(k.conv_wgrad(x, dy, taps, cil, col))
            got.append(k.conv_wgrad(x, dy, taps, cil, col, defer=True))
            keep.append((x, dy))
        xv = bf16r(_rand(1, 32, 8, 16, 16, seed=3)).permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16)
        dv = bf16r(_rand(1, 48, 8, 16, 16, seed=4)).permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16)
        ref.append(k.conv_wgrad(xv, dv, 27, 32, 48))
        got.append(k.conv_wgrad(xv, dv, 27, 32, 48, defer=True))
        assert len(k.wgrad_reducer.items) == len(cases) + 1
        if accumulate:       # 2-D layers report one (shared) slab; volumes keep one slab per split
            assert all(it[1] == 1 for it in k.wgrad_reducer.items[:-1])
        k.wgrad_reducer.flush()
        assert not k.wgrad_reducer.items
        torch.cuda.synchronize()
        for i, (a, b) in enumerate(zip(got, ref)):
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5 * float(b.abs().max()), msg='layer %d' % i)
        if not accumulate:
            again = [k.conv_wgrad(x, dy, c[5], c[6], c[7], defer=True) for (x, dy), c in zip(keep, cases)]
            k.wgrad_reducer.flush()
            torch.cuda.synchronize()
            for i, (a, b) in enumerate(zip(again, got)):
                assert torch.equal(a, b), 'layer %d not reproducible' % i
    finally:
        k.wgrad_reducer.flush()
        k._WGRAD_ACCUMULATE = prev_acc
        k.set_deterministic(prev_det)


# N, H, W, Cin, Cout, taps: 1 / 2 / 6 pixel tiles per cluster, Cout chunks of 48 / 64 / 32, a 1x1 layer, padded inputs
CONV_BN_FUSED_SHAPES = [(12, 2, 2, 192, 192, 9), (12, 4, 4, 192, 192, 9), (12, 8, 8, 192, 192, 9), (12, 8, 8, 256, 256, 9),
                        (5, 4, 4, 256, 256, 9), (12, 8, 8, 64, 64, 9), (12, 2, 2, 16, 64, 9), (3, 8, 8, 64, 48, 9),
                        (12, 8, 8, 192, 32, 1), (16, 8, 8, 32, 32, 9)]


@pytest.mark.parametrize('N,H,W,Cin,Cout,taps', CONV_BN_FUSED_SHAPES)
def test_conv_bn_relu_cluster_kernel(N, H, W, Cin, Cout, taps):
    """uz_conv_bn_act_fused (conv + batch statistics through the cluster's distributed shared memory + normalise + ReLU in
    one launch) against the conv -> uz_bn_finalize -> uz_affine_act path and against torch fp32 F.batch_norm on the
    stored conv output; bit-reproducible run to run."""
    k = kern()
    ks = 3 if taps == 9 else 1
    x = to_nhwc(bf16r(_rand(N, Cin, H, W, seed=1)))
    w = bf16r(_rand(Cout, Cin, ks, ks, seed=2, scale=0.05))
    bias = 0.1 * _rand(Cout, seed=8)
    gamma, beta = 1 + 0.1 * _rand(Cout, seed=3), 0.1 * _rand(Cout, seed=4)
    rm0, rv0 = 0.1 * _rand(Cout, seed=5), _rand(Cout, seed=6).abs() + 0.5
    wf, _ = k.pack_conv_weight(w, need_dgrad=False)
    assert k.conv_bn_fused_supported(x, wf)
    rm, rv = rm0.clone(), rv0.clone()
    a, y, scale, shift, mean, invstd = k.conv_bn_act_fused(x, wf, bias, gamma, beta, rm, rv, relu=True)
    rm_b, rv_b = rm0.clone(), rv0.clone()
    a_b, y_b, scale_b, _, _, _ = k.conv_bn_act_fused(x, wf, bias, gamma, beta, rm_b, rv_b, relu=True)
    assert torch.equal(a, a_b) and torch.equal(y, y_b) and torch.equal(scale, scale_b) and torch.equal(rm, rm_b)
    # split path: same conv kernel (statistics rows reduced in fixed order), then finalize + affine
    prev = k.set_deterministic(True)
    try:
        y2, partial = k.conv_fwd(x, wf, shift=bias, stats=True)
    finally:
        k.set_deterministic(prev)
    rm2, rv2 = rm0.clone(), rv0.clone()
    scale2, shift2, mean2, invstd2 = k.bn_finalize(partial, N * H * W, gamma, beta, rm2, rv2)
    a2 = k.affine_act(y2, scale2, shift2, relu=True)
    import os
    small = os.environ.get('UZ_CONV_SMALL') == '1' and H <= 4 and W <= 4 and N * H * W <= 64 and Cin % 64 == 0 and \
        taps == 9                     # then served by the opt-in CUDA-core small-map kernel (conv_small.cu)
    if small:
        # different fp32 summation order than the tensor-core kernel: a few stored values differ by one bf16 ulp
        d = (y.float() - y2.float()).abs()
        assert float(d.max()) <= 2 ** -7 * float(y2.float().abs().max())
        assert float((d > 0).float().mean()) < 0.05
        tol = 2e-3
    else:
        assert torch.equal(y, y2)                                          # the stored conv output is the same kernel math
        tol = 1e-5
    torch.testing.assert_close(mean, mean2, rtol=tol, atol=tol * float(mean2.abs().max()))
    torch.testing.assert_close(invstd, invstd2, rtol=tol, atol=1e-6)
    torch.testing.assert_close(scale, scale2, rtol=tol, atol=1e-6)
    torch.testing.assert_close(shift, shift2, rtol=10 * tol, atol=tol * float(shift2.abs().max()))
    torch.testing.assert_close(rm, rm2, rtol=tol, atol=tol * float(rm2.abs().max()))
    torch.testing.assert_close(rv, rv2, rtol=tol, atol=1e-6)
    assert float((a.float() - a2.float()).abs().max()) <= 2 ** -7 * float(a2.float().abs().max())     # one bf16 ulp
    # torch fp32 on the stored y
    yq = to_nchw(y)
    ref = F.relu(F.batch_norm(yq, rm0.clone(), rv0.clone(), gamma, beta, True, 0.01, 1e-3))
    err = (to_nchw(a) - ref).abs()
    assert float(err.max()) <= 2 ** -7 * float(ref.abs().max()) + 1e-3
    conv_ref = F.conv2d(to_nchw(x), w, bias, padding=ks // 2)
    assert rel_err(yq, conv_ref) < 5e-3


@pytest.mark.parametrize('N,H,W,C,ld', [(12, 64, 64, 64, 96), (12, 128, 128, 32, 32), (2, 128, 128, 128, 128), (3, 70, 66, 48, 48)])
def test_bn_backward_cooperative_kernel(N, H, W, C, ld):
    """uz_bn_bwd_coop (sums + grid-wide barrier + apply in ONE cooperative launch, large maps) against autograd through
    F.batch_norm(training=True) + ReLU in fp32 and against the two-launch path."""
    k = kern()
    ybuf = to_nhwc(bf16r(_rand(N, ld, H, W, seed=11)))
    y = ybuf[..., :C]
    gamma = (1 + 0.1 * _rand(C, seed=3)).requires_grad_(True)
    beta = (0.1 * _rand(C, seed=4)).requires_grad_(True)
    yq = to_nchw(y.contiguous()).requires_grad_(True)
    ref = F.relu(F.batch_norm(yq, None, None, gamma, beta, True, 0.01, 1e-3))
    da = bf16r(_rand(N, C, H, W, seed=7))
    ref.backward(da)
    mean = yq.detach().mean((0, 2, 3))
    invstd = (yq.detach().var((0, 2, 3), unbiased=False) + 1e-3).rsqrt()
    scale = gamma.detach() * invstd
    shift = beta.detach() - mean * scale
    dout = to_nhwc(da)
    prev, prev_min = k._BN_BWD_COOP, k._BN_BWD_COOP_MIN_PIX
    try:
        k._BN_BWD_COOP, k._BN_BWD_COOP_MIN_PIX = True, 1
        dy, dgamma, dbeta = k.bn_relu_bwd_train(dout, y, scale, shift, gamma.detach(), mean, invstd)
        k._BN_BWD_COOP = False
        dy_u, dgamma_u, dbeta_u = k.bn_relu_bwd_train(dout, y, scale, shift, gamma.detach(), mean, invstd)
    finally:
        k._BN_BWD_COOP, k._BN_BWD_COOP_MIN_PIX = prev, prev_min
    assert rel_err(to_nchw(dy), yq.grad) < 6e-3
    torch.testing.assert_close(dgamma, gamma.grad, rtol=2e-4, atol=2e-4 * float(gamma.grad.abs().max()))
    torch.testing.assert_close(dbeta, beta.grad, rtol=2e-4, atol=2e-4 * float(beta.grad.abs().max()))
    torch.testing.assert_close(dgamma, dgamma_u, rtol=2e-4, atol=2e-4 * float(dgamma_u.abs().max()))
    assert float((dy.float() - dy_u.float()).abs().max()) <= 2 ** -7 * float(dy_u.float().abs().max())
