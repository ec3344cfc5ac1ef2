"""The C-ABI library loads without a GPU and exports every symbol include/unetzoo_b200.h declares; host-side planning
entry points work on CPU; compute wrappers refuse CPU tensors (no fallback)."""
import ctypes
import os

import pytest
import torch

from tests.gpu_util import PKG  # noqa: F401


def test_library_exports_every_declared_symbol():
    from b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 40
    lib = _lib.load()
    for name in protos:
        assert hasattr(lib, name), name
    assert lib.uz_abi_version() == 1
    assert lib.uz_last_error() is not None
    for name in ('uz_conv_fwd', 'uz_conv_wgrad', 'uz_bn_finalize', 'uz_head_fwd', 'uz_kl_fwd', 'uz_residual_ce',
                 'uz_ged_pairwise', 'uz_variance_ncc', 'uz_pack_conv_weights_batched'):
        assert name in protos


def test_host_side_planning_entry_points():
    from b200 import _lib, kern
    tw, th, tn, nt = kern.conv_tile_geometry(12, 128, 128)
    assert tw * th * tn == 128 and nt == 12 * 128 * 128 // 128
    tw, th, tn, nt = kern.conv_tile_geometry(12, 2, 2)
    assert (tw, th, tn, nt) == (2, 2, 32, 1)
    lib = _lib.load()
    assert lib.uz_wgrad_workspace_floats(12, 128, 128, 128, 128, 9) > 0
    assert lib.uz_wgrad_workspace_floats(12, 128, 128, 100, 128, 9) == -1        # channels must be multiples of 16
    assert lib.uz_conv_uses_persistent_kernel(12, 128, 128, 128, 128, 9) == 1
    assert lib.uz_conv_uses_persistent_kernel(12, 8, 8, 192, 192, 9) == 0         # spatial < 16: generic kernel
    assert lib.uz_conv_uses_persistent_kernel(12, 128, 128, 224, 128, 1) == 0     # 1x1: generic kernel
    assert lib.uz_bn_bwd_num_blocks(196608, 128) > 0


def test_argument_errors_are_reported_not_thrown():
    from b200 import _lib
    lib = _lib.load()
    rc = lib.uz_conv_fwd(None, 1, 16, 16, 16, 16, None, 16, 9, None, 16, None, None, 0, None, None)
    assert rc != 0 and b'null pointer' in lib.uz_last_error()


def test_compute_wrappers_refuse_cpu_tensors():
    from b200 import _lib, kern
    x = torch.zeros(1, 16, 16, 16, dtype=torch.bfloat16)
    w = torch.zeros(9, 16, 16, dtype=torch.bfloat16)
    with pytest.raises(_lib.UnetZooLibError):
        kern.conv_fwd(x, w)
    with pytest.raises(_lib.UnetZooLibError):
        kern.ged(torch.zeros(2, 4, 4, dtype=torch.int64), torch.zeros(2, 4, 4), [1])


def test_dropin_state_dict_and_parameter_count():
    """PHISeg has 24 513 330 parameters / 820 state_dict entries."""
    from tests.keygrammar import dropin_phiseg
    net = dropin_phiseg([32, 64, 128, 192, 192, 192, 192])
    assert sum(p.numel() for p in net.parameters()) == 24513330
    assert len(net.state_dict()) == 820
    assert 'prior.sample_z_path.4.mu_conv.0.weight' in net.state_dict()
    rev = dropin_phiseg([32, 64, 128, 192, 192, 192, 192], reversible=True)
    assert sum(p.numel() for p in rev.parameters()) == 16302290          # RevPHISeg
    from tests.keygrammar import dropin_phiseg3d
    vol = dropin_phiseg3d([32, 64, 128], 3, (4, 128, 128, 128))
    assert sum(p.numel() for p in vol.parameters()) == 9265121           # PHISeg3D [32,64,128], L = 3
    rev3 = dropin_phiseg3d([32, 64, 128], 3, (4, 128, 128, 128), reversible=True)
    assert sum(p.numel() for p in rev3.parameters()) == 2455329         # reversible PHISeg3D (phiseg_brats.py:24)


def test_fused_adam_state_layout_and_cpu_failure():
    """b200.optim.FusedAdam keeps torch.optim.Adam's state layout (state_dicts interchange) and, like every product
    path, refuses CPU parameters instead of falling back."""
    import torch
    from tests.gpu_util import PKG  # noqa: F401
    from b200 import _lib
    from b200.optim import FusedAdam
    p_ref = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
    ref = torch.optim.Adam(p_ref, lr=1e-3, weight_decay=1e-5)
    for p in p_ref:
        p.grad = torch.randn_like(p)
    ref.step()
    p_mine = [torch.nn.Parameter(p.detach().clone()) for p in p_ref]
    mine = FusedAdam(p_mine, lr=1e-3, weight_decay=1e-5)
    mine.load_state_dict(ref.state_dict())
    st = mine.state[p_mine[0]]
    assert set(st) == {'step', 'exp_avg', 'exp_avg_sq'} and float(st['step']) == 1.0
    assert torch.equal(st['exp_avg'], ref.state[p_ref[0]]['exp_avg'])
    assert mine.state_dict()['param_groups'][0]['weight_decay'] == 1e-5
    for p in p_mine:
        p.grad = torch.randn_like(p)
    with pytest.raises(_lib.UnetZooLibError):
        mine.step()
