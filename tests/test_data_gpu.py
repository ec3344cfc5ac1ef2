"""Device-side synthetic data generator (uz_synth_lidc_batch) against its numpy restatement:
labels / masks bit exact (integer work), patches to transcendental round-off; statistics of the generated data."""
import numpy as np
import pytest
import torch

from tests.gpu_util import PKG  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('seed,size,m', [(0, 128, 4), (12345, 64, 4), (7, 96, 6)])
def test_device_batch_matches_numpy_restatement(seed, size, m):
    from b200 import data
    gen = data.DeviceLIDC(size=size, annotators=m, seed=seed)
    patch, labels, mask = gen.batch(5, seed=seed)
    rp, rl, rm = data.reference_batch(seed, 5, size, m)
    assert np.array_equal(labels.cpu().numpy(), rl)
    assert np.array_equal(mask.cpu().numpy(), rm)
    np.testing.assert_allclose(patch.cpu().numpy(), rp, rtol=0, atol=2e-6)


def test_device_batches_look_like_lidc():
    from b200 import data
    gen = data.DeviceLIDC(seed=3)
    patch, labels, mask = gen.batch(64)
    p2, _, _ = gen.batch(64)
    assert not torch.equal(patch, p2)                         # the counter advances the seed
    lab = labels.float()
    fg = lab.mean().item()
    empty = (lab.sum((1, 2)) == 0).float().mean().item()
    assert 0.01 < fg < 0.15 and 0.1 < empty < 0.45            # a few percent foreground, ~25 % empty annotations
    assert -0.5 <= patch.min().item() and patch.max().item() <= 0.7 + 1e-6
    assert set(np.unique(mask.cpu().numpy())) <= {0.0, 1.0}


def test_fill_feeds_a_training_step_without_host_data():
    from b200 import build, data, synth, train
    net = build.phiseg([16, 32, 32, 32, 32, 32, 32])
    net.load_state_dict(synth.synth_state_dict(net.state_dict(), seed=1))
    net = net.cuda()
    step = train.TrainStep(net, train.make_adam(net), 4, (1, 128, 128), use_graph=True)
    gen = data.DeviceLIDC(seed=5)
    gen.fill(step)
    step.prepare(warmup=1)
    losses = []
    for _ in range(3):
        gen.fill(step)
        step.step_device()
        losses.append(float(step.loss))
    assert all(np.isfinite(losses)) and len(set(losses)) == 3
