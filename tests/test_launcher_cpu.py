"""The launcher runs the reference's UNMODIFIED train_model.py on the drop-in modules.  Without a GPU the run must get
through the caller's whole set-up (imports, experiment file, UNetModel construction with the drop-in PHISeg, stock
Adam, synthetic data plug-in) and stop at the first forward with the explicit no-CPU-fallback error."""
import os
import subprocess
import sys

import pytest

from oracle.ref_loader import REFERENCE_ROOT, have_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not have_reference(), reason='/root/reference only exists in the build container')
def test_unmodified_train_model_reaches_the_device_boundary(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip('CPU-only check')
    exp = os.path.join(ROOT, 'unet-zoo_b200', 'experiments_b200', 'phiseg_7_5_12_synthetic.py')
    cmd = [sys.executable, os.path.join(ROOT, 'unet-zoo_b200', 'launch.py'), '--reference', REFERENCE_ROOT,
           '--log-root', str(tmp_path), exp, 'local', 'dummy']
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=str(tmp_path))
    out = r.stdout + r.stderr
    assert r.returncode != 0
    assert 'no CPU fallback' in out, out[-3000:]
    assert 'Starting training.' in open(os.path.join(str(tmp_path), 'lidc_synthetic', 'PHISeg_7_5_12_synthetic',
                                                     'training_log.log')).read()
