#!/usr/bin/env python
"""Run the reference's UNMODIFIED train_model.py / test_model.py on the B200 drop-in modules.

    python unet-zoo_b200/launch.py --reference /path/to/UNet-Zoo [--script train_model.py] EXP_PATH LOCAL dummy

What it does (SURVEY.md 8b "Environment facts"):
  * puts this directory FIRST on sys.path, so ``import utils``, ``import torchlayers``, ``from models.phiseg import
    PHISeg`` resolve to the drop-ins, and the reference checkout AFTER it, so ``models.experiments.*``, ``data.*``,
    ``config.*`` and the caller scripts come from the reference untouched (``models`` is a namespace package);
  * stands in for third-party imports that are missing here (compat/), accepts the removed ``verbose=`` keyword of
    ``ReduceLROnPlateau`` (train_model.py:50-51 vs torch >= 2.7), and provides a ``config.system`` /
    ``config.local_config`` with a writable, per-rank ``log_root`` (every rank runs validate()/save_model());
  * under torchrun it initialises NCCL, pins cuda:LOCAL_RANK and attaches the gradient all-reduce to the model the
    caller builds (one process per GPU);
  * then ``runpy``-executes the reference script as ``__main__``.
"""
import argparse
import importlib
import os
import runpy
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))


def _ensure_importable(name):
    try:
        importlib.import_module(name)
    except Exception:
        compat = os.path.join(HERE, 'compat')
        if compat not in sys.path:
            sys.path.append(compat)
        importlib.import_module(name)


def _patch_scheduler():
    import inspect
    import torch
    cls = torch.optim.lr_scheduler.ReduceLROnPlateau
    if 'verbose' in inspect.signature(cls.__init__).parameters:
        return
    orig = cls.__init__

    def init(self, *a, verbose=False, **k):
        orig(self, *a, **k)

    cls.__init__ = init


def _inject_sys_config(log_root, reference_root):
    rank = int(os.environ.get('RANK', '0'))
    root = log_root if rank == 0 else os.path.join(log_root, 'rank%d' % rank)
    os.makedirs(root, exist_ok=True)
    for modname in ('config.system', 'config.local_config'):
        m = types.ModuleType(modname)
        m.at_biwi = False
        m.project_root = reference_root
        m.log_root = root
        m.data_root = os.path.join(root, 'data_lidc.pickle')
        m.preproc_folder = os.path.join(root, 'preproc')
        m.dummy_data_root = None
        sys.modules[modname] = m
    pkg = types.ModuleType('config')
    pkg.__path__ = [os.path.join(reference_root, 'config')]
    pkg.system = sys.modules['config.system']
    pkg.local_config = sys.modules['config.local_config']
    sys.modules['config'] = pkg


def _attach_data_parallel():
    """torchrun launch: wrap nn.Module.to so the model built by the unmodified caller gets its gradient all-reduce
    (hooks + finish-before-optimizer-step) without touching train_model.py."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1:
        return
    import torch
    from b200 import dp
    dp.init_from_env('nccl')
    from models.phiseg import PHISeg
    orig_step = torch.optim.Adam.step
    state = {}

    orig_to = PHISeg.to

    def to(self, *a, **k):
        out = orig_to(self, *a, **k)
        if 'ar' not in state:
            state['ar'] = dp.GradientAllReduce(self.parameters())
            state['steps'] = 0
        return out

    def step(self, *a, **k):
        ar = state.get('ar')
        if ar is not None:
            ar.finish()
            state['steps'] += 1
            if state['steps'] == 1:
                out = orig_step(self, *a, **k)
                ar.freeze_buckets()
                self.zero_grad = lambda set_to_none=True: ar.zero_grad()
                return out
        return orig_step(self, *a, **k)

    PHISeg.to = to
    torch.optim.Adam.step = step


def _install_fused_adam():
    """--graph: the caller's ``torch.optim.Adam(net.parameters(), lr=, weight_decay=)`` (train_model.py:49) resolves to
    b200.optim.FusedAdam -- same update rule and state_dict layout, ONE launch for all parameters instead of ~100
    multi-tensor launches and per-parameter Python work (which would leave the graph-replayed step host-bound)."""
    import torch
    from b200.optim import FusedAdam
    stock = torch.optim.Adam

    class Adam(FusedAdam):
        def __new__(cls, params, *a, **k):
            params = list(params)
            if params and all(torch.is_tensor(p) and p.is_cuda for p in params) and not k.get('amsgrad', False):
                return object.__new__(cls)
            return stock(params, *a, **k)

        def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, **unused):
            FusedAdam.__init__(self, params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)

    torch.optim.Adam = Adam


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default=os.environ.get('UNETZOO_REFERENCE_ROOT'),
                    help='path to a checkout of gigantenbein/UNet-Zoo (unmodified)')
    ap.add_argument('--script', default='train_model.py')
    ap.add_argument('--log-root', default=os.environ.get('UNETZOO_LOG_ROOT', os.path.join(os.getcwd(), 'unetzoo_logs')))
    ap.add_argument('--graph', action='store_true',
                    help='transparent CUDA-graph capture of the training step behind net.forward / loss.backward '
                         '(UNETZOO_TRANSPARENT_GRAPH=1): the unmodified loop stops being host-bound')
    ap.add_argument('script_args', nargs=argparse.REMAINDER)
    args = ap.parse_args()
    if args.graph:
        os.environ['UNETZOO_TRANSPARENT_GRAPH'] = '1'    # read when models.phiseg is imported by the experiment file
    if not args.reference or not os.path.isfile(os.path.join(args.reference, args.script)):
        raise SystemExit('launch.py: --reference must point at a UNet-Zoo checkout containing %s' % args.script)
    ref = os.path.abspath(args.reference)
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or '.') not in (HERE, ref)]
    sys.path.insert(0, HERE)
    sys.path.insert(1, ref)
    for name in ('medpy.metric', 'nibabel', 'h5py', 'matplotlib.pyplot', 'skimage.measure', 'skimage.transform'):
        _ensure_importable(name)
    _patch_scheduler()
    _inject_sys_config(os.path.abspath(args.log_root), ref)
    _attach_data_parallel()
    if args.graph and int(os.environ.get('WORLD_SIZE', '1')) == 1:
        _install_fused_adam()
    script_args = [a for a in args.script_args if a != '--']
    sys.argv = [os.path.join(ref, args.script)] + script_args
    runpy.run_path(os.path.join(ref, args.script), run_name='__main__')


if __name__ == '__main__':
    main()
