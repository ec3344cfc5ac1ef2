"""Load the UNMODIFIED reference modules from ``/root/reference`` under a private namespace (test infra).

The reference uses bare top-level module names (``utils``, ``torchlayers``, ``models.phiseg`` ...,
e.g. models/phiseg.py:4,11) -- the same names the B200 drop-in exports -- so both cannot live in
``sys.modules`` at once.  ``load_reference()`` imports the reference with the shims of
``oracle/shims`` on the path, then detaches every module it pulled in from ``sys.modules`` and
returns them in a namespace object.  The modules keep working because they hold direct references
to each other.

``/root/reference`` exists only in the build container; on the GPU box ``have_reference()`` is False
and everything falls back to the restatement in ``oracle/phiseg_oracle.py`` + committed goldens.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('UNETZOO_REFERENCE_ROOT', '/root/reference')
SHIM_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shims')

_REF_TOPLEVEL = ('utils', 'torchlayers', 'models', 'data', 'config', 'train_model', 'test_model')
_SHIM_TOPLEVEL = ('revtorch', 'medpy', 'nibabel', 'h5py', 'matplotlib')


def have_reference():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'models', 'phiseg.py'))


def _is_ours(name, tops):
    return name.split('.')[0] in tops


def _detach(tops):
    taken = {}
    for name in list(sys.modules):
        if _is_ours(name, tops):
            taken[name] = sys.modules.pop(name)
    return taken


_cache = None


def load_reference(extra=()):
    """Returns a namespace with attributes utils, torchlayers, phiseg, unet, probabilistic_unet,
    phiseg3D (the reference modules).  ``extra``: more dotted module names to import."""
    global _cache
    if _cache is not None and not extra:
        return _cache
    if not have_reference():
        raise RuntimeError('reference not available at %s' % REFERENCE_ROOT)
    # park whatever currently owns these names (e.g. the drop-in)
    parked = _detach(_REF_TOPLEVEL)
    have_shim = {}
    for s in _SHIM_TOPLEVEL:
        try:
            importlib.import_module(s)
            have_shim[s] = True
        except Exception:
            have_shim[s] = False
    old_path = list(sys.path)
    sys.path.insert(0, REFERENCE_ROOT)
    sys.path.insert(0, SHIM_ROOT)  # real packages win only if already imported above
    try:
        ns = types.SimpleNamespace()
        ns.utils = importlib.import_module('utils')
        ns.torchlayers = importlib.import_module('torchlayers')
        ns.phiseg = importlib.import_module('models.phiseg')
        ns.unet = importlib.import_module('models.unet')
        ns.probabilistic_unet = importlib.import_module('models.probabilistic_unet')
        ns.phiseg3D = importlib.import_module('models.phiseg3D')
        for name in extra:
            setattr(ns, name.replace('.', '_'), importlib.import_module(name))
    finally:
        sys.path[:] = old_path
        ns_modules = _detach(_REF_TOPLEVEL)
        # shims that stood in for missing packages are private to the reference too
        for s in _SHIM_TOPLEVEL:
            if not have_shim[s]:
                ns_modules.update(_detach((s,)))
        sys.modules.update(parked)
    ns._modules = ns_modules
    if not extra:
        _cache = ns
    return ns
