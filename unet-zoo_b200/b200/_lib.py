"""ctypes binding of the C ABI declared in include/unetzoo_b200.h.

The prototypes are parsed from the header itself, so the Python side cannot drift from the ABI.  Loading fails loudly:
there is no fallback implementation (SURVEY.md 8b error convention: C returns int status, Python raises).
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(_HERE)
REPO_ROOT = os.path.dirname(PKG_ROOT)
HEADER = os.path.join(REPO_ROOT, 'include', 'unetzoo_b200.h')
# Two builds of the same sources: bf16 storage (the product) and IEEE-half storage (10-bit mantissa like TF32: the
# tolerance-matched parity mode).  UNETZOO_PRECISION=fp16 or set_precision('fp16') selects the second one.
_LIBS = {'bf16': 'libunetzoo_b200.so', 'fp16': 'libunetzoo_b200_fp16.so',
         'prof': 'libunetzoo_b200_prof.so'}      # prof: bf16 + profiling knobs / phase traces (`make prof`, tools only)
PRECISION = os.environ.get('UNETZOO_PRECISION', 'bf16')
if PRECISION not in _LIBS:
    raise ValueError('UNETZOO_PRECISION must be bf16, fp16 or prof (got %r)' % PRECISION)
LIB_PATH = os.path.join(PKG_ROOT, _LIBS[PRECISION])

_SCALARS = {
    'int': ctypes.c_int,
    'long long': ctypes.c_longlong,
    'float': ctypes.c_float,
    'double': ctypes.c_double,
    'unsigned int': ctypes.c_uint,
}


def parse_header(path=HEADER):
    """Returns {name: (restype, [(argtype, argname), ...])} for every prototype in the header."""
    text = open(path).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    text = re.sub(r'//[^\n]*', '', text)
    text = re.sub(r'^\s*#.*$', '', text, flags=re.M)
    protos = {}
    for m in re.finditer(r'([A-Za-z_][\w\s\*]*?)\b(uz_\w+)\s*\(([^)]*)\)\s*;', text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret == 'const char*' or ret == 'const char *':
            restype = ctypes.c_char_p
        else:
            restype = _SCALARS[ret]
        argl = []
        if args and args != 'void':
            for a in args.split(','):
                a = ' '.join(a.split())
                if '*' in a:
                    argl.append((ctypes.c_void_p, a.split('*')[-1].strip()))
                else:
                    ty, an = a.rsplit(' ', 1)
                    argl.append((_SCALARS[ty.replace('const ', '').strip()], an))
        protos[name] = (restype, argl)
    return protos


class UnetZooLibError(RuntimeError):
    pass


_lib = None
_protos = None
_loaded = {}


def act_dtype():
    """torch dtype of activations / packed weights for the selected library"""
    import torch
    return torch.float16 if PRECISION == 'fp16' else torch.bfloat16


def set_precision(name):
    """switch the active library ('bf16' | 'fp16'); tensors produced under one precision must not be fed to the other.
    Returns the previous name."""
    global PRECISION, LIB_PATH, _lib
    if name not in _LIBS:
        raise ValueError('precision must be bf16, fp16 or prof (got %r)' % (name,))
    prev = PRECISION
    if name != prev:
        PRECISION = name
        LIB_PATH = os.path.join(PKG_ROOT, _LIBS[name])
        _lib = _loaded.get(name)
        _fn.clear()
    return prev


def load():
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise UnetZooLibError(
            'library not found at %s -- build it with `python -c "import __graft_entry__ as g; g.build()"` '
            'or `make -C unet-zoo_b200/csrc`; there is no fallback path.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    _protos = parse_header()
    for name, (restype, argl) in _protos.items():
        fn = getattr(lib, name)          # AttributeError => header / library mismatch, fail loudly
        fn.restype = restype
        fn.argtypes = [t for t, _ in argl]
    if lib.uz_abi_version() != 1:
        raise UnetZooLibError('ABI version mismatch')
    if lib.uz_storage_dtype() != (1 if PRECISION == 'fp16' else 0):
        raise UnetZooLibError('%s was not built for %s storage' % (LIB_PATH, PRECISION))
    _lib = lib
    _loaded[PRECISION] = lib
    return lib


SKIP = set()      # measurement aid (tools/family_times.py): entry points whose launches are elided


_fn = {}          # entry point name -> bound ctypes function (attribute lookup on the CDLL is not free)


def call(name, *args):
    """Invoke an int-status entry point; raises with uz_last_error() on failure."""
    if SKIP and name in SKIP:
        return
    fn = _fn.get(name)
    if fn is None:
        fn = _fn[name] = getattr(load(), name)
    rc = fn(*args)
    if rc != 0:
        raise UnetZooLibError('%s failed (%d): %s' % (name, rc, load().uz_last_error().decode()))


def raw(name):
    return getattr(load(), name)
