"""CPU: the C-ABI library loads (no GPU needed) and exports every symbol include/b200_imagen.h declares."""
import ctypes
import os
import re

from imagen_pytorch_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'b200_imagen.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(b200_[a-z0-9_]+)\s*\(', src)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/b200_imagen.h but not exported'


def test_python_binding_covers_the_header():
    assert set(_lib.SIGNATURES) | {'b200_last_error'} == set(_declared())


def test_abi_version_and_npad_without_gpu():
    lib = _lib.load()
    assert lib.b200_abi_version() == 2
    assert [_lib.npad(n) for n in (3, 32, 33, 64, 65, 128, 129, 512)] == [32, 32, 64, 64, 128, 128, 256, 512]
    assert lib.b200_gca_nchunk(64) == 1 and lib.b200_gca_nchunk(4096) == 16


def test_struct_sizes_match_the_c_side():
    lib = _lib.load()
    for which, struct in enumerate((_lib.Src, _lib.Seg, _lib.Epilogue, _lib.TimeRowJob, _lib.DdpmCoef, _lib.EdmCoef, _lib.RowChain)):
        assert ctypes.sizeof(struct) == lib.b200_sizeof(which), struct.__name__
