"""CPU tests of the C-ABI boundary (no GPU): the shared library loads, exports every symbol that
include/xna_basecaller.h declares, and refuses to compute without a device."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from xna_basecaller_b200 import build, _lib
    build.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'xna_basecaller.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(xb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol(lib):
    from xna_basecaller_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(_lib.SYMBOLS) == names           # the ctypes table and the header stay in step


def test_abi_version_and_error_path(lib):
    assert lib.xb_abi_version() == 2
    if torch.cuda.is_available():
        pytest.skip('error path without a device is only observable on a CPU-only host')
    h = ctypes.c_void_p()
    rc = lib.xb_create(ctypes.byref(h), 0, 4, 16, 5, 3, b'NACGTX', 0)
    assert rc < 0 and not h.value
    msg = lib.xb_last_error(None).decode()
    assert 'no CUDA device' in msg and 'no CPU path' in msg
    rc = lib.xb_create(ctypes.byref(h), 0, 4, 16, 7, 3, b'NACGTXYZ', 0)     # unsupported alphabet: argument error first
    assert rc == -4
    assert lib.xb_create(ctypes.byref(h), 0, 0, 16, 5, 3, b'NACGTX', 0) == -1
    assert lib.xb_destroy(None) == 0
    assert lib.xb_launch_count(None) == 0
