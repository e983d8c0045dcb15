"""CPU: the C-ABI library builds, loads, and exports every symbol declared in include/b200ocl.h
(no compute calls -- there is no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, 'include', 'b200ocl.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(b200ocl_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from b200ocl import _build, _native
    _build.build()
    lib = _native.lib()
    names = _declared_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), 'missing export ' + n
        assert n in _native.SIGNATURES, 'ctypes signature missing for ' + n
    assert set(_native.SIGNATURES) == set(names)
    assert lib.b200ocl_version() >= 100


def test_workspace_queries_are_pure_host_calls():
    from b200ocl import _native
    lib = _native.lib()
    assert lib.b200ocl_knn_sv_workspace_bytes(110, 160, 160) >= 256
    assert lib.b200ocl_supcon_workspace_bytes(110, 2, 128) >= 256 + 2 * 220 * 4


def test_ops_refuse_cpu_tensors():
    import torch
    from b200ocl import _native, ops
    with pytest.raises(_native.NativeError):
        ops.knn_sv(torch.zeros(2, 4), torch.zeros(2, dtype=torch.long), torch.zeros(3, 4),
                   torch.zeros(3, dtype=torch.long), 3)
    with pytest.raises(_native.NativeError):
        ops.supcon(torch.zeros(4, 2, 8), torch.zeros(4, dtype=torch.long), 0.07)
