"""CPU: the C-ABI library loads and exports every symbol include/decnet_b200.h declares,
and the Python binding table covers exactly that set (no compute calls without a GPU)."""
import ctypes
import re

import pytest


def _declared_symbols(repo_root):
    text = (repo_root / "include" / "decnet_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(decnet_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(repo_root):
    from decnet_b200 import _lib
    handle = _lib.lib()
    names = _declared_symbols(repo_root)
    assert len(names) >= 10
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"
    assert handle.decnet_abi_version() == 2


def test_binding_table_matches_header(repo_root):
    from decnet_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols(repo_root)


def test_ops_refuse_cpu_tensors():
    import torch
    from decnet_b200 import SpaMat, _lib
    x = torch.zeros(1, 2, 3, 4)
    m = torch.zeros(1, 3, 4)
    with pytest.raises(_lib.DecnetError):
        SpaMat()(x, x, m, m, 4)


def test_module_api_shape():
    """Same construction contract as the reference (SparseDenseNetRefinementMask.py:67-68):
    no ctor args, no parameters or buffers."""
    from decnet_b200 import SpaMat, SpaVar
    for cls in (SpaMat, SpaVar):
        m = cls()
        assert list(m.state_dict().keys()) == []
