"""The C-ABI library loads and exports every symbol include/mgplr.h declares (no compute calls: CPU only)."""
import ctypes
import os
import re

from conftest import ROOT


def test_library_exports_header_symbols():
    from dcd_isaac_b200 import _lib
    if not os.path.exists(_lib.SO_PATH):
        _lib.build()
    L = ctypes.CDLL(_lib.SO_PATH)
    hdr = open(os.path.join(ROOT, 'include', 'mgplr.h')).read()
    declared = set(re.findall(r'\b(mgplr_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for sym in declared:
        assert hasattr(L, sym), 'libmgplr.so does not export %s' % sym
    assert L.mgplr_abi_version() == 4


def test_no_cpu_fallback_without_device():
    """Creating a venv on a box without a CUDA device must fail loudly."""
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    from dcd_isaac_b200 import _lib
    from dcd_isaac_b200.vec_env import CudaAdversarialVecEnv
    with pytest.raises(_lib.MgplrError):
        CudaAdversarialVecEnv('MultiGrid-GoalLastAdversarial-v0', 4)


def test_product_never_imports_oracle():
    bad = []
    pkg = os.path.join(ROOT, 'dcd_isaac_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r'^\s*(from|import)\s+oracle', src, re.M) or 'mg_oracle' in src:
                    bad.append(f)
    assert not bad, bad
