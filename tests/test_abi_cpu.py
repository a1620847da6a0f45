"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports exactly the entry points that
include/progen_b200.h declares (no compute is launched without a GPU)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from progen_b200 import lib as L
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__ as G
        G.build()
    return L


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'progen_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return set(re.findall(r'\b(progen_[a-z0-9_]+)\s*\(', src))


def test_library_exports_every_declared_symbol(lib):
    handle = lib.load()
    declared = header_functions()
    assert declared, 'no declarations parsed from include/progen_b200.h'
    for name in declared:
        assert hasattr(handle, name), f'{name} declared in the header but not exported by the .so'
    assert declared == set(lib.PROTOTYPES), (declared ^ set(lib.PROTOTYPES))


def test_version_and_error_text(lib):
    assert 'sm_100a' in lib.version()
    assert lib.load().progen_last_error() is not None


def test_product_has_no_oracle_or_cpu_fallback():
    """The product path must not import the oracle and must fail loudly without a GPU."""
    import torch
    pkg = os.path.join(ROOT, 'progen_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+\.*oracle', text, flags=re.M), f'{f} imports the oracle'
                assert 'progen_ref' not in text and 'progen_torch' not in text, f'{f} references the oracle'
    if not torch.cuda.is_available():
        from progen_b200 import lib as L
        with pytest.raises(L.ProgenError):
            L.require_device()
