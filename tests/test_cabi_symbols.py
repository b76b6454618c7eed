"""CPU-only: the C-ABI library builds/loads and exports every symbol include/mn_b200.h declares
(no compute calls: there is no GPU here)."""
import os
import re

import cases  # noqa: F401  (sys.path setup)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'mn_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(mn_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_header_symbols():
    from mega_nerf_b200 import build, _cabi
    build.build()
    lib = _cabi.load_library()
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f'{n} declared in mn_b200.h but not exported'
        assert n in _cabi.SIGNATURES, f'{n} has no ctypes signature'
    assert lib.mn_abi_version() == 1


def test_no_cpu_fallback():
    import pytest
    import torch
    import mega_nerf_b200 as M
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError):
        M.get_ray_directions(4, 4, 1.0, 1.0, 2.0, 2.0, True, torch.device('cpu'))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'mega_nerf_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            assert 'oracle' not in open(os.path.join(pkg, fn)).read(), fn
