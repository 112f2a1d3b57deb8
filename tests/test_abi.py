"""CPU tests: the C-ABI library builds, loads and exports every symbol include/mimeo_b200.h declares."""
import os
import re

import pytest

from tests.helpers import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, 'include', 'mimeo_b200.h')).read()
    return sorted(set(re.findall(r'MB2_API[^;(]*?\b(mb2_\w+)\s*\(', txt)))


def test_header_declares_something():
    syms = declared_symbols()
    assert 'mb2_init' in syms and 'mb2_coverage_segments' in syms


def test_library_exports_every_declared_symbol():
    from mimeo_b200 import build, _lib
    build.build_library()
    l = _lib.lib()
    for s in declared_symbols():
        assert hasattr(l, s), f'{s} is declared in include/mimeo_b200.h but not exported'
        assert s in _lib.SIGNATURES, f'{s} has no ctypes signature in mimeo_b200/_lib.py'
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_no_cpu_fallback_without_gpu():
    """On a box without a GPU the product path must raise, not silently compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from mimeo_b200 import coverage, _lib
    with pytest.raises(_lib.Mb2Error):
        coverage.coverage_segments([0], [1], [5], [100], 1, 1)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, 'mimeo_b200')
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dp, fn)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, re.M), f'{fn} imports the oracle'
                for needle in ('libannot_oracle', 'liblastz_oracle', 'oracle/_build', 'oracle/_ref', 'annot_oracle'):
                    assert needle not in txt, f'{fn} references {needle}'
