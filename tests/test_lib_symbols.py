"""The C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol include/xtd_sigma.h declares."""
import ctypes
import os
import re

import pytest

from xtddft_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    return build.build()


def test_header_symbols_are_exported(lib_path):
    hdr = open(os.path.join(ROOT, "include", "xtd_sigma.h")).read()
    declared = set(re.findall(r"\b(xtd_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"xtd_engine"}
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_loader_binds_everything(lib_path):
    lib = _lib.load(lib_path)
    assert lib.xtd_version() >= 100
    assert isinstance(lib.xtd_last_error(), bytes)


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from xtddft_b200.engine import SigmaEngine
    from xtddft_b200.plan import build_sf_plan
    from xtddft_b200.synth import make_problem
    p = make_problem(8, 3, 2, 3, 6, 16, seed=1)
    with pytest.raises(_lib.XtdError):
        SigmaEngine.from_problem(build_sf_plan(p), p)
