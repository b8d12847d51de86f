"""The C-ABI library loads and exports every symbol include/ace_b200.h declares; without a GPU the compute
entry points fail loudly (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

from additivecausalexpansion_b200 import _lib, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ace_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ace_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    syms = _declared_symbols()
    assert len(syms) >= 45
    raw = ctypes.CDLL(_lib.SO_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/ace_b200.h but not exported"
    # every declared symbol has a typed ctypes signature, and nothing is bound that the header lacks
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)
    assert _lib.lib().ace_version().startswith(b"ace_b200")


def test_reference_exports_have_counterparts():
    """The 19 routines registered in the reference's src/RcppExports.cpp:301-327."""
    ref = ["kernmat_Matern32_cpp", "kernmat_Matern32_symmetric_cpp", "grad_Matern_cpp", "kernmat_SE_cpp",
           "kernmat_SE_symmetric_cpp", "invkernel_cpp", "grad_SE_cpp", "ncs_basis", "ncs_basis_deriv",
           "Nesterov_cpp", "Nadam_cpp", "Adam_cpp", "pred_cpp", "pred_marginal_cpp", "stats_cpp", "mu_solution_cpp",
           "normalize_train", "normalize_test", "norm_clip_cpp"]
    syms = set(_declared_symbols())
    for r in ref:
        assert "ace_" + r in syms
        assert hasattr(api, r)


@pytest.mark.skipif(_lib.lib().ace_device_count() > 0, reason="a CUDA device is present")
def test_no_gpu_means_loud_failure_not_fallback():
    with pytest.raises(_lib.AceError) as ei:
        api.invkernel_cpp(np.eye(4), 0.0)
    assert ei.value.status == -3  # ACE_ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)
    with pytest.raises(_lib.AceError):
        api.kernmat_SE_symmetric_cpp(np.zeros((4, 2)), np.ones((4, 1)), np.zeros(2 + 2 + 4))
    from additivecausalexpansion_b200.fit import AceFit

    with pytest.raises(_lib.AceError):
        AceFit(np.zeros(4), np.zeros((4, 2)), np.ones((4, 1)), np.zeros(8))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "additivecausalexpansion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "ace_oracle" not in txt, f
