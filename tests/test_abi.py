"""The C-ABI library loads on a CPU-only box and exports every symbol include/ccrs_b200.h declares;
without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, have_gpu


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "ccrs_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    # function declarations only (not the callback members of ccrs_backend)
    return sorted(set(re.findall(r"^\s*(?:const\s+)?[A-Za-z_0-9]+\*?\s+\*?(ccrs_[a-z0-9_]+)\s*\(", txt, flags=re.M)))


def test_header_symbols_all_exported(pkg):
    lib = pkg._abi.load()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ccrs_b200.h but not exported"
    assert sorted(pkg.SYMBOLS) == declared


def test_model_nparams_and_bounds(pkg):
    lib = pkg._abi.load()
    assert [lib.ccrs_model_nparams(m) for m in range(6)] == [5, 6, 8, 8, 9, 8]
    assert lib.ccrs_model_nparams(17) == -1
    lo, hi = pkg.model_bounds("eucm", 1024, 768)
    assert (lo[0], hi[0], hi[2], hi[3]) == (0.0, 1e4, 1024.0, 768.0)   # util.rs:36-39
    assert 0 < lo[4] < hi[4] <= 1.0
    import numpy as np
    for m in ("kb4", "opencv5", "ftheta"):      # polynomial coefficients are not boxed (no invented [-1, 1] clamp)
        lo, hi = pkg.model_bounds(m, 1024, 768)
        assert np.all(np.isneginf(lo[4:])) and np.all(np.isposinf(hi[4:]))
    lo, hi = pkg.model_bounds("eucmt", 1024, 768)
    assert np.all(np.isinf(lo[6:])) and np.all(np.isinf(hi[6:])) and hi[4] == 1.0


def test_default_options_are_tiny_solver_defaults(pkg):
    o = pkg.default_options()
    assert (o.max_iteration, o.min_abs_decrease, o.min_rel_decrease, o.min_error) == (100, 1e-5, 1e-5, 1e-10)
    assert (o.lm_initial_radius, o.lm_min_diag, o.lm_max_diag) == (1e4, 1e-6, 1e32)


@pytest.mark.skipif(have_gpu(), reason="CPU-only behaviour")
def test_no_cpu_fallback(pkg):
    s = pkg.synth.make_calib("eucm", 5, seed=0)
    with pytest.raises(pkg.CcrsError) as e:
        pkg.Problem.from_synth(s)
    assert e.value.code == -3  # CCRS_ERR_NO_DEVICE
    with pytest.raises(pkg.CcrsError):
        pkg.measure_fp64_peak(0)


def test_invalid_arguments_are_rejected(pkg):
    lib = pkg._abi.load()
    h = C.c_void_p()
    fo = np.array([0, 5, 3], dtype=np.int32)  # not monotone
    z = np.zeros(5)
    dp = z.ctypes.data_as(C.POINTER(C.c_double))
    code = lib.ccrs_problem_create(C.byref(h), 1, 10, 10, 0, 2, fo.ctypes.data_as(C.POINTER(C.c_int32)), dp, dp, dp, dp, dp, 1.0, 0)
    assert code == -1 and b"monotone" in lib.ccrs_last_error()
    code = lib.ccrs_problem_create(C.byref(h), 9, 10, 10, 0, 2, fo.ctypes.data_as(C.POINTER(C.c_int32)), dp, dp, dp, dp, dp, 1.0, 0)
    assert code == -1
