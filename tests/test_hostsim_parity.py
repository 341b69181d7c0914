"""Kernel functors (compiled for the host: tests/hostsim, TEST-ONLY) vs the oracle.  This is how the
hand-derived reverse sweep is checked on a machine without a GPU; the -m gpu tests repeat the same
checks through the CUDA library."""
import numpy as np

from tests.common import HOSTSIM, check_functions, check_parity, setup, rel_err


def test_forward_and_reverse_parity_host_build():
    worst = check_parity(HOSTSIM, tol=1e-10)
    assert worst < 1e-10


def test_force_function_and_dfdw_host_build():
    dirv = [float(np.cos(0.05)), float(np.sin(0.05)), 0.0]
    fn = {"CD": {"type": "force", "source": "patchToFace", "patches": ["wing"], "directionMode": "fixedDirection",
                 "direction": dirv, "scale": 0.02}}
    mesh, bcs, orc, sol, W, _ = setup("naca", True, lib_path=HOSTSIM, extra_options=dict(function=fn))
    sol.updateOFFields(W)
    F, Fo = sol.calcFunction("CD"), orc.force(W, 0, dirv, 0.02)
    assert abs(F - Fo) <= 1e-12 * abs(Fo)
    prod = np.zeros(orc.ndof)
    sol.calcJacTVecProduct("states", "stateVar", W, "CD", "function", np.array([1.0]), prod)
    assert rel_err(prod, orc.dforce_dw(W, 0, dirv, 0.02)) < 1e-12


def test_force_moment_and_direction_modes_host_build():
    assert check_functions(HOSTSIM)


def test_dot_product_identity_host_build():
    # <J^T psi, v> == <psi, J v> with J v from central differences of the engine's own R(W)
    mesh, bcs, orc, sol, W, _ = setup("channel", True, lib_path=HOSTSIM, nk=1,
                                      extra_options=dict(normalizeStates=dict(U=1.0, p=1.0, nuTilda=1.0, phi=1.0)))
    sol.updateOFFields(W)
    rng = np.random.default_rng(5)
    n = orc.ndof
    psi = rng.uniform(-1, 1, n)
    y = np.zeros(n)
    sol.calcdRdWTPsiAD(psi, y)
    magSf = orc.geometry("magSf")
    y[-mesh.n_faces:] /= magSf  # undo the phi scaling (normalizeStates phi * magSf)
    v = rng.uniform(-1, 1, n) * 1e-3
    eps = 1e-4
    Rp, Rm = np.zeros(n), np.zeros(n)
    sol.updateOFFields(W + eps * v)
    sol.getResiduals(Rp)
    sol.updateOFFields(W - eps * v)
    sol.getResiduals(Rm)
    Jv = (Rp - Rm) / (2 * eps)
    assert abs(psi @ Jv - v @ y) <= 1e-6 * abs(v @ y)


def test_lane_per_face_reva_pilot_matches():
    """DAB_LANES=1 (RevALanes: 8 lanes per cell, butterfly reduction) computes the same transpose product as the cell-per-thread RevA,
    on hexahedral and on triangular-prism cells (5 faces: idle lanes).  The mapping is chosen
    once per process, so the check runs in a child process."""
    import os
    import subprocess
    import sys
    code = r'''
import numpy as np
from tests.common import HOSTSIM, setup, rel_err
worst = 0.0
for kind, turb, nk in (("naca", True, 2), ("channel", True, 1), ("naca", False, 1), ("prism", True, 1)):
    mesh, bcs, orc, sol, W, _ = setup(kind, turb, nk=nk, lib_path=HOSTSIM)
    sol.updateOFFields(W)
    orc.record(W)
    rng = np.random.default_rng(3)
    for _ in range(2):
        psi = rng.uniform(-1, 1, orc.ndof)
        y = np.zeros(orc.ndof)
        sol.calcdRdWTPsiAD(psi, y)
        worst = max(worst, rel_err(y, orc.jtvec(psi)))
print("WORST %.3e" % worst)
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = {}
    for lanes in ("0", "1"):
        env = dict(os.environ, DAB_LANES=lanes, PYTHONPATH=root)
        r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        out[lanes] = float(r.stdout.strip().split("WORST")[-1])
    assert out["0"] < 1e-10 and out["1"] < 1e-10, out
