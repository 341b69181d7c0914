"""Shared helpers of the test-suite: build a synthetic case, its oracle and a solver on it."""
import os
import tempfile

import numpy as np

from dafoam_b200 import cases
from dafoam_b200.pyDASolvers import pyDASolvers
from oracle.pyoracle import Oracle, synthetic_state

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTSIM = os.path.join(ROOT, "tests", "hostsim", "libdab200_hostsim.so")
NORM_STATES = dict(U=10.0, p=50.0, nuTilda=1e-3, phi=1.0)
ALL_RES = ("URes", "pRes", "nuTildaRes", "phiRes")


def make_mesh(kind, nk=2, scale=1):
    if kind in ("naca", "nacawf"):
        return cases.naca0012_ogrid(ni=40 * scale, nj=20 * scale, nk=nk)
    if kind == "prism":
        return cases.prism_channel(nx=10 * scale, ny=6 * scale)
    return cases.channel(nx=12 * scale, ny=8 * scale, nz=nk)


def make_bcs(kind, turbulent):
    if kind == "nacawf":  # nutUSpaldingWallFunction on the wing (useWallFunction True in the reference's NACA0012 cases)
        return cases.default_bcs_naca(turbulent=turbulent, wall_function=True)
    return cases.default_bcs_naca(turbulent=turbulent) if kind == "naca" else cases.default_bcs_channel(turbulent=turbulent)


def setup(kind="naca", turbulent=True, divU="linearUpwind", nk=2, nres=ALL_RES, lib_path=None, scale=1, binary=False,
          extra_options=None, with_oracle=True):
    mesh = make_mesh(kind, nk, scale)
    bcs = make_bcs(kind, turbulent)
    d = tempfile.mkdtemp(prefix="dab_case_")
    div_u = "bounded Gauss %s%s" % (divU, " grad(U)" if divU.startswith("linearUpwind") else "")
    cases.write_case(d, mesh, bcs, binary=binary, div_u=div_u)
    opts = dict(normalizeStates=NORM_STATES, normalizeResiduals=list(nres))
    opts.update(extra_options or {})
    orc = Oracle(mesh, bcs, normalizeStates=NORM_STATES, divU=divU, normalizeResiduals=nres) if with_oracle else None
    sol = pyDASolvers("DASimpleFoam -python", opts, caseDir=d, _lib_path=lib_path)
    if orc is not None:
        W = synthetic_state(mesh, orc.geometry("C"), orc.geometry("Sf"), turbulent=turbulent)
    else:
        W = None
    return mesh, bcs, orc, sol, W, d


def segments(mesh, turbulent, ndof):
    nC = mesh.n_cells
    segs = [("U", 0, 3 * nC), ("p", 3 * nC, 4 * nC)]
    if turbulent:
        segs.append(("nuTilda", 4 * nC, 5 * nC))
    segs.append(("phi", (5 if turbulent else 4) * nC, ndof))
    return segs


def rel_err(a, b):
    n = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (n if n > 0 else 1.0))


CONFIGS = [
    # kind, turbulent, divU, nk, normalizeResiduals
    ("naca", True, "linearUpwind", 2, ALL_RES),
    ("naca", False, "linearUpwind", 1, ALL_RES),
    ("channel", True, "linearUpwind", 2, ALL_RES),
    ("channel", False, "upwind", 1, ALL_RES),
    ("channel", True, "linear", 1, ALL_RES),
    ("naca", True, "linearUpwind", 1, ("pRes",)),
    ("naca", True, "upwind", 1, ()),
    ("naca", True, "linearUpwindV", 2, ALL_RES),   # the div(phi,U) scheme of the reference's NACA0012 tutorial cases
    ("channel", True, "linearUpwindV", 1, ALL_RES),
    ("nacawf", True, "linearUpwindV", 1, ALL_RES),  # Spalding wall function (Newton solve per wall face)
    ("prism", True, "linearUpwind", 1, ALL_RES),    # triangular prisms: 5 faces per cell, triangles + quads
]


def check_parity(lib_path, tol=1e-10):
    """R(W), dRdW^T psi, force and dF/dW of the engine vs the oracle on every configuration."""
    worst = 0.0
    for kind, turb, divU, nk, nres in CONFIGS:
        mesh, bcs, orc, sol, W, _ = setup(kind, turb, divU, nk, nres, lib_path)
        assert sol.getNLocalAdjointStates() == orc.ndof
        sol.updateOFFields(W)
        for isPC in (0, 1):
            R = np.zeros(orc.ndof)
            sol.getResiduals(R, isPC)
            Ro = orc.residual(W, isPC)
            for name, a, b in segments(mesh, turb, orc.ndof):
                e = rel_err(R[a:b], Ro[a:b])
                worst = max(worst, e)
                assert e < tol, ("residual", kind, turb, divU, isPC, name, e)
        orc.record(W)
        rng = np.random.default_rng(4321)
        for trial in range(2):
            psi = rng.uniform(-1, 1, orc.ndof) if trial == 0 else np.full(orc.ndof, 1e-3)
            y = np.zeros(orc.ndof)
            sol.calcdRdWTPsiAD(psi, y)
            yo = orc.jtvec(psi)
            for name, a, b in segments(mesh, turb, orc.ndof):
                e = rel_err(y[a:b], yo[a:b])
                worst = max(worst, e)
                assert e < tol, ("jtvec", kind, turb, divU, name, e)
    return worst
