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
    if kind == "wing":  # swept tapered 3-D wing section (BASELINE config 4 topology): no face aligned with the axes but the symmetry planes
        return cases.naca0012_ogrid(ni=16 * scale, nj=8 * scale, nk=max(nk, 3), span=1.5, sweep=0.5, taper=0.5, radius=8.0, tile=(4, 4, 3))
    if kind == "prism":
        return cases.prism_channel(nx=10 * scale, ny=6 * scale)
    return cases.channel(nx=12 * scale, ny=8 * scale, nz=nk)


def make_bcs(kind, turbulent):
    if kind == "nacawf":  # nutUSpaldingWallFunction on the wing (useWallFunction True in the reference's NACA0012 cases)
        return cases.default_bcs_naca(turbulent=turbulent, wall_function=True)
    return cases.default_bcs_naca(turbulent=turbulent) if kind in ("naca", "wing") else cases.default_bcs_channel(turbulent=turbulent)


def setup(kind="naca", turbulent=True, divU="linearUpwind", nk=2, nres=ALL_RES, lib_path=None, scale=1, binary=False,
          extra_options=None, with_oracle=True, ras_model="SpalartAllmaras"):
    mesh = make_mesh(kind, nk, scale)
    bcs = make_bcs(kind, turbulent)
    d = tempfile.mkdtemp(prefix="dab_case_")
    div_u = "bounded Gauss %s%s" % (divU, " grad(U)" if divU.startswith("linearUpwind") else "")
    cases.write_case(d, mesh, bcs, binary=binary, div_u=div_u, **({} if ras_model == "SpalartAllmaras" else dict(ras_model=ras_model)))
    opts = dict(normalizeStates=NORM_STATES, normalizeResiduals=list(nres))
    opts.update(extra_options or {})
    orc = Oracle(mesh, bcs, normalizeStates=NORM_STATES, divU=divU, normalizeResiduals=nres, rasModel=ras_model) if with_oracle else None
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
    ("naca", True, "linearUpwindV", 1, ALL_RES, "SpalartAllmarasFv3"),  # fv3 production term (the reference's compressible tests use it)
    ("channel", True, "linearUpwind", 2, ALL_RES, "SpalartAllmarasFv3"),
    ("prism", True, "linearUpwind", 1, ALL_RES),    # triangular prisms: 5 faces per cell, triangles + quads
    ("wing", True, "linearUpwindV", 3, ALL_RES),    # swept tapered wing, brick-major numbering: skewed fully 3-D hexahedra
]


def check_parity(lib_path, tol=1e-10):
    """R(W), dRdW^T psi, force and dF/dW of the engine vs the oracle on every configuration."""
    worst = 0.0
    for cfg in CONFIGS:
        kind, turb, divU, nk, nres = cfg[:5]
        ras = cfg[5] if len(cfg) > 5 else "SpalartAllmaras"
        mesh, bcs, orc, sol, W, _ = setup(kind, turb, divU, nk, nres, lib_path, ras_model=ras)
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


def check_functions(lib_path, tol=1e-12):
    """DAFunctionForce (fixedDirection / parallelToFlow / normalToFlow) and DAFunctionMoment of the engine vs
    the oracle: values, [dF/dW]^T and the direct dF/d(aoa) term of the flow-aligned modes."""
    aoa = 3.0
    a = np.deg2rad(aoa)
    dirv = [float(np.cos(0.05)), float(np.sin(0.05)), 0.0]
    axis, center = [0.0, 0.0, 1.0], [0.25, 0.01, 0.05]
    fn = {"CD": {"type": "force", "source": "patchToFace", "patches": ["wing"], "directionMode": "fixedDirection",
                 "direction": dirv, "scale": 0.02},
          "CDa": {"type": "force", "source": "patchToFace", "patches": ["wing"], "directionMode": "parallelToFlow",
                  "patchVelocityInputName": "patchV", "scale": 0.5},
          "CLa": {"type": "force", "source": "patchToFace", "patches": ["wing"], "directionMode": "normalToFlow",
                  "patchVelocityInputName": "patchV", "scale": 0.5},
          "CMZ": {"type": "moment", "source": "patchToFace", "patches": ["wing"], "axis": axis, "center": center,
                  "scale": 2.0}}
    inp = {"patchV": {"type": "patchVelocity", "patches": ["inout"], "flowAxis": "x", "normalAxis": "y"}}
    mesh, bcs, orc, sol, W, _ = setup("naca", True, lib_path=lib_path, extra_options=dict(function=fn, inputInfo=inp))
    sol.updateOFFields(W)
    x = np.array([10.0, aoa])
    one = np.array([1.0])
    dpar = [np.cos(a), np.sin(a), 0.0]
    dnor = [-np.sin(a), np.cos(a), 0.0]
    # the far-field reference velocity changes with the input: keep the oracle consistent
    orc.set_bc_value("U", 1, [10.0 * np.cos(a), 10.0 * np.sin(a), 0.0])
    dFdx = np.zeros(2)
    sol.calcJacTVecProduct("patchV", "patchVelocity", x, "CD", "function", one, dFdx)
    assert np.all(dFdx == 0.0)
    cases = {"CD": (dirv, 0.02, None), "CDa": (dpar, 0.5, None), "CLa": (dnor, 0.5, None), "CMZ": (axis, 2.0, center)}
    for name, (d, scale, ctr) in cases.items():
        F, Fo = sol.calcFunction(name), orc.force(W, 0, d, scale, center=ctr)
        assert abs(F - Fo) <= tol * abs(Fo), (name, F, Fo)
        prod = np.zeros(orc.ndof)
        sol.calcJacTVecProduct("states", "stateVar", W, name, "function", one, prod)
        assert rel_err(prod, orc.dforce_dw(W, 0, d, scale, center=ctr)) < tol, name
    # d(force . dir(aoa))/d aoa[deg] at fixed states: d(par)/da = nor, d(nor)/da = -par
    for name, d, sgn in (("CDa", dnor, 1.0), ("CLa", dpar, -1.0)):
        sol.calcJacTVecProduct("patchV", "patchVelocity", x, name, "function", one, dFdx)
        ref = sgn * orc.force(W, 0, d, 0.5) * np.pi / 180.0
        assert dFdx[0] == 0.0 and abs(dFdx[1] - ref) <= tol * abs(ref), (name, dFdx, ref)
    sol.calcJacTVecProduct("patchV", "patchVelocity", x, "CMZ", "function", one, dFdx)
    assert np.all(dFdx == 0.0)
    return True


def mrf_zone(mesh, wall="wing", omega=40.0, axis=(0.0, 0.1, 1.0)):
    """A rotating cellZone for the MRF tests: the cells within a radius of the aerofoil (centres estimated from the quad-face
    means); every patch but the wall itself is a nonRotatingPatch."""
    Sf, Cf = cases.quad_face_geometry(mesh)
    nC = mesh.n_cells
    cc, cnt = np.zeros((nC, 3)), np.zeros(nC)
    np.add.at(cc, mesh.owner, Cf)
    np.add.at(cnt, mesh.owner, 1.0)
    np.add.at(cc, mesh.neighbour, Cf[:mesh.n_internal_faces])
    np.add.at(cnt, mesh.neighbour, 1.0)
    cc /= cnt[:, None]
    origin = np.array([0.3, 0.02, 0.0])
    r = np.linalg.norm((cc - origin)[:, :2], axis=1)
    cells = np.nonzero(r < 0.45 * r.max())[0]
    return dict(cellZone="rotor", cells=cells, origin=origin.tolist(), axis=list(axis), omega=omega,
                nonRotatingPatches=[p["name"] for p in mesh.patches if p["name"] != wall])
