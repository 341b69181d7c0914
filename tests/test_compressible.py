"""DARhoSimpleFoam: the compressible residual kernels (comp_kernels.hpp) against the oracle's
restatement of DAResidualRhoSimpleFoam + updateThermoVars + the compressible SA model, over both energy variables, both
transport models, both SA variants, every div(phi,U) scheme and the wall function."""
import tempfile

import numpy as np
import pytest

from dafoam_b200 import cases
from dafoam_b200.pyDASolvers import pyDASolvers
from oracle.pyoracle import Oracle, synthetic_state
from tests.common import HOSTSIM, rel_err

NS = dict(U=50.0, p=101325.0, T=300.0, nuTilda=1e-3, phi=1.0)
NRES = ("URes", "pRes", "TRes", "nuTildaRes", "phiRes")
CONFIGS = [
    # mesh, energy, transport, RAS model, div(phi,U), div(phi,e|h), wall function, listed residuals
    ("naca", "sensibleInternalEnergy", "const", "SpalartAllmaras", "linearUpwindV", "upwind", False, NRES),
    ("naca", "sensibleEnthalpy", "sutherland", "SpalartAllmarasFv3", "linearUpwind", "linearUpwind", True, NRES),
    ("channel", "sensibleInternalEnergy", "sutherland", "SpalartAllmaras", "upwind", "linear", False, NRES),
    ("prism", "sensibleEnthalpy", "const", "SpalartAllmaras", "linearUpwind", "upwind", False, ("pRes", "TRes")),
]


def setup_comp(cfg, lib_path):
    kind, energy, transport, ras, divU, divE, wf, nres = cfg
    if kind == "naca":
        mesh = cases.naca0012_ogrid(ni=40, nj=20, nk=2)
        base = cases.default_bcs_naca(U0=(50.0, 2.0, 0.0), wall_function=wf)
    elif kind == "prism":
        mesh = cases.prism_channel(nx=10, ny=6)
        base = cases.default_bcs_channel(U0=(50.0, 0.0, 0.0))
    else:
        mesh = cases.channel(nx=12, ny=8, nz=1)
        base = cases.default_bcs_channel(U0=(50.0, 0.0, 0.0))
    th = cases.default_thermo(energy=energy, transport=transport, divE=divE, divEkp="linear" if divE == "linear" else "upwind")
    bcs = cases.compressible_bcs(base)
    d = tempfile.mkdtemp(prefix="dab_comp_")
    div_u = "bounded Gauss %s%s" % (divU, " grad(U)" if divU.startswith("linearUpwind") else "")
    cases.write_case(d, mesh, bcs, div_u=div_u, ras_model=ras, thermo=th)
    orc = Oracle(mesh, bcs, normalizeStates=NS, normalizeResiduals=nres, thermo=th, rasModel=ras, divU=divU)
    sol = pyDASolvers("DARhoSimpleFoam -python", dict(normalizeStates=NS, normalizeResiduals=list(nres)), caseDir=d, _lib_path=lib_path)
    W = synthetic_state(mesh, orc.geometry("C"), orc.geometry("Sf"), U0=(50.0, 2.0, 0.0), thermo=th)
    return mesh, orc, sol, W


def segments(mesh, ndof):
    nC = mesh.n_cells
    return (("U", 0, 3 * nC), ("p", 3 * nC, 4 * nC), ("T", 4 * nC, 5 * nC), ("nuTilda", 5 * nC, 6 * nC), ("phi", 6 * nC, ndof))


def check_forward(lib_path, tol=1e-10):
    for cfg in CONFIGS:
        mesh, orc, sol, W = setup_comp(cfg, lib_path)
        assert sol.getNLocalAdjointStates() == orc.ndof == 6 * mesh.n_cells + mesh.n_faces
        sol.updateOFFields(W)
        W2 = np.zeros(orc.ndof)
        sol.getOFFields(W2)
        assert np.array_equal(W, W2)
        for isPC in (0, 1):
            R = np.zeros(orc.ndof)
            sol.getResiduals(R, isPC)
            Ro = orc.residual(W, isPC)
            for name, a, b in segments(mesh, orc.ndof):
                e = rel_err(R[a:b], Ro[a:b])
                assert e < tol, (cfg[:6], isPC, name, e)


def check_reverse(lib_path, tol=1e-10):
    """dRdW^T psi of the hand-derived compressible reverse sweep (comp_rev_kernels.hpp) vs the oracle's tape."""
    worst = 0.0
    for cfg in CONFIGS:
        mesh, orc, sol, W = setup_comp(cfg, lib_path)
        sol.updateOFFields(W)
        orc.record(W)
        rng = np.random.default_rng(4321)
        for trial in range(2):
            psi = rng.uniform(-1, 1, orc.ndof) if trial == 0 else np.full(orc.ndof, 1e-3)
            y = np.zeros(orc.ndof)
            sol.calcdRdWTPsiAD(psi, y)
            yo = orc.jtvec(psi)
            for name, a, b in segments(mesh, orc.ndof):
                e = rel_err(y[a:b], yo[a:b])
                worst = max(worst, e)
                assert e < tol, (cfg[:6], trial, name, e)
        # linearity and reproducibility of the operator
        a_, b_ = rng.uniform(-1, 1, orc.ndof), rng.uniform(-1, 1, orc.ndof)
        ya, yb, yab, ya2 = (np.zeros(orc.ndof) for _ in range(4))
        sol.calcdRdWTPsiAD(a_, ya)
        sol.calcdRdWTPsiAD(b_, yb)
        sol.calcdRdWTPsiAD(2.0 * a_ - 3.0 * b_, yab)
        sol.calcdRdWTPsiAD(a_, ya2)
        assert np.array_equal(ya, ya2) and rel_err(yab, 2.0 * ya - 3.0 * yb) < 1e-12
    return worst


def check_function_and_adjoint(lib_path):
    """Force / moment (compressible devRhoReff), their state derivatives, and the adjoint solve dRdW^T psi = dF/dW with the
    coloured-FD preconditioner + GMRES on the 6-state layout; the solution is checked against the oracle's own J^T."""
    from dafoam_b200.pyDASolvers import KSP, Mat
    cfg = CONFIGS[0]
    mesh, orc, sol, W = setup_comp(cfg, lib_path)
    d = [0.6, 0.8, 0.0]
    ctr = [0.25, 0.0, 0.05]
    fn = {"CD": {"type": "force", "source": "patchToFace", "patches": ["wing"], "directionMode": "fixedDirection", "direction": d, "scale": 0.01},
          "CM": {"type": "moment", "source": "patchToFace", "patches": ["wing"], "axis": [0.0, 0.0, 1.0], "center": ctr, "scale": 0.02}}
    sol.updateDAOption(dict(normalizeStates=NS, normalizeResiduals=list(cfg[7]), function=fn,
                            adjEqnOption=dict(gmresRelTol=1e-5, gmresMaxIters=800, gmresRestart=800, pcConLevel=3)))
    sol.updateOFFields(W)
    one = np.array([1.0])
    tol = 1e-12 if lib_path is not None else 1e-9  # GPU: FMA contraction on pressures of 1e5
    for name, dirv, scale, c_ in (("CD", d, 0.01, None), ("CM", [0.0, 0.0, 1.0], 0.02, ctr)):
        F, Fo = sol.calcFunction(name), orc.force(W, 0, dirv, scale, center=c_)
        assert abs(Fo) > 0 and abs(F - Fo) <= tol * abs(Fo), (name, F, Fo)
        g = np.zeros(orc.ndof)
        sol.calcJacTVecProduct("states", "stateVar", W, name, "function", one, g)
        go = orc.dforce_dw(W, 0, dirv, scale, center=c_)
        assert np.linalg.norm(go) > 0 and rel_err(g, go) < tol, name
    dFdW, psi = np.zeros(orc.ndof), np.zeros(orc.ndof)
    sol.calcJacTVecProduct("states", "stateVar", W, "CD", "function", one, dFdW)
    pc, ksp = Mat(), KSP()
    sol.calcdRdWT(1, pc)
    sol.createMLRKSPMatrixFree(pc, ksp)
    assert sol.solveLinearEqn(ksp, dFdW, psi) == 0
    orc.record(W)
    assert rel_err(orc.jtvec(psi), dFdW) < 2e-5


def check_patch_velocity(lib_path):
    """[dR/d(|U|, aoa)]^T psi for DARhoSimpleFoam (central differences on the device kernels) vs the oracle's tape."""
    cfg = CONFIGS[0]
    mesh, orc, sol, W = setup_comp(cfg, lib_path)
    inp = {"patchV": {"type": "patchVelocity", "patches": ["inout"], "flowAxis": "x", "normalAxis": "y"}}
    sol.updateDAOption(dict(normalizeStates=NS, normalizeResiduals=list(cfg[7]), inputInfo=inp))
    sol.updateOFFields(W)
    x = np.array([50.0, 3.0])
    a = np.deg2rad(x[1])
    psi = np.random.default_rng(11).uniform(-1, 1, orc.ndof)
    prod = np.zeros(2)
    sol.calcJacTVecProduct("patchV", "patchVelocity", x, "R", "residual", psi, prod)
    ip = [p["name"] for p in mesh.patches].index("inout")
    orc.set_bc_value("U", ip, [x[0] * np.cos(a), x[0] * np.sin(a), 0.0])
    rb = orc.jtvec_bcU(W, psi, ip)
    ref = np.array([rb[0] * np.cos(a) + rb[1] * np.sin(a), (-rb[0] * x[0] * np.sin(a) + rb[1] * x[0] * np.cos(a)) * np.pi / 180.0])
    assert np.allclose(prod, ref, rtol=1e-6), (prod, ref)


def check_primal(lib_path):
    """DARhoSimpleFoam::solvePrimal on the device: the SIMPLE fixed point is the root of the compressible residual (oracle),
    and equals the oracle's Newton-converged state."""
    from tests.test_converged_primal import newton
    host = lib_path is not None  # the GPU run stops one decade earlier (round-off floor of the normalised residuals at p ~ 1e5)
    mesh = cases.channel(nx=14, ny=8, nz=1)
    bcs = cases.compressible_bcs(cases.default_bcs_channel(U0=(60.0, 0.0, 0.0)))
    th = cases.default_thermo()
    d = tempfile.mkdtemp(prefix="dab_cprimal_")
    cases.write_case(d, mesh, bcs, thermo=th)
    sol = pyDASolvers("DARhoSimpleFoam -python", dict(normalizeStates=NS, primalMinResTol=1e-12 if host else 1e-11, primalMaxIters=3000), caseDir=d, _lib_path=lib_path)
    orc = Oracle(mesh, bcs, normalizeStates=NS, normalizeResiduals=NRES, thermo=th)
    n = orc.ndof
    W0 = np.zeros(n)
    sol.getOFFields(W0)
    assert sol.solvePrimal() == 0
    st = sol.primalStats
    assert st.converged == 1 and 10 < st.iterations < 3000
    W = np.zeros(n)
    sol.getOFFields(W)
    r0, r1 = np.linalg.norm(orc.residual(W0)), np.linalg.norm(orc.residual(W))
    assert r1 < (1e-8 if host else 1e-6) * r0, (r0, r1)
    Wn = newton(orc, W.copy() * (1.0 + 1e-6), tol=1e-7 * r0 * 1e-3, maxit=20)
    for name, a, b in segments(mesh, n):
        err = np.linalg.norm(W[a:b] - Wn[a:b]) / np.linalg.norm(Wn[a:b])
        assert err < (1e-7 if host else 1e-6), (name, err)


def test_compressible_primal_fixed_point_host_build():
    check_primal(HOSTSIM)


def test_compressible_patch_velocity_product_host_build():
    check_patch_velocity(HOSTSIM)


def test_compressible_function_and_adjoint_solve_host_build():
    check_function_and_adjoint(HOSTSIM)


def test_compressible_residual_parity_host_build():
    check_forward(HOSTSIM)


def test_compressible_transpose_product_parity_host_build():
    assert check_reverse(HOSTSIM) < 1e-10


@pytest.mark.gpu
def test_compressible_residual_parity_cuda():
    check_forward(None, tol=1e-9)


@pytest.mark.gpu
def test_compressible_primal_fixed_point_cuda():
    check_primal(None)


@pytest.mark.gpu
def test_compressible_patch_velocity_product_cuda():
    check_patch_velocity(None)


@pytest.mark.gpu
def test_compressible_function_and_adjoint_solve_cuda():
    check_function_and_adjoint(None)


@pytest.mark.gpu
def test_compressible_transpose_product_parity_cuda():
    assert check_reverse(None, tol=1e-9) < 1e-9


def check_patch_functions(lib_path):
    """DAFunctionTotalPressure and DAFunctionMassFlowRate (patch reductions of p_b, U_b, rho_b) and their state derivatives,
    incompressible (rho = 1) and compressible, vs the oracle."""
    from tests.common import setup
    fn = {"TP": {"type": "totalPressure", "source": "patchToFace", "patches": ["inout"], "scale": 0.5},
          "MFR": {"type": "massFlowRate", "source": "patchToFace", "patches": ["inout"], "scale": 2.0}}
    one = np.array([1.0])
    for comp in (False, True):
        if comp:
            mesh, orc, sol, W = setup_comp(CONFIGS[0], lib_path)
            sol.updateDAOption(dict(normalizeStates=NS, normalizeResiduals=list(NRES), function=fn))
        else:
            mesh, bcs, orc, sol, W, _ = setup("naca", True, lib_path=lib_path, extra_options=dict(function=fn))
        sol.updateOFFields(W)
        ip = [p["name"] for p in mesh.patches].index("inout")
        tol = 1e-12 if lib_path is not None else 1e-9
        for name, mode, scale in (("TP", 2, 0.5), ("MFR", 3, 2.0)):
            F, Fo = sol.calcFunction(name), orc.force(W, ip, [1.0, 0.0, 0.0], scale, mode=mode)
            assert abs(Fo) > 0 and abs(F - Fo) <= tol * abs(Fo), (comp, name, F, Fo)
            g = np.zeros(orc.ndof)
            sol.calcJacTVecProduct("states", "stateVar", W, name, "function", one, g)
            go = orc.dforce_dw(W, ip, [1.0, 0.0, 0.0], scale, mode=mode)
            assert np.linalg.norm(go) > 0 and rel_err(g, go) < tol, (comp, name, rel_err(g, go))


def check_total_pressure_ratio(lib_path):
    """DAFunctionTotalPressureRatio (outlet / inlet area-averaged isentropic total pressure): value and state derivative vs the
    oracle; its mesh derivative, and the one of the area-averaged totalPressure, vs central differences through updateOFMesh
    (the area sums move with the points)."""
    cfg = CONFIGS[2]
    mesh, orc, sol, W = setup_comp(cfg, lib_path)
    gamma = 1.4
    fn = {"TPR": {"type": "totalPressureRatio", "source": "patchToFace", "patches": ["inlet", "outlet"], "inletPatches": ["inlet"],
                  "outletPatches": ["outlet"], "scale": 1.0},
          "TP": {"type": "totalPressure", "source": "patchToFace", "patches": ["outlet"], "scale": 0.5}}
    sol.updateDAOption(dict(normalizeStates=NS, normalizeResiduals=list(cfg[7]), function=fn))
    sol.updateOFFields(W)
    names = [p["name"] for p in mesh.patches]
    i_in, i_out = names.index("inlet"), names.index("outlet")
    g3 = [gamma, 0.0, 0.0]
    A, B = orc.force(W, i_out, g3, 1.0, mode=4), orc.force(W, i_in, g3, 1.0, mode=4)
    tol = 1e-12 if lib_path is not None else 1e-9
    F = sol.calcFunction("TPR")
    assert abs(F - A / B) <= tol * abs(A / B), (F, A / B)
    g = np.zeros(orc.ndof)
    sol.calcJacTVecProduct("states", "stateVar", W, "TPR", "function", np.array([1.0]), g)
    go = orc.dforce_dw(W, i_out, g3, 1.0, mode=4) / B - A / B**2 * orc.dforce_dw(W, i_in, g3, 1.0, mode=4)
    assert np.linalg.norm(go) > 0 and rel_err(g, go) < max(tol, 1e-11), rel_err(g, go)
    nP3 = 3 * sol.getNLocalPoints()
    pts = np.zeros(nP3)
    sol.getOFMeshPoints(pts)
    X = pts.reshape(-1, 3)
    v = np.stack([np.sin(3.0 * X[:, 1]) * 1e-3, np.cos(2.0 * X[:, 0] + X[:, 1]) * 1e-3, np.zeros(len(X))], axis=1).ravel()
    h = 1e-3
    for name in ("TPR", "TP"):
        dFdx = np.zeros(nP3)
        sol.calcJacTVecProduct("aero_vol_coords", "volCoord", pts, name, "function", np.array([1.0]), dFdx)
        sol.updateOFMesh(pts + h * v)
        Fp = sol.calcFunction(name)
        sol.updateOFMesh(pts - h * v)
        Fm = sol.calcFunction(name)
        sol.updateOFMesh(pts)
        fd = (Fp - Fm) / (2 * h)
        assert abs(fd) > 0 and abs(fd - dFdx @ v) <= 2e-5 * abs(fd), (name, fd, dFdx @ v)


def test_total_pressure_ratio_host_build():
    check_total_pressure_ratio(HOSTSIM)


@pytest.mark.gpu
def test_total_pressure_ratio_cuda():
    check_total_pressure_ratio(None)


def test_total_pressure_and_mass_flow_rate_host_build():
    check_patch_functions(HOSTSIM)


@pytest.mark.gpu
def test_total_pressure_and_mass_flow_rate_cuda():
    check_patch_functions(None)


FVSOURCE = {"disk1": {"type": "actuatorDisk", "source": "cylinderAnnulusSmooth", "center": [1.6, 0.1, 0.05], "direction": [1.0, 0.0, 0.0],
                      "innerRadius": 0.05, "outerRadius": 0.6, "rotDir": "right", "scale": 100.0, "POD": 0.8, "eps": 0.25, "expM": 1.0,
                      "expN": 0.5, "adjustThrust": 0, "targetThrust": 1.0}}


def check_fvsource(lib_path):
    """fvSource actuator disk (DAFvSourceActuatorDisk, cylinderAnnulusSmooth) in the momentum and energy rows, the fvSourcePar
    input and the mesh dependence of the source in the volCoord product -- incompressible and compressible, vs the oracle."""
    from tests.common import setup
    inp = {"actuator_disk": {"type": "fvSourcePar", "fvSourceName": "disk1", "indices": [0, 7, 8, 9]}}
    for comp in (False, True):
        if comp:
            mesh, orc, sol, W = setup_comp(("naca", "sensibleInternalEnergy", "const", "SpalartAllmaras", "linearUpwind", "upwind", False, NRES), lib_path)
            sol.updateDAOption(dict(normalizeStates=NS, normalizeResiduals=list(NRES), fvSource=FVSOURCE, inputInfo=inp))
        else:
            mesh, bcs, orc, sol, W, _ = setup("naca", True, lib_path=lib_path, extra_options=dict(fvSource=FVSOURCE, inputInfo=inp))
        base = orc.residual(W)
        orc.set_fvsource(FVSOURCE)
        sol.updateOFFields(W)
        R = np.zeros(orc.ndof)
        sol.getResiduals(R)
        Ro = orc.residual(W)
        assert rel_err(Ro, base) > 1e-6  # the disk sits inside the mesh
        tolp = 1e-11 if lib_path is not None else 1e-9
        assert rel_err(R, Ro) < tolp, (comp, rel_err(R, Ro))
        orc.record(W)
        psi = np.random.default_rng(9).uniform(-1, 1, orc.ndof)
        y = np.zeros(orc.ndof)
        sol.calcdRdWTPsiAD(psi, y)
        assert rel_err(y, orc.jtvec(psi)) < tolp
        # fvSourcePar: d(psi.R)/d(center_x, outerRadius, scale, POD) vs central differences of the oracle
        d = FVSOURCE["disk1"]
        x = np.array([d["center"][0], d["outerRadius"], d["scale"], d["POD"]])
        assert sol.getInputSize("actuator_disk", "fvSourcePar") == 4
        prod = np.zeros(4)
        sol.calcJacTVecProduct("actuator_disk", "fvSourcePar", x, "R", "residual", psi, prod)
        ref = np.zeros(4)
        import copy
        for k, key in enumerate(("center", "outerRadius", "scale", "POD")):
            h = 1e-5 * max(1.0, abs(x[k]))
            vals = []
            for sgn in (1.0, -1.0):
                fs = copy.deepcopy(FVSOURCE)
                if key == "center":
                    fs["disk1"]["center"][0] += sgn * h
                else:
                    fs["disk1"][key] += sgn * h
                orc.set_fvsource(fs)
                vals.append(psi @ orc.residual(W))
            ref[k] = (vals[0] - vals[1]) / (2 * h)
        orc.set_fvsource(FVSOURCE)
        # both sides are central differences of psi.R; R ~ 1e6 in the compressible case limits them to ~1e-4
        assert np.allclose(prod, ref, rtol=3e-4 if comp else 1e-5, atol=1e-7 * np.abs(ref).max()), (comp, prod, ref)
        # the source moves with the cell centres: volCoord product vs the oracle's tape through the geometry
        nP3 = 3 * sol.getNLocalPoints()
        pts = np.zeros(nP3)
        sol.getOFMeshPoints(pts)
        pv = np.zeros(nP3)
        sol.calcJacTVecProduct("aero_vol_coords", "volCoord", pts, "R", "residual", psi, pv)
        refv = orc.jtvec_xv(W, psi)
        mask = np.ones((nP3 // 3, 3), dtype=bool)
        for pch in mesh.patches:
            if pch["type"] == "symmetry":
                fp = mesh.faces[pch["start"]:pch["start"] + pch["size"]]
                mask[np.unique(fp[fp >= 0]), 2] = False
        mask = mask.ravel()
        assert rel_err(pv[mask], refv[mask]) < 1e-7


def test_fvsource_actuator_disk_host_build():
    check_fvsource(HOSTSIM)


@pytest.mark.gpu
def test_fvsource_actuator_disk_cuda():
    check_fvsource(None)
