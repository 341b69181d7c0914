"""MRF (multiple reference frame) zones on the DASimpleFoam / DARhoSimpleFoam / DATurboFoam residual (reference
src/adjoint/DAMisc/MRFDF/MRFZoneDF.C, DAResidualSimpleFoam.C:144,183,246, DAResidualRhoSimpleFoam.C:128,188,226,
DAResidualTurboFoam.C:108-110,193-195): Coriolis term, rotating-wall velocity, relative face fluxes."""
import numpy as np
import pytest

from dafoam_b200 import cases
from oracle.pyoracle import Oracle, synthetic_state
from tests.common import HOSTSIM, mrf_zone


def mrf_spec(mesh, C=None, wall="wing", omega=40.0):
    return mrf_zone(mesh, wall=wall, omega=omega)


def test_oracle_mrf_changes_the_residual_and_tape_matches_fd():
    mesh = cases.naca0012_ogrid(ni=24, nj=12, nk=2)
    for comp in (False, True):
        if comp:
            th = cases.default_thermo()
            bcs = cases.compressible_bcs(cases.default_bcs_naca(U0=(50.0, 2.0, 0.0)))
            ns = dict(U=50.0, p=101325.0, T=300.0, nuTilda=1e-3, phi=1.0)
            orc = Oracle(mesh, bcs, normalizeStates=ns, normalizeResiduals=("URes", "pRes", "TRes", "nuTildaRes", "phiRes"), thermo=th)
            W = synthetic_state(mesh, orc.geometry("C"), orc.geometry("Sf"), U0=(50.0, 2.0, 0.0), thermo=th)
        else:
            bcs = cases.default_bcs_naca()
            orc = Oracle(mesh, bcs, normalizeStates=dict(U=10.0, p=50.0, nuTilda=1e-3, phi=1.0))
            W = synthetic_state(mesh, orc.geometry("C"), orc.geometry("Sf"))
        R0 = orc.residual(W)
        mrf = mrf_spec(mesh, orc.geometry("C").reshape(-1, 3))
        assert 0 < len(mrf["cells"]) < mesh.n_cells
        orc.set_mrf(mesh, mrf)
        R = orc.residual(W)
        nC = mesh.n_cells
        assert np.linalg.norm((R - R0)[:3 * nC]) > 1e-3 * np.linalg.norm(R0[:3 * nC])  # Coriolis + rotating wall
        assert np.linalg.norm((R - R0)[-mesh.n_faces:]) > 0                            # relative fluxes
        orc.record(W)
        rng = np.random.default_rng(0)
        psi = rng.uniform(-1, 1, orc.ndof)
        g = orc.jtvec(psi, normalize=False)
        v = rng.uniform(-1, 1, orc.ndof) * np.abs(W) * 1e-1 + 1e-12
        eps = 1e-6
        fd = psi @ (orc.residual(W + eps * v) - orc.residual(W - eps * v)) / (2 * eps)
        assert abs(g @ v - fd) <= 1e-7 * abs(fd), (comp, g @ v, fd)


NS_C = dict(U=50.0, p=101325.0, T=300.0, nuTilda=1e-3, phi=1.0)
NRES_C = ("URes", "pRes", "TRes", "nuTildaRes", "phiRes")
NS_I = dict(U=10.0, p=50.0, nuTilda=1e-3, phi=1.0)


def setup_mrf(solver, lib_path, function=None, omega=40.0, energy="sensibleInternalEnergy"):
    import tempfile
    from dafoam_b200.pyDASolvers import pyDASolvers
    mesh = cases.naca0012_ogrid(ni=24, nj=12, nk=2)
    comp = solver != "DASimpleFoam"
    if comp:
        th = cases.default_thermo(energy=energy)
        bcs = cases.compressible_bcs(cases.default_bcs_naca(U0=(50.0, 2.0, 0.0)))
        orc = Oracle(mesh, bcs, normalizeStates=NS_C, normalizeResiduals=NRES_C, thermo=th, divU="linearUpwindV")
        orc.set_turbo(solver == "DATurboFoam")
        W = synthetic_state(mesh, orc.geometry("C"), orc.geometry("Sf"), U0=(50.0, 2.0, 0.0), thermo=th)
        opts = dict(normalizeStates=NS_C, normalizeResiduals=list(NRES_C))
        kw = dict(thermo=th)
    else:
        bcs = cases.default_bcs_naca()
        orc = Oracle(mesh, bcs, normalizeStates=NS_I, divU="linearUpwindV")
        W = synthetic_state(mesh, orc.geometry("C"), orc.geometry("Sf"))
        opts = dict(normalizeStates=NS_I)
        kw = {}
    if function:
        opts["function"] = function
    mrf = mrf_spec(mesh, orc.geometry("C").reshape(-1, 3), omega=omega)
    orc.set_mrf(mesh, mrf)
    d = tempfile.mkdtemp(prefix="dab_mrf_")
    cases.write_case(d, mesh, bcs, div_u="bounded Gauss linearUpwindV grad(U)", mrf=mrf, **kw)
    sol = pyDASolvers("%s -python" % solver, opts, caseDir=d, _lib_path=lib_path)
    return mesh, orc, sol, W


def check_engine(lib_path, tol=1e-10):
    from tests.common import rel_err
    for solver, energy in (("DASimpleFoam", None), ("DARhoSimpleFoam", "sensibleInternalEnergy"), ("DATurboFoam", "sensibleInternalEnergy"),
                           ("DATurboFoam", "sensibleEnthalpy")):
        mesh, orc, sol, W = setup_mrf(solver, lib_path, energy=energy)
        n = orc.ndof
        assert sol.getNLocalAdjointStates() == n
        sol.updateOFFields(W)
        nC = mesh.n_cells
        if energy == "sensibleEnthalpy":
            # the viscous-work and p(U - URel) terms of DATurboFoam's enthalpy equation are really there
            Rt = orc.residual(W)
            orc.set_turbo(False)
            assert np.linalg.norm((Rt - orc.residual(W))[4 * nC:5 * nC]) > 1e-3 * np.linalg.norm(Rt[4 * nC:5 * nC])
            orc.set_turbo(True)
        ns = (n - mesh.n_faces) // nC
        segs = [("U", 0, 3 * nC)] + [("s%d" % k, k * nC, (k + 1) * nC) for k in range(3, ns)] + [("phi", ns * nC, n)]
        for isPC in (0, 1):
            R = np.zeros(n)
            sol.getResiduals(R, isPC)
            Ro = orc.residual(W, isPC)
            for name, a, b in segs:
                assert rel_err(R[a:b], Ro[a:b]) < tol, (solver, isPC, name, rel_err(R[a:b], Ro[a:b]))
        orc.record(W)
        rng = np.random.default_rng(11)
        psi = rng.uniform(-1, 1, n)
        y = np.zeros(n)
        sol.calcdRdWTPsiAD(psi, y)
        yo = orc.jtvec(psi)
        for name, a, b in segs:
            assert rel_err(y[a:b], yo[a:b]) < tol, (solver, "JT", name, rel_err(y[a:b], yo[a:b]))


def test_mrf_residual_and_transpose_product_host_build():
    check_engine(HOSTSIM)


@pytest.mark.gpu
def test_mrf_residual_and_transpose_product_cuda():
    check_engine(None, tol=1e-9)


def check_function_volcoord_primal(lib_path, tol=1e-10):
    from tests.common import rel_err
    dirv = [0.8, 0.6, 0.0]
    fn = {"CD": {"type": "force", "source": "patchToFace", "patches": ["wing"], "directionMode": "fixedDirection", "direction": dirv,
                 "scale": 0.02}}
    # force on the rotating wall (its velocity enters the wall shear stress) and its state derivative
    for solver in ("DASimpleFoam", "DATurboFoam"):
        mesh, orc, sol, W = setup_mrf(solver, lib_path, function=fn)
        sol.updateOFFields(W)
        wing = [p["name"] for p in mesh.patches].index("wing")
        F, Fo = sol.calcFunction("CD"), orc.force(W, wing, dirv, 0.02)
        assert abs(F - Fo) <= max(1e-11, 0.1 * tol) * abs(Fo), (solver, F, Fo)
        n = orc.ndof
        g = np.zeros(n)
        sol.calcJacTVecProduct("states", "stateVar", W, "CD", "function", np.array([1.0]), g)
        assert rel_err(g, orc.dforce_dw(W, wing, dirv, 0.02)) < tol, solver
    # [dR/dx_v]^T psi: the relative fluxes and the wall velocity follow the face centres and areas
    mesh, orc, sol, W = setup_mrf("DASimpleFoam", lib_path)
    sol.updateOFFields(W)
    nP3 = 3 * sol.getNLocalPoints()
    pts = np.zeros(nP3)
    sol.getOFMeshPoints(pts)
    psi = np.random.default_rng(5).uniform(-1, 1, orc.ndof)
    prod = np.zeros(nP3)
    sol.calcJacTVecProduct("aero_vol_coords", "volCoord", pts, "R", "residual", psi, prod)
    ref = orc.jtvec_xv(W, psi)
    mask = np.ones((nP3 // 3, 3), dtype=bool)
    for pch in mesh.patches:
        if pch["type"] == "symmetry":
            fp = mesh.faces[pch["start"]:pch["start"] + pch["size"]]
            mask[np.unique(fp[fp >= 0]), 2] = False  # |n_k| kink of the symmetry transform (see test_volcoord.py)
    mask = mask.ravel()
    assert rel_err(prod[mask], ref[mask]) < 1e-7, rel_err(prod[mask], ref[mask])
    # SIMPLE with the Coriolis source, the rotating wall and the relative fluxes: its fixed point is the root of the same R(W)
    for solver, its, energy in (("DASimpleFoam", 1000, None), ("DATurboFoam", 4000, "sensibleInternalEnergy"),
                                ("DATurboFoam", 4000, "sensibleEnthalpy")):
        mesh, orc, sol, W = setup_mrf(solver, lib_path, omega=0.3, energy=energy)
        sol.updateDAOption(dict(primalMinResTol=1e-9, primalMaxIters=its))
        n = orc.ndof
        W0 = np.zeros(n)
        sol.getOFFields(W0)
        assert sol.solvePrimal() == 0, (solver, sol.primalStats.max_residual)
        W1 = np.zeros(n)
        sol.getOFFields(W1)
        r0, r1 = np.linalg.norm(orc.residual(W0)), np.linalg.norm(orc.residual(W1))
        assert r1 < 1e-6 * r0, (solver, r0, r1)


def test_mrf_function_volcoord_and_primal_host_build():
    check_function_volcoord_primal(HOSTSIM)


@pytest.mark.gpu
def test_mrf_function_volcoord_and_primal_cuda():
    check_function_volcoord_primal(None, tol=1e-9)


def test_primal_bc_option_host_build():
    """primalBC (reference DAField::setPrimalBoundaryConditions): boundary values, MRF speed, laminar viscosity and the nut wall
    treatment from the options give the residual of a case written with those values."""
    import tempfile
    from dafoam_b200.pyDASolvers import pyDASolvers
    from tests.common import rel_err
    mesh = cases.naca0012_ogrid(ni=24, nj=12, nk=2)
    th = cases.default_thermo(mu=2.5e-5)
    U1 = (60.0, 1.0, 0.0)
    # the target: written BCs U1, T 310, wall function, omega 20, mu 2.5e-5
    bcs1 = cases.compressible_bcs(cases.default_bcs_naca(U0=U1, wall_function=True), T0=310.0)
    orc = Oracle(mesh, bcs1, normalizeStates=NS_C, normalizeResiduals=NRES_C, thermo=th, divU="linearUpwindV")
    orc.set_mrf(mesh, mrf_zone(mesh, omega=20.0))
    W = synthetic_state(mesh, orc.geometry("C"), orc.geometry("Sf"), U0=(50.0, 2.0, 0.0), thermo=th)
    # the case on disk: other values everywhere, corrected through primalBC
    bcs0 = cases.compressible_bcs(cases.default_bcs_naca(U0=(50.0, 2.0, 0.0)))
    d = tempfile.mkdtemp(prefix="dab_pbc_")
    cases.write_case(d, mesh, bcs0, div_u="bounded Gauss linearUpwindV grad(U)", mrf=mrf_zone(mesh, omega=5.0), thermo=cases.default_thermo())
    pbc = {"U0": {"variable": "U", "patches": ["inout"], "value": list(U1)},
           "T0": {"variable": "T", "patches": ["inout"], "value": [310.0]},
           "k0": {"variable": "k", "patches": ["inout"], "value": [0.1]},  # not a field of this solver: skipped
           "useWallFunction": True, "MRF": 20.0, "thermo:mu": 2.5e-5}
    sol = pyDASolvers("DATurboFoam -python", dict(normalizeStates=NS_C, normalizeResiduals=list(NRES_C), primalBC=pbc), caseDir=d,
                      _lib_path=HOSTSIM)
    sol.updateOFFields(W)
    R = np.zeros(orc.ndof)
    sol.getResiduals(R)
    assert rel_err(R, orc.residual(W)) < 1e-10
    # ... and again after a later option update
    sol.updateDAOption(dict(primalBC={"MRF": 5.0}))
    R2 = np.zeros(orc.ndof)
    sol.getResiduals(R2)
    orc.set_mrf(mesh, mrf_zone(mesh, omega=5.0))
    assert rel_err(R2, orc.residual(W)) < 1e-10
