"""Cyclic (periodic) patch pairs: one passage of an annular duct with `cyclic` sides against the closed ring of n passages.

The reference couples the two patches of a pair through cyclicFvPatchField (neighbour values rotated by the patch transform inside
every fvc:: operator) and counts the coupled faces as states (reference src/adjoint/DAIndex/DAIndex.C:151-167); BASELINE config 5
(DATurboFoam rotor passage) needs them.  Known answer: the periodic problem written out -- the same passage repeated n times around
the axis, closed into a ring, NO cyclic patch anywhere, evaluated by the oracle.  With a state that repeats from passage to passage
(vectors rotated by the passage angle), the residual rows of passage 0 and the rows of [dR/dW]^T psi of passage 0 must be the ones
the engine computes on the single passage with cyclic sides.
"""
import tempfile

import numpy as np
import pytest

from dafoam_b200 import cases
from dafoam_b200.pyDASolvers import pyDASolvers
from oracle.pyoracle import Oracle
from tests.common import HOSTSIM, NORM_STATES, ALL_RES, rel_err

N_SECTORS = 5


def merged_faces(mesh):
    """Face numbering of the engine on a mesh with cyclic patches (mesh.hpp mergeCyclics): internal faces, one face per coupled
    pair (owner = the cell on the first patch of the pair), the remaining boundary faces patch by patch.
    Returns owner, neighbour (-1: boundary), patch name per face ('' for internal)."""
    nIF = mesh.n_internal_faces
    own, nei, pname = list(mesh.owner[:nIF]), list(mesh.neighbour), [""] * nIF
    names = [p["name"] for p in mesh.patches]
    for pi, p in enumerate(mesh.patches):
        if p["type"] != "cyclic":
            continue
        qi = names.index(p["neighbourPatch"])
        if qi < pi:
            continue
        q = mesh.patches[qi]
        for i in range(p["size"]):
            own.append(mesh.owner[p["start"] + i])
            nei.append(mesh.owner[q["start"] + i])
            pname.append("")
    for p in mesh.patches:
        if p["type"] == "cyclic":
            continue
        for i in range(p["size"]):
            own.append(mesh.owner[p["start"] + i])
            nei.append(-1)
            pname.append(p["name"])
    return np.array(own), np.array(nei), pname


def rotz(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


class Pair:
    """One passage with cyclic sides (engine) and the ring of N_SECTORS passages (oracle), with the maps between them."""

    def __init__(self, turbulent=True, divU="linearUpwind", lib_path=HOSTSIM, dims=(4, 4, 6), mrf_omega=None, extra=None, solver="DASimpleFoam",
                 energy="sensibleEnthalpy"):
        nr, nt, nz = dims
        self.turb = turbulent
        self.comp = solver != "DASimpleFoam"
        self.ns = 4 + int(turbulent) + int(self.comp)
        self.Uax = 60.0 if self.comp else 10.0
        self.sec = cases.annular_passage(nr=nr, nt=nt, nz=nz, n_sectors=N_SECTORS, sectors=1)
        self.full = cases.annular_passage(nr=nr, nt=nt, nz=nz, n_sectors=N_SECTORS, sectors=N_SECTORS)
        self.nCs = self.sec.n_cells
        bs = cases.default_bcs_passage(Uin=(0.0, 0.0, self.Uax), turbulent=turbulent, cyclic=True)
        bf = cases.default_bcs_passage(Uin=(0.0, 0.0, self.Uax), turbulent=turbulent, cyclic=False)
        self.thermo = None
        ns_, nres_ = NORM_STATES, ALL_RES
        if self.comp:
            bs, bf = cases.compressible_bcs(bs), cases.compressible_bcs(bf)
            self.thermo = cases.default_thermo(energy=energy)
            ns_ = dict(U=50.0, p=101325.0, T=300.0, nuTilda=1e-3, phi=1.0)
            nres_ = ("URes", "pRes", "TRes", "nuTildaRes", "phiRes")
        div_u = "bounded Gauss %s%s" % (divU, " grad(U)" if divU.startswith("linearUpwind") else "")
        d = tempfile.mkdtemp(prefix="dab_cyc_")
        kw = {}
        self.mrf_s = self.mrf_f = None
        if mrf_omega is not None:
            self.mrf_s = dict(cellZone="rotor", cells=np.arange(self.nCs), origin=(0.0, 0.0, 0.0), axis=(0.0, 0.0, 1.0), omega=mrf_omega,
                              nonRotatingPatches=["inlet", "outlet", "shroud"])
            self.mrf_f = dict(self.mrf_s, cells=np.arange(self.full.n_cells))
            kw["mrf"] = self.mrf_s
        if self.thermo is not None:
            kw["thermo"] = self.thermo
        cases.write_case(d, self.sec, bs, div_u=div_u, **kw)
        opts = dict(normalizeStates=ns_, normalizeResiduals=list(nres_))
        opts.update(extra or {})
        self.sol = pyDASolvers(solver + " -python", opts, caseDir=d, _lib_path=lib_path)
        okw = dict(thermo=self.thermo) if self.comp else {}
        self.orc = Oracle(self.full, bf, normalizeStates=ns_, divU=divU, normalizeResiduals=nres_, **okw)
        if self.comp:
            self.orc.set_turbo(solver == "DATurboFoam")
        if self.mrf_f is not None:
            self.orc.set_mrf(self.full, self.mrf_f)
        self.case_dir = d
        # faces of the ring -> merged faces of the passage, by the pair of passage-local cells (boundary: cell + patch)
        so, sn, sp = merged_faces(self.sec)
        self.nFs = so.size
        key = {}
        for g in range(self.nFs):
            k = (min(so[g], sn[g]), max(so[g], sn[g])) if sn[g] >= 0 else (so[g], sp[g])
            assert k not in key
            key[k] = g
        nF, nIF = self.full.n_faces, self.full.n_internal_faces
        pn = [""] * nF
        for p in self.full.patches:
            for i in range(p["size"]):
                pn[p["start"] + i] = p["name"]
        self.f2s = np.zeros(nF, dtype=np.int64)
        self.fsign = np.ones(nF)
        self.s2f = -np.ones(self.nFs, dtype=np.int64)  # the ring face that is the passage-0 instance of a merged face
        for f in range(nF):
            o = self.full.owner[f]
            if f < nIF:
                n = self.full.neighbour[f]
                ol, nl = o % self.nCs, n % self.nCs
                g = key[(min(ol, nl), max(ol, nl))]
                same = ol == so[g]
                self.fsign[f] = 1.0 if same else -1.0
                owner_cell = o if same else n  # the ring cell playing the owner of the merged face
            else:
                g = key[(o % self.nCs, pn[f])]
                owner_cell = o
            self.f2s[f] = g
            if owner_cell < self.nCs:
                self.s2f[g] = f
        assert (self.s2f >= 0).all()
        # local (engine) layout -> merged numbering
        self.idx = self.sol.localStateIndex(self.nCs, self.nFs, turbulent=turbulent, compressible=self.comp)
        self.owned = np.concatenate([np.ones(self.ns * self.sol.getNLocalCells(), dtype=bool),
                                     self.sol.getLocalToGlobal("faceOwned").astype(bool)])
        assert self.sol.getNLocalCells() == self.nCs

    # ---- vectors in the merged (passage) numbering <-> ring numbering
    def n_sec(self):
        return self.ns * self.nCs + self.nFs

    def to_ring(self, v):
        nCs, nCf = self.nCs, self.full.n_cells
        ns = self.ns
        out = np.zeros(ns * nCf + self.full.n_faces)
        U = v[:3 * nCs].reshape(nCs, 3)
        for s in range(N_SECTORS):
            out[3 * s * nCs:3 * (s + 1) * nCs] = (U @ rotz(s * self.sec.sector_angle).T).ravel()
            for k in range(ns - 3):
                out[(3 + k) * nCf + s * nCs:(3 + k) * nCf + (s + 1) * nCs] = v[(3 + k) * nCs:(4 + k) * nCs]
        out[ns * nCf:] = self.fsign * v[ns * nCs + self.f2s]
        return out

    def from_ring(self, w):
        """rows of passage 0"""
        nCs, nCf = self.nCs, self.full.n_cells
        ns = self.ns
        out = np.zeros(self.n_sec())
        out[:3 * nCs] = w[:3 * nCs]
        for k in range(ns - 3):
            out[(3 + k) * nCs:(4 + k) * nCs] = w[(3 + k) * nCf:(3 + k) * nCf + nCs]
        out[ns * nCs:] = self.fsign[self.s2f] * w[ns * nCf + self.s2f]
        return out

    def state(self, seed=7):
        """a flow through the passage: axial velocity with swirl and a boundary-layer-like profile, passage-periodic"""
        rng = np.random.default_rng(seed)
        nCs = self.nCs
        C = np.asarray(self.orc.geometry("C")).reshape(-1, 3)[:nCs]
        r = np.hypot(C[:, 0], C[:, 1])
        th = np.arctan2(C[:, 1], C[:, 0])
        z = C[:, 2]
        eta = (r - 0.2) / 0.15
        prof = 4.0 * eta * (1.0 - eta) + 0.2
        ur = 0.4 * np.sin(N_SECTORS * th) * np.sin(np.pi * eta)
        ut = 0.3 * self.Uax * prof * (1.0 + 0.2 * np.cos(N_SECTORS * th)) + 0.1 * self.Uax * z
        uz = self.Uax * prof * (1.0 + 0.1 * np.sin(N_SECTORS * th + 3.0 * z))
        U = np.stack([ur * np.cos(th) - ut * np.sin(th), ur * np.sin(th) + ut * np.cos(th), uz], axis=1)
        U *= 1.0 + 0.01 * rng.uniform(-1, 1, U.shape)
        p = 20.0 * (1.0 - z / 0.3) + 5.0 * np.cos(N_SECTORS * th) * eta + 0.2 * rng.uniform(-1, 1, nCs)
        rho = np.ones(nCs)
        if self.comp:
            p = 101325.0 + 40.0 * p
            T = 300.0 * (1.0 + 0.02 * np.sin(N_SECTORS * th) * np.cos(9.0 * z)) * (1.0 + 0.001 * rng.uniform(-1, 1, nCs))
            rho = p / (8314.4700665 / self.thermo["molWeight"] * T)
        parts = [U.ravel(), p]
        if self.comp:
            parts.append(T)
        if self.turb:
            parts.append(4.5e-5 * (1.0 + 3.0 * prof) * (1.0 + 0.01 * rng.uniform(-1, 1, nCs)))
        so, sn, _ = merged_faces(self.sec)
        # flux through the merged faces from the ring geometry of their passage-0 instances
        S = np.asarray(self.orc.geometry("Sf")).reshape(-1, 3)[self.s2f] * self.fsign[self.s2f][:, None]
        ring_o = np.where(self.fsign[self.s2f] > 0, self.full.owner[self.s2f], 0)
        Uo = U[so]
        # the neighbour's velocity in the owner's frame: across the coupled pair it is the rotated image
        Un = np.where((sn >= 0)[:, None], U[np.maximum(sn, 0)], Uo)
        nIF = self.sec.n_internal_faces
        nCyc = sum(p["size"] for p in self.sec.patches if p["type"] == "cyclic") // 2
        Rm = rotz(-self.sec.sector_angle)  # per_hi cells seen from per_lo: one passage back
        Un[nIF:nIF + nCyc] = Un[nIF:nIF + nCyc] @ Rm.T
        del ring_o
        rf = 0.5 * (rho[so] + np.where(sn >= 0, rho[np.maximum(sn, 0)], rho[so]))
        phi = rf * np.einsum("ij,ij->i", 0.5 * (Uo + Un), S) * (1.0 + 0.01 * rng.uniform(-1, 1, self.nFs))
        _, _, pname = merged_faces(self.sec)
        for g in range(self.nFs):
            if pname[g] in ("hub", "shroud"):
                phi[g] = 0.0
        parts.append(phi)
        return np.concatenate(parts)

    def local(self, v):
        return np.ascontiguousarray(v[self.idx])

    def merged(self, vloc):
        out = np.zeros(self.n_sec())
        out[self.idx[self.owned]] = vloc[self.owned]
        return out

    def segments(self):
        nC = self.nCs
        segs = [("U", 0, 3 * nC), ("p", 3 * nC, 4 * nC)]
        k = 4
        if self.comp:
            segs.append(("T", k * nC, (k + 1) * nC))
            k += 1
        if self.turb:
            segs.append(("nuTilda", k * nC, (k + 1) * nC))
            k += 1
        segs.append(("phi", k * nC, self.n_sec()))
        return segs


def check_pair(P, tol=1e-9):
    W = P.state()
    Wr = P.to_ring(W)
    # the ring state repeats: its passage-0 rows are the passage state
    assert np.array_equal(P.from_ring(Wr), W)
    P.sol.updateOFFields(P.local(W))
    worst = 0.0
    for isPC in (0, 1):
        R = np.zeros(P.idx.size)
        P.sol.getResiduals(R, isPC)
        Rs, Ro = P.merged(R), P.from_ring(P.orc.residual(Wr, isPC))
        for name, a, b in P.segments():
            e = rel_err(Rs[a:b], Ro[a:b])
            worst = max(worst, e)
            assert e < tol, ("residual", isPC, name, e)
    P.orc.record(Wr)
    rng = np.random.default_rng(99)
    for trial in range(2):
        psi = rng.uniform(-1, 1, P.n_sec()) if trial == 0 else np.full(P.n_sec(), 1e-3)
        if trial == 1:
            psi[:3 * P.nCs] = (np.array([0.3e-3, -0.7e-3, 1e-3])[None, :] * np.ones((P.nCs, 1))).ravel()
        y = np.zeros(P.idx.size)
        P.sol.calcdRdWTPsiAD(P.local(psi), y)
        ys, yo = P.merged(y), P.from_ring(P.orc.jtvec(P.to_ring(psi)))
        for name, a, b in P.segments():
            e = rel_err(ys[a:b], yo[a:b])
            worst = max(worst, e)
            assert e < tol, ("jtvec", trial, name, e)
    return worst


def test_passage_generator_is_periodic():
    sec = cases.annular_passage(n_sectors=N_SECTORS)
    lo = next(p for p in sec.patches if p["name"] == "per_lo")
    hi = next(p for p in sec.patches if p["name"] == "per_hi")
    assert lo["size"] == hi["size"] > 0
    R = rotz(sec.sector_angle)
    for i in range(lo["size"]):
        a = sec.points[sec.faces[lo["start"] + i]].mean(axis=0)
        b = sec.points[sec.faces[hi["start"] + i]].mean(axis=0)
        assert np.allclose(R @ a, b, atol=1e-13)


@pytest.mark.parametrize("turbulent,divU", [(True, "linearUpwind"), (False, "upwind"), (True, "linearUpwindV")])
def test_cyclic_passage_equals_ring_host_build(turbulent, divU):
    worst = check_pair(Pair(turbulent, divU))
    assert worst < 1e-9


def test_cyclic_passage_with_mrf_host_build():
    worst = check_pair(Pair(True, "linearUpwind", mrf_omega=30.0))
    assert worst < 1e-9


@pytest.mark.parametrize("solver,energy,omega", [("DARhoSimpleFoam", "sensibleInternalEnergy", None), ("DATurboFoam", "sensibleEnthalpy", 300.0),
                                                 ("DATurboFoam", "sensibleInternalEnergy", 300.0)])
def test_cyclic_compressible_passage_host_build(solver, energy, omega):
    """BASELINE config 5's combination: DATurboFoam + SA + MRF rotor zone + cyclic sides"""
    worst = check_pair(Pair(True, "linearUpwindV", solver=solver, energy=energy, mrf_omega=omega))
    assert worst < 1e-9


def _channel_pair(lib_path, nz=3):
    """the convergent channel once with its two z-planes as a translational cyclic pair, once as symmetry planes"""
    out = []
    for cyc in (True, False):
        mesh = cases.channel(nx=12, ny=8, nz=nz)
        bcs = cases.default_bcs_channel()
        if cyc:
            for p in mesh.patches:
                if p["name"] in ("sym1", "sym2"):
                    p.update(type="cyclic", neighbourPatch="sym2" if p["name"] == "sym1" else "sym1", transform="translational",
                             separationVector=(0.0, 0.0, 0.1 if p["name"] == "sym1" else -0.1))
            for f in bcs.values():
                f[3]["sym1"] = dict(type="cyclic")
                f[3]["sym2"] = dict(type="cyclic")
        d = tempfile.mkdtemp(prefix="dab_cyc_")
        cases.write_case(d, mesh, bcs)
        sol = pyDASolvers("DASimpleFoam -python", dict(normalizeStates=NORM_STATES, normalizeResiduals=list(ALL_RES)), caseDir=d, _lib_path=lib_path)
        out.append((mesh, sol))
    return out


def test_translational_cyclic_host_build():
    """A flow that does not vary along z and has no z-velocity sees a translational cyclic pair exactly like two symmetry planes
    (momentum and nuTilda rows of the residual); and the transposed product of the cyclic mesh passes the dot-product test against central
    differences of its own residual."""
    (mc, sc), (ms, ss) = _channel_pair(HOSTSIM)
    nC, nz = mc.n_cells, 3
    nxy = nC // nz
    rng = np.random.default_rng(3)
    # z-invariant state on the symmetric mesh (its state vector is the polyMesh one)
    Ws = np.zeros(ss.getNLocalAdjointStates())
    ss.getOFFields(Ws)
    U = Ws[:3 * nC].reshape(nC, 3).copy()
    col = np.arange(nC) % nxy  # cells are numbered i + nx*(j + ny*k)
    U2 = (np.array([10.0, 0.4, 0.0])[None, :] * (1.0 + 0.05 * rng.uniform(-1, 1, (nxy, 3))))[col]
    U2[:, 2] = 0.0
    p2 = (3.0 * rng.uniform(-1, 1, nxy))[col]
    nt2 = (4.5e-5 * (1.0 + 0.3 * rng.uniform(0, 1, nxy)))[col]
    Sf, Cf = cases.quad_face_geometry(ms)
    nIF = ms.n_internal_faces
    nei = np.concatenate([ms.neighbour, ms.owner[nIF:]])
    phi = np.einsum("ij,ij->i", 0.5 * (U2[ms.owner] + U2[nei]), Sf)
    for pch in ms.patches:
        if pch["type"] in ("wall", "symmetry"):
            phi[pch["start"]:pch["start"] + pch["size"]] = 0.0
    Wsym = np.concatenate([U2.ravel(), p2, nt2, phi])
    ss.updateOFFields(Wsym)
    Rs = np.zeros(Wsym.size)
    ss.getResiduals(Rs)
    # the same state on the cyclic mesh: merged numbering = polyMesh faces with the pair merged (phi through the pair is zero: U_z = 0)
    order = cases.merged_face_order(mc)
    Wm = np.concatenate([U2.ravel(), p2, nt2, phi[order]])
    idx = sc.localStateIndex(nC, order.size)
    owned = np.concatenate([np.ones(5 * nC, dtype=bool), sc.getLocalToGlobal("faceOwned").astype(bool)])
    sc.updateOFFields(np.ascontiguousarray(Wm[idx]))
    Rc = np.zeros(idx.size)
    sc.getResiduals(Rc)
    Rm = np.zeros(Wm.size)
    Rm[idx[owned]] = Rc[owned]
    # (not the pressure rows: rAU carries the component-averaged boundary diagonal of the momentum matrix, to which a symmetry plane
    # contributes through its normal component and a cyclic pair does not)
    for a, b in ((0, 3 * nC), (4 * nC, 5 * nC)):
        assert rel_err(Rm[a:b], Rs[a:b]) < 1e-11, (a, b)
    # dot-product test of the cyclic operator: psi^T (R(W + h v) - R(W - h v)) / 2h = v^T (J^T psi)
    psi = rng.uniform(-1, 1, idx.size) * owned
    v = rng.uniform(-1, 1, idx.size) * owned
    v[5 * nC:] = 0.0  # cell states (every row of psi, the phi rows included, still takes part)
    y = np.zeros(idx.size)
    W0 = np.ascontiguousarray(Wm[idx])
    sc.calcdRdWTPsiAD(psi, y)
    # the product is y_j = s_j (J^T psi)_j with s_j the state scaling (normalizeStates; phi rows also carry the face area): perturb by s_j v_j
    fa = np.linalg.norm(Sf, axis=1)[order][sc.getLocalToGlobal("faces")]
    scale = np.concatenate([np.full(3 * nC, NORM_STATES["U"]), np.full(nC, NORM_STATES["p"]), np.full(nC, NORM_STATES["nuTilda"]),
                            NORM_STATES["phi"] * fa])
    h = 1e-5
    Rp, Rn = np.zeros(idx.size), np.zeros(idx.size)
    sc.updateOFFields(W0 + h * v * scale)
    sc.getResiduals(Rp)
    sc.updateOFFields(W0 - h * v * scale)
    sc.getResiduals(Rn)
    lhs = float(psi @ (Rp - Rn)) / (2.0 * h)
    rhs = float(v @ y)
    assert abs(lhs - rhs) < 1e-6 * max(abs(lhs), abs(rhs)), (lhs, rhs)


def solve_on_passage(lib_path, solver="DATurboFoam", dims=(6, 6, 12)):
    """the bench's config-5 workload in small: state from cases.passage_state, dRdWTPC + GMRES, the solution checked with the product"""
    from dafoam_b200.pyDASolvers import KSP, Mat
    comp = solver != "DASimpleFoam"
    mesh = cases.annular_passage(nr=dims[0], nt=dims[1], nz=dims[2], n_sectors=36)
    Uax = 100.0 if comp else 10.0
    bcs = cases.default_bcs_passage(Uin=(0.0, 0.0, Uax))
    th = cases.default_thermo(energy="sensibleEnthalpy") if comp else None
    mrf = dict(cellZone="rotor", cells=np.arange(mesh.n_cells), origin=(0.0, 0.0, 0.0), axis=(0.0, 0.0, 1.0), omega=300.0 if comp else 30.0,
               nonRotatingPatches=["inlet", "outlet", "shroud"])
    d = tempfile.mkdtemp(prefix="dab_cyc_")
    kw = dict(thermo=th) if comp else {}
    cases.write_case(d, mesh, cases.compressible_bcs(bcs) if comp else bcs, mrf=mrf, **kw)
    fn = {"FZ": {"type": "force", "source": "patchToFace", "patches": ["hub"], "directionMode": "fixedDirection", "direction": [0.0, 0.0, 1.0],
                 "scale": 1.0}}
    ns = dict(U=100.0, p=101325.0, T=300.0, nuTilda=1e-3, phi=1.0) if comp else NORM_STATES
    sol = pyDASolvers(solver + " -python", dict(normalizeStates=ns, function=fn,
                                                adjEqnOption=dict(gmresRelTol=1e-8, gmresMaxIters=1500, gmresRestart=1500, pcConLevel=2)),
                      caseDir=d, _lib_path=lib_path)
    Wg = cases.passage_state(mesh, Uax=Uax, thermo=th, n_sectors=36)
    nFg = cases.merged_face_order(mesh).size
    idx = sol.localStateIndex(mesh.n_cells, nFg, compressible=comp)
    owned = np.concatenate([np.ones((6 if comp else 5) * mesh.n_cells, dtype=bool), sol.getLocalToGlobal("faceOwned").astype(bool)])
    assert idx.size == sol.getNLocalAdjointStates() and int(owned.sum()) == Wg.size
    W = np.ascontiguousarray(Wg[idx])
    sol.updateOFFields(W)
    R = np.zeros(idx.size)
    sol.getResiduals(R)
    assert np.all(np.isfinite(R)) and np.all(R[~owned] == 0.0)
    b = np.zeros(idx.size)
    sol.calcJacTVecProduct("s", "stateVar", W, "FZ", "function", np.array([1.0]), b)
    assert np.linalg.norm(b) > 0
    pc, ksp = Mat(), KSP()
    sol.calcdRdWT(1, pc)
    x = np.zeros(idx.size)
    fail = sol.solveLinearEqn(ksp, b, x)
    assert fail == 0, (fail, ksp.stats.iterations, ksp.stats.final_residual / ksp.stats.initial_residual)
    y = np.zeros(idx.size)
    sol.calcdRdWTPsiAD(x, y)
    res = np.linalg.norm((y - b)[owned]) / np.linalg.norm(b[owned])
    assert res < 1e-6, res
    return ksp.stats.iterations, res


@pytest.mark.parametrize("solver", ["DASimpleFoam", "DATurboFoam"])
def test_adjoint_solve_on_passage_host_build(solver):
    its, res = solve_on_passage(HOSTSIM, solver)
    assert its > 0


@pytest.mark.gpu
def test_adjoint_solve_on_passage_cuda():
    its, res = solve_on_passage(None, "DATurboFoam", dims=(12, 12, 24))
    assert its > 0


def check_golden(lib_path, tol=1e-10):
    """the committed vectors of tests/golden/make_golden_cyclic.py (oracle on the ring of passages): DATurboFoam + SA + MRF, cyclic sides"""
    import os
    from tests.golden.make_golden_cyclic import SPEC
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "passage_turbo_cyclic_4x4x6.npz"))
    P = Pair(SPEC["turbulent"], SPEC["divU"], lib_path=lib_path, solver=SPEC["solver"], energy=SPEC["energy"], mrf_omega=SPEC["mrf_omega"])
    assert np.allclose(P.state(), g["W"], rtol=1e-12, atol=0.0)  # the generator of the state has not drifted either (libm may differ in the last bit)
    P.sol.updateOFFields(P.local(g["W"]))
    R = np.zeros(P.idx.size)
    P.sol.getResiduals(R)
    y = np.zeros(P.idx.size)
    P.sol.calcdRdWTPsiAD(P.local(g["psi"]), y)
    Rm, ym = P.merged(R), P.merged(y)
    for name, a, b in P.segments():
        assert rel_err(Rm[a:b], g["R"][a:b]) < tol, ("R", name)
        assert rel_err(ym[a:b], g["y"][a:b]) < tol, ("JTpsi", name)


def test_cyclic_golden_host_build():
    check_golden(HOSTSIM)


@pytest.mark.gpu
def test_cyclic_golden_cuda():
    check_golden(None)


@pytest.mark.gpu
def test_cyclic_passage_equals_ring_cuda():
    worst = check_pair(Pair(True, "linearUpwindV", lib_path=None))
    assert worst < 1e-9
    worst = check_pair(Pair(True, "linearUpwind", lib_path=None, mrf_omega=30.0))
    assert worst < 1e-9
