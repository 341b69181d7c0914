"""ctypes wrapper of the CPU oracle (oracle/oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (dafoam_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

BC_KIND = {"fixedValue": 0, "zeroGradient": 1, "inletOutlet": 2, "outletInlet": 3, "symmetry": 4,
           "calculated": 5, "nutLowReWallFunction": 6, "nutUSpaldingWallFunction": 7}
GEOM_KIND = {"patch": 0, "wall": 1, "symmetry": 2}
DIV_SCHEME = {"upwind": 0, "linearUpwind": 1, "linear": 2, "linearUpwindV": 3}
FIELDS = ["U", "p", "nuTilda", "nut"]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        L = _LIB
        L.orc_create.restype = C.c_void_p
        L.orc_ndof.argtypes = [C.c_void_p]
        L.orc_ncells.argtypes = [C.c_void_p]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_record.restype = C.c_long
        L.orc_force.restype = C.c_double
        L.orc_tape_size.restype = C.c_long
    return _LIB


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def bc_tables(mesh, bcs):
    """Return (kind[4,nPatch] i32, value[4,nPatch,3] f64) from a cases.default_bcs_* dict."""
    npatch = len(mesh.patches)
    kind = np.full((4, npatch), BC_KIND["zeroGradient"], dtype=np.int32)
    value = np.zeros((4, npatch, 3))
    for fi, fname in enumerate(FIELDS):
        if fname not in bcs:
            continue
        pb = bcs[fname][3]
        for pi, patch in enumerate(mesh.patches):
            bc = pb[patch["name"]]
            kind[fi, pi] = BC_KIND[bc["type"]]
            ref = None
            for key in ("inletValue", "outletValue", "value"):
                if key in bc:
                    ref = bc[key]
                    break
            if ref is not None:
                ref = np.atleast_1d(np.asarray(ref, dtype=np.float64))
                value[fi, pi, :ref.size] = ref
    return kind, value


class Oracle:
    def __init__(self, mesh, bcs, nu=1.5e-5, alphaU=0.7, divU="linearUpwind", divNut="upwind",
                 normalizeStates=None, normalizeResiduals=("URes", "pRes", "nuTildaRes", "phiRes"),
                 constrainHbyA=True, yWall=None, rasModel="SpalartAllmaras", thermo=None):
        """thermo: None (DASimpleFoam) or the dict of cases.default_thermo() -> DARhoSimpleFoam (bcs must carry "T")."""
        L = lib()
        self.mesh = mesh
        self.turb = "nuTilda" in bcs
        ns = dict(U=1.0, p=1.0, nuTilda=1.0, phi=1.0, T=1.0)
        ns.update(normalizeStates or {})
        kind, value = bc_tables(mesh, bcs)
        nf = mesh.faces.shape[0]
        foff, flab = mesh.face_offsets_labels()
        foff = np.ascontiguousarray(foff, dtype=np.int32)
        flab = np.ascontiguousarray(flab, dtype=np.int32)
        pstart = np.array([p["start"] for p in mesh.patches], dtype=np.int32)
        psize = np.array([p["size"] for p in mesh.patches], dtype=np.int32)
        pgeom = np.array([GEOM_KIND.get(p["type"], 0) for p in mesh.patches], dtype=np.int32)
        dpar = np.array([nu, alphaU, ns["U"], ns["p"], ns["nuTilda"], ns.get("phi", 1.0)], dtype=np.float64)
        ipar = np.array([int(self.turb) * (2 if rasModel == "SpalartAllmarasFv3" else 1), DIV_SCHEME[divU], DIV_SCHEME[divNut],
                         int("URes" in normalizeResiduals), int("pRes" in normalizeResiduals),
                         int("nuTildaRes" in normalizeResiduals), int("phiRes" in normalizeResiduals),
                         int(constrainHbyA), int("phi" in (normalizeStates or {"phi": 1.0}))], dtype=np.int32)
        yw = None if yWall is None else _p(np.ascontiguousarray(yWall, dtype=np.float64))
        self._keep = (foff, flab, pstart, psize, pgeom, kind, value, dpar, ipar)
        self.h = C.c_void_p(L.orc_create(
            C.c_int(mesh.n_points), _p(mesh.points), C.c_int(nf), _p(foff, C.c_int), _p(flab, C.c_int),
            _p(mesh.owner, C.c_int), C.c_int(mesh.n_internal_faces), _p(mesh.neighbour, C.c_int),
            C.c_int(len(mesh.patches)), _p(pstart, C.c_int), _p(psize, C.c_int), _p(pgeom, C.c_int),
            _p(kind, C.c_int), _p(value), _p(dpar), _p(ipar, C.c_int), yw))
        self.compressible = thermo is not None
        if thermo is not None:
            th = thermo
            pbT = bcs["T"][3]
            kindT = np.array([BC_KIND[pbT[p_["name"]]["type"]] for p_ in mesh.patches], dtype=np.int32)
            valT = np.zeros(len(mesh.patches))
            for pi, p_ in enumerate(mesh.patches):
                bc = pbT[p_["name"]]
                for key in ("inletValue", "outletValue", "value"):
                    if key in bc:
                        valT[pi] = float(bc[key])
                        break
            Rg = 8314.4700665 / th["molWeight"]
            dth = np.array([Rg, th["Cp"], th["mu"], th["Pr"], th["Prt"], th.get("As", 1.4792e-6), th.get("Ts", 116.0), 298.15,
                            ns.get("T", 1.0)], dtype=np.float64)
            ith = np.array([int(th["energy"] == "sensibleInternalEnergy"), int(th["transport"] == "sutherland"),
                            DIV_SCHEME[th.get("divE", "upwind")], DIV_SCHEME[th.get("divEkp", "upwind")],
                            int("TRes" in normalizeResiduals)], dtype=np.int32)
            self._keep_th = (dth, ith, kindT, valT)
            L.orc_set_compressible(self.h, _p(dth), _p(ith, C.c_int), _p(kindT, C.c_int), _p(valT))
        self.ndof = L.orc_ndof(self.h)
        self.ncells = L.orc_ncells(self.h)

    def __del__(self):
        try:
            lib().orc_destroy(self.h)
        except Exception:
            pass

    def geometry(self, what):
        sizes = {"V": (0, self.ncells), "magSf": (1, self.mesh.n_faces), "w": (2, self.mesh.n_faces),
                 "delta": (3, self.mesh.n_faces), "yWall": (4, self.ncells), "C": (5, 3 * self.ncells),
                 "Sf": (6, 3 * self.mesh.n_faces), "Cf": (7, 3 * self.mesh.n_faces), "corr": (8, 3 * self.mesh.n_faces)}
        code, n = sizes[what]
        out = np.zeros(n)
        lib().orc_get_geometry(self.h, C.c_int(code), _p(out))
        return out

    def residual(self, W, isPC=0):
        W = np.ascontiguousarray(W, dtype=np.float64)
        R = np.zeros(self.ndof)
        lib().orc_residual(self.h, _p(W), C.c_int(isPC), _p(R))
        return R

    def record(self, W, isPC=0):
        W = np.ascontiguousarray(W, dtype=np.float64)
        return lib().orc_record(self.h, _p(W), C.c_int(isPC))

    def nut_fvmatrix(self, W, alpha=0.7):
        """(D, upper, lower) of the relaxed nuTilda fvMatrix with the div(pc) scheme (DASpalartAllmaras::getFvMatrixFields)."""
        D, up, lo = np.zeros(self.ncells), np.zeros(self.mesh.n_internal_faces), np.zeros(self.mesh.n_internal_faces)
        lib().orc_nut_fvmatrix(self.h, _p(np.ascontiguousarray(W, dtype=np.float64)), C.c_double(alpha), _p(D), _p(up), _p(lo))
        return D, up, lo

    def jvec(self, W, v, isPC=0):
        """J v by forward-mode dual numbers (exact tangent; the independent check of the tape)."""
        out = np.zeros(self.ndof)
        lib().orc_jvec(self.h, _p(np.ascontiguousarray(W, dtype=np.float64)), _p(np.ascontiguousarray(v, dtype=np.float64)), C.c_int(isPC), _p(out))
        return out

    def jtvec(self, psi, normalize=True):
        psi = np.ascontiguousarray(psi, dtype=np.float64)
        out = np.zeros(self.ndof)
        lib().orc_jtvec(self.h, _p(psi), _p(out), C.c_int(int(normalize)))
        return out

    def force(self, W, patch, direction, scale=1.0, center=None, mode=None):
        """force . direction, or (center given) the moment ((Cf - center) x force) . direction; mode 2: area-averaged total
        pressure, mode 3: mass flow rate of the patch (direction unused)."""
        W = np.ascontiguousarray(W, dtype=np.float64)
        d = np.ascontiguousarray(direction, dtype=np.float64)
        ctr = np.zeros(3) if center is None else np.ascontiguousarray(center, dtype=np.float64)
        m = mode if mode is not None else (0 if center is None else 1)
        return lib().orc_force(self.h, _p(W), C.c_int(patch), _p(d), C.c_double(scale), C.c_int(m), _p(ctr))

    def dforce_dw(self, W, patch, direction, scale=1.0, seed=1.0, normalize=True, center=None, mode=None):
        W = np.ascontiguousarray(W, dtype=np.float64)
        d = np.ascontiguousarray(direction, dtype=np.float64)
        ctr = np.zeros(3) if center is None else np.ascontiguousarray(center, dtype=np.float64)
        out = np.zeros(self.ndof)
        lib().orc_dforce_dw(self.h, _p(W), C.c_int(patch), _p(d), C.c_double(scale), C.c_double(seed), _p(out),
                            C.c_int(int(normalize)), C.c_int(mode if mode is not None else (0 if center is None else 1)), _p(ctr))
        return out

    def jtvec_bcU(self, W, psi, patch):
        """[dR/d(U boundary reference value of `patch`)]^T psi (3 numbers)."""
        W = np.ascontiguousarray(W, dtype=np.float64)
        psi = np.ascontiguousarray(psi, dtype=np.float64)
        out = np.zeros(3)
        lib().orc_jtvec_bcU(self.h, _p(W), _p(psi), C.c_int(patch), _p(out))
        return out

    def set_fvsource(self, fvSource):
        """fvSource option of the reference (actuatorDisk / cylinderAnnulusSmooth, adjustThrust 0)."""
        rows = []
        for d in fvSource.values():
            rows.append(list(d["center"]) + list(d["direction"]) + [d["innerRadius"], d["outerRadius"], d["scale"], d["POD"], d["expM"], d["expN"],
                                                                       d.get("targetThrust", 1.0), d["eps"], 1.0 if d.get("rotDir", "right") == "left" else 0.0])
        a = np.ascontiguousarray(rows, dtype=np.float64).ravel()
        self._keep_disks = a
        lib().orc_set_fvsource(self.h, C.c_int(len(rows)), _p(a))

    def set_turbo(self, on=True):
        """DATurboFoam's energy equation (the enthalpy form carries the viscous-work and p(U - URel) terms)."""
        lib().orc_set_turbo(self.h, C.c_int(int(on)))

    def set_transonic(self, on=True, scheme=0, k=1.0, pc_option=-1):
        """simple_.transonic(): pEqn = fvm::div(phid, p) - fvm::laplacian(rho rAU, p) (DAResidualTurboFoam.C:148-189,
        DAResidualRhoSimpleCFoam.C:148-200).  scheme: 0 upwind, 2 linear, 4 limitedLinear k; pc_option = transonicPCOption."""
        lib().orc_set_transonic(self.h, C.c_int(int(on)), C.c_int(int(scheme)), C.c_double(float(k)), C.c_int(int(pc_option)))

    def set_mrf(self, mesh, mrf):
        """One MRF zone (see dafoam_b200.cases.write_mrf for the dict)."""
        axis = np.asarray(mrf["axis"], dtype=np.float64)
        omega = np.ascontiguousarray(mrf["omega"] * axis / np.linalg.norm(axis))
        origin = np.ascontiguousarray(mrf["origin"], dtype=np.float64)
        mask = np.zeros(mesh.n_cells, dtype=np.int32)
        mask[np.asarray(mrf["cells"], dtype=np.int64)] = 1
        excl = np.ascontiguousarray([1 if p["name"] in mrf.get("nonRotatingPatches", []) else 0 for p in mesh.patches], dtype=np.int32)
        lib().orc_set_mrf(self.h, _p(omega), _p(origin), mask.ctypes.data_as(C.POINTER(C.c_int)), excl.ctypes.data_as(C.POINTER(C.c_int)))

    def jtvec_bc(self, W, psi, field, patch):
        """[dR/d(boundary reference value of `field` on `patch`)]^T psi (3 numbers; scalars use the first)."""
        W = np.ascontiguousarray(W, dtype=np.float64)
        psi = np.ascontiguousarray(psi, dtype=np.float64)
        out = np.zeros(3)
        lib().orc_jtvec_bc(self.h, _p(W), _p(psi), C.c_int(FIELDS.index(field)), C.c_int(patch), _p(out))
        return out

    def set_bc_value(self, field, patch, value):
        v = np.zeros(3)
        v[:len(np.atleast_1d(value))] = value
        lib().orc_set_bc_value(self.h, C.c_int(FIELDS.index(field)), C.c_int(patch), _p(v))

    def jtvec_xv(self, W, psi):
        W = np.ascontiguousarray(W, dtype=np.float64)
        psi = np.ascontiguousarray(psi, dtype=np.float64)
        out = np.zeros(3 * self.mesh.n_points)
        lib().orc_jtvec_xv(self.h, _p(W), _p(psi), _p(out))
        return out


def synthetic_state(mesh, geomC, Sf, U0=(10.0, 0.5, 0.0), nuTilda0=4.5e-5, turbulent=True, seed=1234, noise=0.01, thermo=None,
                    p0=101325.0, T0=300.0):
    """Smooth analytic field + seeded 1 % noise: U, p, nuTilda at cells and a
    face flux phi = U_f . Sf (+ noise) on all faces, in the reference's state ordering."""
    rng = np.random.default_rng(seed)
    nC = geomC.size // 3
    Cc = geomC.reshape(nC, 3)
    U0 = np.asarray(U0, dtype=np.float64)
    Umag = np.linalg.norm(U0)
    r2 = (Cc[:, 0] - 0.5) ** 2 + Cc[:, 1] ** 2
    damp = 1.0 - np.exp(-r2 / 0.5)
    U = U0[None, :] * damp[:, None]
    U[:, 0] += 0.1 * Umag * np.sin(1.3 * Cc[:, 1]) * damp
    U[:, 1] += 0.1 * Umag * np.cos(0.7 * Cc[:, 0]) * damp
    U *= 1.0 + noise * rng.uniform(-1, 1, U.shape)
    p = 0.5 * Umag**2 * (np.exp(-r2) - 0.3 * np.sin(Cc[:, 0])) * (1.0 + noise * rng.uniform(-1, 1, nC))
    nt = nuTilda0 * (1.0 + 3.0 * np.exp(-r2 / 4.0)) * (1.0 + noise * rng.uniform(-1, 1, nC))
    nF = mesh.n_faces
    nIF = mesh.n_internal_faces
    S = Sf.reshape(nF, 3)
    Uf = np.empty((nF, 3))
    Uf[:nIF] = 0.5 * (U[mesh.owner[:nIF]] + U[mesh.neighbour])
    Uf[nIF:] = U[mesh.owner[nIF:]]
    phi = np.einsum("ij,ij->i", Uf, S)
    magS = np.linalg.norm(S, axis=1)
    phi += noise * Umag * magS * rng.uniform(-1, 1, nF)
    # symmetry-plane and wall faces carry no flux
    for pch in mesh.patches:
        if pch["type"] in ("symmetry", "wall"):
            phi[pch["start"]:pch["start"] + pch["size"]] = 0.0
    parts = [U.ravel(), p]
    if thermo is not None:
        # compressible: absolute pressure, temperature field, mass flux
        Rg = 8314.4700665 / thermo["molWeight"]
        p = p0 + p
        Tt = T0 * (1.0 + 0.02 * np.sin(0.9 * Cc[:, 0]) * np.cos(1.1 * Cc[:, 1])) * (1.0 + 0.1 * noise * rng.uniform(-1, 1, nC))
        rho = p / (Rg * Tt)
        rf = np.empty(nF)
        rf[:nIF] = 0.5 * (rho[mesh.owner[:nIF]] + rho[mesh.neighbour])
        rf[nIF:] = rho[mesh.owner[nIF:]]
        phi = phi * rf
        parts = [U.ravel(), p, Tt]
    if turbulent:
        parts.append(nt)
    parts.append(phi)
    return np.concatenate(parts)
