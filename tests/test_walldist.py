"""Option wallDistCorrectWalls (OpenFOAM wallDist `correctWalls`; reference src/adjoint/DAMisc/meshWaveFrozen keeps OpenFOAM's patchWave
with its wall correction): cells that touch a wall get the exact distance to the wall faces around them instead of the distance to the
nearest wall-face centre.  Checked against a brute-force numpy evaluation (closest point on the triangle fan of every wall face), and
the residual / transposed product against the oracle fed with that wall distance."""
import tempfile

import numpy as np

from dafoam_b200 import cases
from dafoam_b200.pyDASolvers import pyDASolvers
from oracle.pyoracle import Oracle, synthetic_state
from tests.common import HOSTSIM, NORM_STATES, ALL_RES, rel_err


def _closest_on_triangle(a, b, c, p):
    ab, ac, ap = b - a, c - a, p - a
    d1, d2 = ab @ ap, ac @ ap
    if d1 <= 0 and d2 <= 0:
        return a
    bp = p - b
    d3, d4 = ab @ bp, ac @ bp
    if d3 >= 0 and d4 <= d3:
        return b
    vc = d1 * d4 - d3 * d2
    if vc <= 0 and d1 >= 0 and d3 <= 0:
        return a + d1 / (d1 - d3) * ab
    cp = p - c
    d5, d6 = ab @ cp, ac @ cp
    if d6 >= 0 and d5 <= d6:
        return c
    vb = d5 * d2 - d1 * d6
    if vb <= 0 and d2 >= 0 and d6 <= 0:
        return a + d2 / (d2 - d6) * ac
    va = d3 * d6 - d5 * d4
    if va <= 0 and d4 - d3 >= 0 and d5 - d6 >= 0:
        return b + (d4 - d3) / ((d4 - d3) + (d5 - d6)) * (c - b)
    den = 1.0 / (va + vb + vc)
    return a + ab * (vb * den) + ac * (vc * den)


def _dist_to_face(pts, ctr, p):
    n = len(pts)
    return min(np.linalg.norm(p - _closest_on_triangle(pts[i], pts[(i + 1) % n], ctr, p)) for i in range(n))


def test_corrected_wall_distance_host_build():
    mesh = cases.naca0012_ogrid(ni=40, nj=20, nk=2)
    bcs = cases.default_bcs_naca()
    d = tempfile.mkdtemp(prefix="dab_yw_")
    cases.write_case(d, mesh, bcs)
    opts = dict(normalizeStates=NORM_STATES, normalizeResiduals=list(ALL_RES))
    plain = pyDASolvers("DASimpleFoam -python", opts, caseDir=d, _lib_path=HOSTSIM)
    corr = pyDASolvers("DASimpleFoam -python", dict(opts, wallDistCorrectWalls=True), caseDir=d, _lib_path=HOSTSIM)
    nC = mesh.n_cells
    y0, y1, C = np.zeros(nC), np.zeros(nC), np.zeros(3 * nC)
    plain.getOFField("yWall", "scalar", y0)
    corr.getOFField("yWall", "scalar", y1)
    orc0 = Oracle(mesh, bcs, normalizeStates=NORM_STATES, normalizeResiduals=ALL_RES)
    C = np.asarray(orc0.geometry("C")).reshape(nC, 3)
    Cf = np.asarray(orc0.geometry("Cf")).reshape(-1, 3)
    assert rel_err(y0, np.asarray(orc0.geometry("yWall"))) < 1e-13
    wall = next(p for p in mesh.patches if p["type"] == "wall")
    wf = range(wall["start"], wall["start"] + wall["size"])
    wall_pts = set(int(v) for f in wf for v in mesh.faces[f] if v >= 0)
    # cells touching the wall with at least a point
    touching = set()
    for f in range(mesh.n_faces):
        if any(int(v) in wall_pts for v in mesh.faces[f] if v >= 0):
            touching.add(int(mesh.owner[f]))
            if f < mesh.n_internal_faces:
                touching.add(int(mesh.neighbour[f]))
    changed = np.nonzero(y1 != y0)[0]
    assert set(changed.tolist()) <= touching and len(changed) >= wall["size"]
    for c in sorted(touching):
        ref = min(_dist_to_face(mesh.points[[v for v in mesh.faces[f] if v >= 0]], Cf[f], C[c]) for f in wf)
        assert abs(y1[c] - ref) <= 1e-12 * ref, (c, y1[c], ref)
        assert y1[c] <= y0[c] * (1.0 + 1e-12)  # the closest point of a face is never farther than its centre
    assert np.array_equal(y1[[c for c in range(nC) if c not in touching]], y0[[c for c in range(nC) if c not in touching]])
    # residual and transposed product with the corrected distance: the oracle takes it as an input
    orc = Oracle(mesh, bcs, normalizeStates=NORM_STATES, normalizeResiduals=ALL_RES, yWall=y1)
    W = synthetic_state(mesh, orc.geometry("C"), orc.geometry("Sf"))
    corr.updateOFFields(W)
    R = np.zeros(orc.ndof)
    corr.getResiduals(R)
    assert rel_err(R, orc.residual(W)) < 1e-10
    orc.record(W)
    psi = np.random.default_rng(5).uniform(-1, 1, orc.ndof)
    y = np.zeros(orc.ndof)
    corr.calcdRdWTPsiAD(psi, y)
    assert rel_err(y, orc.jtvec(psi)) < 1e-10
    # and it matters: the nuTilda rows differ from the uncorrected ones
    plain.updateOFFields(W)
    Rp = np.zeros(orc.ndof)
    plain.getResiduals(Rp)
    assert rel_err(Rp[4 * nC:5 * nC], R[4 * nC:5 * nC]) > 1e-6
