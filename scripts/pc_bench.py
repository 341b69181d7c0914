"""Adjoint solve on the bench O-grid for a list of adjEqnOption overrides (preconditioner ordering / block size / Krylov method).
env: PB_NJ (radial cells; ni = 2 nj), PB_TILE ("16x12"), PB_LIB (library; default the CUDA build), PB_CFGS (JSON list of adjEqnOption
overrides), PB_LVL (pcConLevel), PB_RESTART"""
import json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dafoam_b200 import cases
from dafoam_b200.pyDASolvers import pyDASolvers, Mat, KSP

nj = int(os.environ.get("PB_NJ", 720))
tile = tuple(int(v) for v in os.environ.get("PB_TILE", "16x12").split("x"))
mesh = cases.naca0012_ogrid(ni=2 * nj, nj=nj, nk=1, tile=tile)
d = tempfile.mkdtemp(prefix="dab_pb_")
cases.write_case(d, mesh, cases.default_bcs_naca(), binary=True)
fn = {"CD": {"type": "force", "source": "patchToFace", "patches": ["wing"], "directionMode": "fixedDirection", "direction": [1.0, 0.0, 0.0], "scale": 1.0}}
adj = dict(gmresRelTol=1e-6, gmresMaxIters=3000, gmresRestart=int(os.environ.get("PB_RESTART", 1500)), printInfo=0,
           pcConLevel=int(os.environ.get("PB_LVL", 2)), coarseAggregates=0)
opts = dict(normalizeStates=dict(U=10.0, p=50.0, nuTilda=1e-3, phi=1.0), function=fn, adjEqnOption=adj)
sol = pyDASolvers("DASimpleFoam -python", opts, caseDir=d, _lib_path=os.environ.get("PB_LIB") or None)
n = sol.getNLocalAdjointStates()
y = np.zeros(sol.getNLocalCells()); sol.getOFField("yWall", "scalar", y)
W = cases.boundary_layer_state(mesh, y, noise=0.001)
sol.updateOFFields(W)
dFdW = np.zeros(n)
sol.calcJacTVecProduct("states", "stateVar", W, "CD", "function", np.array([1.0]), dFdW)
ref = None
out = []
for cfg in json.loads(os.environ.get("PB_CFGS", "[{}]")):
    a = dict(adj, **cfg)
    if a.get("kspType") == "idrs":
        a["gmresMaxIters"] = 3 * adj["gmresMaxIters"]
    try:
        sol.updateDAOption(dict(opts, adjEqnOption=a))
        pc, ksp = Mat(), KSP()
        t0 = time.time(); sol.calcdRdWT(1, pc); sol.createMLRKSPMatrixFree(pc, ksp); tpc = time.time() - t0
        psi = np.zeros(n)
        t0 = time.time(); fail = sol.solveLinearEqn(ksp, dFdW, psi); t = time.time() - t0
        st = ksp.stats
        if ref is None and not fail:
            ref = psi.copy()
        res = dict(cells=sol.getNLocalCells(), cfg=cfg, its=st.iterations, n_matvec=st.n_matvec, fail=fail, pc_s=round(tpc, 2), solve_s=round(t, 2),
                   device_s=round(st.solve_seconds, 2), rel=st.final_residual / st.initial_residual,
                   psi_diff=float(np.linalg.norm(psi - ref) / np.linalg.norm(ref)) if ref is not None else None)
    except Exception as e:
        res = dict(cfg=cfg, error=str(e))
    print(json.dumps(res), flush=True)
    out.append(res)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/pc_bench_%s.json" % os.environ.get("PB_TAG", "x"), "w"), indent=1)
