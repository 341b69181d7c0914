"""Adjoint solve on the bench mesh for several preconditioner settings in one process (one PC assembly).
env: AB_CONFIGS="gp:omega,gp:omega,..." (globalPCIters:richardsonOmega), AB_CELLS"""
import json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dafoam_b200 import cases
from dafoam_b200.pyDASolvers import pyDASolvers, Mat, KSP

cells = int(os.environ.get("AB_CELLS", 980000))
nj = max(8, int(round((cells / 2.0) ** 0.5 / 2.0)) * 2)
mesh = cases.naca0012_ogrid(ni=2 * nj, nj=nj, nk=1)
d = tempfile.mkdtemp(prefix="dab_ab_")
cases.write_case(d, mesh, cases.default_bcs_naca(), binary=True)
fn = {"CD": {"type": "force", "source": "patchToFace", "patches": ["wing"], "directionMode": "fixedDirection", "direction": [1.0, 0.0, 0.0], "scale": 1.0}}
adj = dict(gmresRelTol=1e-6, gmresMaxIters=3000, gmresRestart=1500, printInfo=0, pcConLevel=3, coarseAggregates=1000)
opts = dict(normalizeStates=dict(U=10.0, p=50.0, nuTilda=1e-3, phi=1.0), function=fn, adjEqnOption=adj)
sol = pyDASolvers("DASimpleFoam -python", opts, caseDir=d)
n = sol.getNLocalAdjointStates()
y = np.zeros(sol.getNLocalCells()); sol.getOFField("yWall", "scalar", y)
W = cases.boundary_layer_state(mesh, y, noise=0.001)
sol.updateOFFields(W)
dFdW = np.zeros(n)
sol.calcJacTVecProduct("states", "stateVar", W, "CD", "function", np.array([1.0]), dFdW)
pc, ksp = Mat(), KSP()
t0 = time.time(); sol.calcdRdWT(1, pc); sol.createMLRKSPMatrixFree(pc, ksp); t_pc = time.time() - t0
out = []
for cfg in os.environ.get("AB_CONFIGS", "0:1.0,1:1.0,2:1.0").split(","):
    gp, om = cfg.split(":")
    o2 = dict(opts, adjEqnOption=dict(adj, globalPCIters=int(gp), richardsonOmega=float(om)))
    sol.updateDAOption(o2)
    psi = np.zeros(n)
    t0 = time.time(); fail = sol.solveLinearEqn(ksp, dFdW, psi); t = time.time() - t0
    st = ksp.stats
    r = dict(globalPCIters=int(gp), omega=float(om), fail=fail, iterations=st.iterations, n_matvec=st.n_matvec, solve_s=t, gmres_device_s=st.solve_seconds,
             rel_residual=st.final_residual / st.initial_residual, pc_s=t_pc)
    print(json.dumps(r), flush=True)
    out.append(r)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/adjoint_bench.json", "w"), indent=1)
