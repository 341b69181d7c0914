"""SIMPLE primal on the bench mesh (NACA0012 O-grid) on the GPU: iterations, residuals, seconds; then |R(W)|.
env: PB_NI, PB_NJ (default 1400x700), PB_ITERS (default 2000), PB_TOL (default 1e-8)"""
import json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dafoam_b200 import cases
from dafoam_b200.pyDASolvers import pyDASolvers

ni, nj = int(os.environ.get("PB_NI", 1400)), int(os.environ.get("PB_NJ", 700))
iters, tol = int(os.environ.get("PB_ITERS", 2000)), float(os.environ.get("PB_TOL", 1e-8))
t0 = time.time()
mesh = cases.naca0012_ogrid(ni=ni, nj=nj, nk=1)
d = tempfile.mkdtemp(prefix="dab_pb_")
cases.write_case(d, mesh, cases.default_bcs_naca(U0=(10.0 * np.cos(np.deg2rad(3.0)), 10.0 * np.sin(np.deg2rad(3.0)), 0.0)), binary=True)
fn = {"CD": {"type": "force", "source": "patchToFace", "patches": ["wing"], "directionMode": "fixedDirection",
             "direction": [float(np.cos(np.deg2rad(3.0))), float(np.sin(np.deg2rad(3.0))), 0.0], "scale": 1.0 / (0.5 * 100 * 0.1)},
      "CL": {"type": "force", "source": "patchToFace", "patches": ["wing"], "directionMode": "fixedDirection",
             "direction": [-float(np.sin(np.deg2rad(3.0))), float(np.cos(np.deg2rad(3.0))), 0.0], "scale": 1.0 / (0.5 * 100 * 0.1)}}
sol = pyDASolvers("DASimpleFoam -python", dict(normalizeStates=dict(U=10.0, p=50.0, nuTilda=1e-3, phi=1.0), function=fn,
                                                primalMinResTol=tol, primalMaxIters=iters, printInterval=50,
                                                adjEqnOption=dict(printInfo=1)), caseDir=d)
n = sol.getNLocalAdjointStates()
R = np.zeros(n)
sol.getResiduals(R)
r0 = float(np.linalg.norm(R))
t1 = time.time()
fail = sol.solvePrimal()
st = sol.primalStats
sol.getResiduals(R)
out = dict(cells=mesh.n_cells, setup_sec=t1 - t0, fail=fail, iterations=st.iterations, converged=st.converged, max_residual=st.max_residual,
           res_u=list(st.res_u), res_p=st.res_p, res_nutilda=st.res_nutilda, p_iterations=st.p_iterations, seconds=st.seconds,
           ms_per_iteration=1e3 * st.seconds / max(st.iterations, 1), R0=r0, R=float(np.linalg.norm(R)),
           CD=sol.calcFunction("CD"), CL=sol.calcFunction("CL"))
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/primal_bench.json", "w"), indent=1)
