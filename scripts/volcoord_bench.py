"""[dR/dx_v]^T psi on the bench mesh on the GPU: set-up (colouring) and product seconds.  env: VB_NI, VB_NJ"""
import json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dafoam_b200 import cases
from dafoam_b200.pyDASolvers import pyDASolvers

ni, nj = int(os.environ.get("VB_NI", 1400)), int(os.environ.get("VB_NJ", 700))
mesh = cases.naca0012_ogrid(ni=ni, nj=nj, nk=1)
d = tempfile.mkdtemp(prefix="dab_vb_")
cases.write_case(d, mesh, cases.default_bcs_naca(), binary=True)
fn = {"CD": {"type": "force", "source": "patchToFace", "patches": ["wing"], "directionMode": "fixedDirection", "direction": [1.0, 0.0, 0.0], "scale": 1.0}}
sol = pyDASolvers("DASimpleFoam -python", dict(normalizeStates=dict(U=10.0, p=50.0, nuTilda=1e-3, phi=1.0), function=fn, adjEqnOption=dict(printInfo=1)), caseDir=d)
n, nP3 = sol.getNLocalAdjointStates(), 3 * sol.getNLocalPoints()
pts = np.zeros(nP3)
sol.getOFMeshPoints(pts)
psi = np.random.default_rng(7).uniform(-1, 1, n)
prod = np.zeros(nP3)
t0 = time.time()
sol.calcJacTVecProduct("x", "volCoord", pts, "R", "residual", psi, prod)
t1 = time.time()
sol.calcJacTVecProduct("x", "volCoord", pts, "R", "residual", psi, prod)
t2 = time.time()
dFdx = np.zeros(nP3)
sol.calcJacTVecProduct("x", "volCoord", pts, "CD", "function", np.array([1.0]), dFdx)
t3 = time.time()
out = dict(cells=mesh.n_cells, points=nP3 // 3, first_call_s=t1 - t0, product_s=t2 - t1, function_s=t3 - t2, norm=float(np.linalg.norm(prod)), norm_dFdx=float(np.linalg.norm(dFdx)))
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/volcoord_bench.json", "w"), indent=1)
