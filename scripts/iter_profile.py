"""A short IDR(s) adjoint solve (a few dozen operator applications) on the bench O-grid, for a per-kernel launch list under
ncu (scripts/gpu_iter_profile.sh): where the ~5 ms per application go."""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dafoam_b200 import cases
from dafoam_b200.pyDASolvers import pyDASolvers, Mat, KSP

nj = int(os.environ.get("IP_NJ", 720))
mesh = cases.naca0012_ogrid(ni=2 * nj, nj=nj, nk=1, tile=(16, 12))
d = tempfile.mkdtemp(prefix="dab_ip_")
cases.write_case(d, mesh, cases.default_bcs_naca(), binary=True)
fn = {"CD": {"type": "force", "source": "patchToFace", "patches": ["wing"], "directionMode": "fixedDirection", "direction": [1.0, 0.0, 0.0], "scale": 1.0}}
adj = dict(gmresRelTol=1e-6, gmresMaxIters=int(os.environ.get("IP_ITERS", 40)), gmresRestart=100, printInfo=0, pcConLevel=3, coarseAggregates=2000,
           kspType="idrs", idrS=8, pcStorage=os.environ.get("IP_STORAGE", "fp32"))
sol = pyDASolvers("DASimpleFoam -python", dict(normalizeStates=dict(U=10.0, p=50.0, nuTilda=1e-3, phi=1.0), function=fn, adjEqnOption=adj), caseDir=d)
n = sol.getNLocalAdjointStates()
y = np.zeros(sol.getNLocalCells()); sol.getOFField("yWall", "scalar", y)
W = cases.boundary_layer_state(mesh, y, noise=0.001)
sol.updateOFFields(W)
dFdW = np.zeros(n)
sol.calcJacTVecProduct("states", "stateVar", W, "CD", "function", np.array([1.0]), dFdW)
pc, ksp = Mat(), KSP()
sol.calcdRdWT(1, pc); sol.createMLRKSPMatrixFree(pc, ksp)
print("PC DONE", flush=True)
psi = np.zeros(n)
sol.solveLinearEqn(ksp, dFdW, psi)
print("applications", ksp.stats.iterations, "device_s", ksp.stats.solve_seconds, flush=True)
