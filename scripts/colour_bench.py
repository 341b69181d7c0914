"""Adjoint solve for several multicolour-ILU ordering radii (adjEqnOption.pcColourRadius) and Krylov methods (CB_KSP) on the bench O-grid.
env: CB_SOLVER (DASimpleFoam | DARhoSimpleFoam), CB_CELLS, CB_RADII="0,2,4", CB_LIB (library path; default = the CUDA build), CB_AGG (coarse aggregates), CB_RESTART"""
import json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dafoam_b200 import cases
from dafoam_b200.pyDASolvers import pyDASolvers, Mat, KSP

cells = int(os.environ.get("CB_CELLS", 250000))
nj = max(8, int(round((cells / 2.0) ** 0.5 / 2.0)) * 2)
mesh = cases.naca0012_ogrid(ni=2 * nj, nj=nj, nk=1)
d = tempfile.mkdtemp(prefix="dab_cb_")
comp = os.environ.get("CB_SOLVER", "DASimpleFoam") == "DARhoSimpleFoam"
U0c = (100.0, 0.0, 0.0)
thermo = cases.default_thermo() if comp else None
if comp:
    cases.write_case(d, mesh, cases.compressible_bcs(cases.default_bcs_naca(U0=U0c)), binary=True, thermo=thermo)
else:
    cases.write_case(d, mesh, cases.default_bcs_naca(), binary=True)
fn = {"CD": {"type": "force", "source": "patchToFace", "patches": ["wing"], "directionMode": "fixedDirection", "direction": [1.0, 0.0, 0.0], "scale": 1.0}}
adj = dict(gmresRelTol=1e-6, gmresMaxIters=3000, gmresRestart=int(os.environ.get("CB_RESTART", 1500)), printInfo=0, pcConLevel=3,
           coarseAggregates=int(os.environ.get("CB_AGG", 1000)))
ns = dict(U=100.0, p=101325.0, T=300.0, nuTilda=1e-3, phi=1.0) if comp else dict(U=10.0, p=50.0, nuTilda=1e-3, phi=1.0)
opts = dict(normalizeStates=ns, function=fn, adjEqnOption=adj)
if comp:
    opts["normalizeResiduals"] = ["URes", "pRes", "TRes", "nuTildaRes", "phiRes"]
sol = pyDASolvers(("DARhoSimpleFoam" if comp else "DASimpleFoam") + " -python", opts, caseDir=d, _lib_path=os.environ.get("CB_LIB") or None)
n = sol.getNLocalAdjointStates()
y = np.zeros(sol.getNLocalCells()); sol.getOFField("yWall", "scalar", y)
W = cases.boundary_layer_state(mesh, y, U0=U0c, seed=1234, noise=0.001) if comp else cases.boundary_layer_state(mesh, y, noise=0.001)
if comp:
    W = cases.to_compressible_state(mesh, W, thermo)
sol.updateOFFields(W)
dFdW = np.zeros(n)
sol.calcJacTVecProduct("states", "stateVar", W, "CD", "function", np.array([1.0]), dFdW)
out = []
ksps = os.environ.get("CB_KSP", "gmres").split(",")  # e.g. "gmres,idrs:4,idrs:8"
for r, ksp_name in [(int(x), kk) for x in os.environ.get("CB_RADII", "0,2,4").split(",") for kk in ksps]:
    kt, ks = (ksp_name.split(":") + ["4"])[:2]
    sol.updateDAOption(dict(opts, adjEqnOption=dict(adj, pcColourRadius=r, kspType=kt, idrS=int(ks),
                                                    gmresMaxIters=adj["gmresMaxIters"] * (3 if kt == "idrs" else 1))))
    pc, ksp = Mat(), KSP()
    t0 = time.time(); sol.calcdRdWT(1, pc); sol.createMLRKSPMatrixFree(pc, ksp); t_pc = time.time() - t0
    psi = np.zeros(n)
    t0 = time.time(); fail = sol.solveLinearEqn(ksp, dFdW, psi); t = time.time() - t0
    st = ksp.stats
    res = dict(cells=sol.getNLocalCells(), radius=r, ksp=ksp_name, n_matvec=st.n_matvec, fail=fail, iterations=st.iterations, pc_s=round(t_pc, 3), solve_s=round(t, 3),
               gmres_device_s=round(st.solve_seconds, 3), rel_residual=st.final_residual / st.initial_residual)
    print(json.dumps(res), flush=True)
    out.append(res)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/colour_bench_%d.json" % cells, "w"), indent=1)
