#!/usr/bin/env python
"""bench.py -- the adjoint hot path on B200: dRdW^T*psi throughput (GCells/s) and adjoint-solve wall time.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cells C]

A "step" is one matrix-free product y = diag(n) (dR/dW)^T psi over the whole mesh (the body of the
reference's GMRES shell-matrix callback, DASolver.C:1364-1409).  Workload: BASELINE.json configs[1],
"DASimpleFoam NACA0012 SA turbulence 1M cells" (synthetic O-grid 1400x700x1 = 980k cells, analytic state
+ 1 % seeded noise; the reference ships no mesh).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

NORM_STATES = dict(U=10.0, p=50.0, nuTilda=1e-3, phi=1.0)


def grid_for(cells):
    nj = max(8, int(round((cells / 2.0) ** 0.5 / 2.0)) * 2)
    ni = 2 * nj
    return ni, nj


def smooth_state(sol, mesh, seed=1234, noise=0.001):
    """Smooth analytic boundary-layer state (+0.1 % seeded noise) -- SURVEY.md section 8d; the reference
    would supply a converged primal here, which needs the (out-of-scope) primal solver."""
    from dafoam_b200 import cases
    y = np.zeros(sol.getNLocalCells())
    sol.getOFField("yWall", "scalar", y)
    return cases.boundary_layer_state(mesh, y, seed=seed, noise=noise)


class ClockSampler:
    def __init__(self):
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def start(self):
        def run():
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap")
            while not self._stop.is_set():
                try:
                    out = subprocess.run(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-i", "0"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
                except Exception:
                    pass
                self._stop.wait(0.1)
        self._t = threading.Thread(target=run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU reference arm: the oracle port of the reference algorithm (tape record once, evaluate per product)
# ---------------------------------------------------------------------------------------------------
_WORKER = r"""
import sys, time, json
sys.path.insert(0, %(root)r)
import numpy as np
from dafoam_b200 import cases
from oracle.pyoracle import Oracle, synthetic_state
ni, nj, reps = %(ni)d, %(nj)d, %(reps)d
m = cases.naca0012_ogrid(ni=ni, nj=nj, nk=1)
o = Oracle(m, cases.default_bcs_naca(), normalizeStates=%(ns)r)
W = synthetic_state(m, o.geometry("C"), o.geometry("Sf"))
t0 = time.time(); o.record(W); trec = time.time() - t0
psi = np.random.default_rng(4321).uniform(-1, 1, o.ndof)
o.jtvec(psi)
print("READY", flush=True)
sys.stdin.readline()
t0 = time.time()
for _ in range(reps):
    o.jtvec(psi)
dt = time.time() - t0
print(json.dumps(dict(cells=m.n_cells, reps=reps, seconds=dt, record_seconds=trec)), flush=True)
"""


def cpu_reference(cores, cells_per_proc=6000, reps=20):
    """All `cores` host cores, one process per core (the reference runs one MPI rank per core), each
    evaluating the recorded tape of a `cells_per_proc`-cell partition `reps` times."""
    ni, nj = grid_for(cells_per_proc)
    code = _WORKER % dict(root=ROOT, ni=ni, nj=nj, reps=reps, ns=NORM_STATES)
    procs = [subprocess.Popen([sys.executable, "-c", code], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True)
             for _ in range(cores)]
    for p in procs:
        assert p.stdout.readline().strip() == "READY"
    t0 = time.time()
    for p in procs:
        p.stdin.write("go\n")
        p.stdin.flush()
    res = [json.loads(p.stdout.readline()) for p in procs]
    wall = time.time() - t0
    for p in procs:
        p.wait()
    cells = sum(r["cells"] for r in res)
    reps_ = res[0]["reps"]
    return dict(value=cells * reps_ / wall / 1e9, unit="GCells/s", cores=cores, kind="port",
                sample="%d processes x %d-cell NACA0012 partition x %d tape evaluations (oracle port of the reference's "
                       "CoDiPack tape-evaluate matvec; wall %.2f s)" % (cores, res[0]["cells"], reps_, wall),
                seconds_per_product_per_partition=wall / reps_, tape_record_seconds=res[0]["record_seconds"])


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    ni, nj = grid_for(args.cells)
    steps = max(1, args.steps)
    t_all = []
    base = None
    for i in range(args.warmup + steps):
        base = cpu_reference(cores, reps=5)
        if i >= args.warmup:
            t_all.append(base["value"])
        if i >= args.warmup + 2:  # bounded: the CPU arm is slow
            break
    v = float(np.mean(t_all))
    base["value"] = v
    out = {"impl": "reference", "metric": "dRdWTPsi_GCells_per_s", "value": v, "unit": "GCells/s", "n_gpus": args.gpus,
           "steps": len(t_all), "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": args.solver + " NACA0012 SA %dx%dx1 O-grid (CPU arm: bounded sample of partitions)" % (ni, nj)},
           "cpu_baseline": base,
           "e2e": {"value": v, "unit": "GCells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


# ---------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def emit(obj):
    """The one JSON line goes to the real stdout; everything else any library prints (e.g. NCCL's version banner)
    was redirected to stderr at start-up."""
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--cells", type=int, default=980000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-solve", action="store_true")
    ap.add_argument("--no-idrs", action="store_true", help="skip the IDR(s) leg of the adjoint solve")
    ap.add_argument("--solve-multi", action="store_true")
    ap.add_argument("--restart", type=int, default=1500)
    ap.add_argument("--pc-level", type=int, default=3)
    ap.add_argument("--coarse", type=int, default=1000)
    ap.add_argument("--max-iters", type=int, default=3000)
    ap.add_argument("--solver", default="DASimpleFoam", choices=["DASimpleFoam", "DARhoSimpleFoam"],
                    help="DARhoSimpleFoam: BASELINE config 3 (compressible airfoil; use --cells 2000000); single GPU")
    ap.add_argument("--primal-iters", type=int, default=0,
                    help="run that many SIMPLE iterations (solvePrimal on the GPU) from the synthetic state before the adjoint legs (1 GPU)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    from dafoam_b200 import cases
    from dafoam_b200.pyDASolvers import pyDASolvers

    # weak scaling: the global mesh has world x args.cells cells and is split into `world` sub-domains (RCB)
    ni, nj = grid_for(args.cells * world)
    t_setup = time.time()
    mesh = cases.naca0012_ogrid(ni=ni, nj=nj, nk=1)
    shared = [None]
    comp = args.solver == "DARhoSimpleFoam"
    U0c = (100.0, 0.0, 0.0)  # M ~ 0.29 at 300 K
    thermo = cases.default_thermo() if comp else None
    if rank == 0:
        shared[0] = tempfile.mkdtemp(prefix="dab_bench_")
        if comp:
            cases.write_case(shared[0], mesh, cases.compressible_bcs(cases.default_bcs_naca(U0=U0c)), binary=True, thermo=thermo)
        else:
            cases.write_case(shared[0], mesh, cases.default_bcs_naca(), binary=True)
    uid = None
    if world > 1:
        from dafoam_b200.pyDASolvers import nccl_unique_id
        box = [shared[0], nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        shared[0], uid = box
    case_dir = shared[0]
    fn = {"CD": {"type": "force", "source": "patchToFace", "patches": ["wing"], "directionMode": "fixedDirection",
                 "direction": [1.0, 0.0, 0.0], "scale": 1.0}}
    ns_opt = dict(U=100.0, p=101325.0, T=300.0, nuTilda=1e-3, phi=1.0) if comp else NORM_STATES
    opts = dict(normalizeStates=ns_opt, function=fn, primalMaxIters=max(args.primal_iters, 1), primalMinResTol=1e-8, printInterval=100,
                adjEqnOption=dict(gmresRelTol=1e-6, gmresMaxIters=args.max_iters, gmresRestart=args.restart, printInfo=1, pcConLevel=args.pc_level, coarseAggregates=args.coarse))
    sol = pyDASolvers(args.solver + " -python", opts, caseDir=case_dir, device=local_rank, rank=rank, nRanks=world, ncclUniqueId=uid)
    n = sol.getNLocalAdjointStates()
    nC = sol.getNLocalCells()
    if world == 1:
        if comp:
            y_ = np.zeros(sol.getNLocalCells())
            sol.getOFField("yWall", "scalar", y_)
            W = cases.to_compressible_state(mesh, cases.boundary_layer_state(mesh, y_, U0=U0c, seed=1234, noise=0.001), thermo)
        else:
            W = smooth_state(sol, mesh)
    else:
        # global analytic state (wall distance by a KD-tree on the wall-face centres), then this rank's slice
        from scipy.spatial import cKDTree
        Sf, Cf = cases.quad_face_geometry(mesh)
        wall = [p for p in mesh.patches if p["type"] == "wall"][0]
        nIF = mesh.n_internal_faces
        Cc = np.zeros((mesh.n_cells, 3))
        cnt = np.zeros(mesh.n_cells)
        np.add.at(Cc, mesh.owner, Cf)
        np.add.at(cnt, mesh.owner, 1.0)
        np.add.at(Cc, mesh.neighbour, Cf[:nIF])
        np.add.at(cnt, mesh.neighbour, 1.0)
        Cc /= cnt[:, None]
        yw = cKDTree(Cf[wall["start"]:wall["start"] + wall["size"]]).query(Cc)[0]
        Wg = cases.boundary_layer_state(mesh, yw, U0=U0c if comp else (10.0, 0.0, 0.0), seed=1234, noise=0.001)
        if comp:
            Wg = cases.to_compressible_state(mesh, Wg, thermo)
        W = np.ascontiguousarray(Wg[sol.localStateIndex(mesh.n_cells, mesh.n_faces, compressible=comp)])
        del Wg
    sol.updateOFFields(W)
    t_setup = time.time() - t_setup
    primal = None
    if args.primal_iters > 0 and world == 1:
        # the step before the path (solve_nonlinear): SIMPLE iterations on the device, then the adjoint at that state
        pfail = sol.solvePrimal()
        ps = sol.primalStats
        sol.getOFFields(W)
        primal = {"iterations": ps.iterations, "seconds": ps.seconds, "max_residual": ps.max_residual, "converged": int(ps.converged),
                  "fail": pfail, "p_iterations": ps.p_iterations, "ms_per_iteration": 1e3 * ps.seconds / max(ps.iterations, 1),
                  "CD": sol.calcFunction("CD")}

    # pinned host buffers for the end-to-end (host-buffer) leg
    psi_h = torch.empty(n, dtype=torch.float64).pin_memory()
    y_h = torch.empty(n, dtype=torch.float64).pin_memory()
    psi = psi_h.numpy()
    y = y_h.numpy()
    psi[:] = np.random.default_rng(4321 + rank).uniform(-1, 1, n)
    sol.benchSetVector(psi)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg (value): K products, CUDA events on the solver's stream inside the C ABI
    sol.benchDevice(0, args.warmup)
    sampler = ClockSampler()
    if rank == 0:
        sampler.start()
    barrier()
    ms, launches = sol.benchDevice(0, args.steps)
    barrier()
    per_kernel = {name: sol.benchDevice(which, max(5, args.steps // 5))[0] for name, which in (("RevA", 2), ("RevB", 3), ("RevC", 4))}
    ms_fwd, _ = sol.benchDevice(1, max(5, args.steps // 5))

    # ---- end-to-end leg: the public call with HOST buffers (H2D psi + 3 kernels + D2H y every step)
    for _ in range(3):
        sol.calcdRdWTPsiAD(psi, y)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sol.calcdRdWTPsiAD(psi, y)
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop() if rank == 0 else None

    # max over ranks
    tt = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = float(tt[0]), float(tt[1])

    adjoint = None
    if not args.no_solve and (world == 1 or args.solve_multi):
        try:
            from dafoam_b200.pyDASolvers import Mat, KSP
            t0 = time.perf_counter()
            pc = Mat()
            sol.calcdRdWT(1, pc)
            ksp = KSP()
            sol.createMLRKSPMatrixFree(pc, ksp)
            t_pc = time.perf_counter() - t0
            dFdW = np.zeros(n)
            sol.calcJacTVecProduct("states", "stateVar", W, "CD", "function", np.array([1.0]), dFdW)
            psi_sol = np.zeros(n)
            t0 = time.perf_counter()
            fail = sol.solveLinearEqn(ksp, dFdW, psi_sol)
            t_solve = time.perf_counter() - t0
            st = ksp.stats
            adjoint = {"wall_s": t_pc + t_solve, "pc_s": t_pc, "solve_s": t_solve, "fail": fail, "iterations": st.iterations,
                       "rel_residual": st.final_residual / st.initial_residual if st.initial_residual else None,
                       "n_matvec": st.n_matvec, "gmres_device_s": st.solve_seconds, "method": "GMRES(%d), the reference's KSP" % args.restart}
            # the same system with IDR(s) on the same preconditioner (adjEqnOption.kspType idrs, an extension: short recurrences,
            # no orthogonalisation against the whole basis); reported beside the GMRES number, never instead of it
            if not args.no_idrs:
                try:
                    sol.updateDAOption(dict(adjEqnOption=dict(kspType="idrs", idrS=4, gmresMaxIters=3 * args.max_iters)))
                    psi2 = np.zeros(n)
                    t0 = time.perf_counter()
                    fail2 = sol.solveLinearEqn(ksp, dFdW, psi2)
                    t2 = time.perf_counter() - t0
                    st = ksp.stats
                    dn = float(np.linalg.norm(psi_sol))
                    adjoint["idrs"] = {"method": "IDR(4)", "wall_s": t_pc + t2, "solve_s": t2, "fail": fail2, "operator_applications": st.iterations,
                                       "rel_residual": st.final_residual / st.initial_residual if st.initial_residual else None,
                                       "n_matvec": st.n_matvec, "device_s": st.solve_seconds,
                                       "psi_rel_diff_vs_gmres": float(np.linalg.norm(psi2 - psi_sol)) / dn if dn > 0 else None}
                except Exception as e:
                    adjoint["idrs"] = {"error": str(e)}
                finally:
                    sol.updateDAOption(dict(adjEqnOption=dict(kspType="gmres", gmresMaxIters=args.max_iters)))
        except Exception as e:  # reported, never hidden
            adjoint = {"error": str(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    traffic, traffic_src = None, None
    if comp:
        args.no_cpu_baseline = True  # the CPU port leg times the incompressible oracle; not comparable
    tf = os.path.join(ROOT, "profiles", "r01d_ncu_kernels_dram.json")
    if os.path.exists(tf) and world == 1 and args.cells == 980000 and not comp:
        tj = json.load(open(tf))
        traffic = sum(tj[k]["dram__bytes_read.sum"] + tj[k]["dram__bytes_write.sum"] for k in ("RevA", "RevB", "RevC"))
        traffic_src = "profiles/r01d_ncu_kernels_dram.json (ncu dram__bytes_read+write of RevA+RevB+RevC, same workload)"
    alg = sol.algorithmicBytes(0)
    achieved = alg / (ms_max * 1e-3) / 1e9
    nC_global = sol.getNGlobalCells()
    value = nC_global / (ms_max * 1e-3) / 1e9
    out = {
        "metric": "dRdWTPsi_GCells_per_s", "value": value, "unit": "GCells/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "DASimpleFoam NACA0012 SA %dx%dx1 O-grid, %d cells, %d DOF per GPU; adjoint matvec "
                               "dRdW^T*psi; working set per product ~%.0f MB >> 126 MB L2 (no explicit flush)"
                               % (ni, nj, nC, n, (alg + 60 * 8 * nC) / 1e6),
                   "parallelism": ("domain decomposition (RCB) over %d GPUs, NCCL ghost-cell exchange" % world) if world > 1 else "single GPU", "setup_s": t_setup},
        "gpu_launches": launches,
        "e2e": {"value": nC_global / (e2e_ms_max * 1e-3) / 1e9, "unit": "GCells/s", "ms_per_step": e2e_ms_max,
                "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n,
                "call": "pyDASolvers.calcdRdWTPsiAD(psi_host, y_host) -> dab_drdwt_mat_vec (pinned host buffers)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_product": alg,
                     "kernels_ms": per_kernel, "forward_R_ms": ms_fwd,
                     "note": "one product = RevA+RevB+RevC; achieved = algorithmic bytes of the product / its device time"},
        "adjoint_solve": adjoint,
        "primal_solve": primal,
        "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            out["cpu_baseline"] = cpu_reference(os.cpu_count() or 1, reps=10)
            # config 1 of BASELINE.json (the reference's own ~5k-cell case on ONE CPU rank): the same port on one core
            out["cpu_baseline_1core"] = cpu_reference(1, reps=10)
        except Exception as e:
            out["cpu_baseline"] = {"error": str(e)}
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
