"""Worker of the world_size-2 gloo test (launched by tests/test_multirank.py through torch.distributed.run).
The test-only host build of the engine runs the kernels; torch.distributed/gloo carries the ghost-cell
exchanges and the GMRES all-reduces through the library's communication callbacks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dafoam_b200 import cases  # noqa: E402
from dafoam_b200.pyDASolvers import KSP, Mat, pyDASolvers, set_comm_callbacks  # noqa: E402
from tests.common import HOSTSIM, NORM_STATES  # noqa: E402


def main():
    case_dir, kind = sys.argv[1], sys.argv[2]
    cuda = len(sys.argv) > 3 and sys.argv[3] == "cuda"
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = None if cuda else HOSTSIM

    def exchange(peers, sends, recvs):
        reqs = []
        for p, s, r in zip(peers, sends, recvs):
            if r.size:
                reqs.append(dist.irecv(torch.from_numpy(r), src=p))
            if s.size:
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(s)), dst=p))
        for q in reqs:
            q.wait()

    def allreduce(a):
        dist.all_reduce(torch.from_numpy(a))

    if kind.startswith("passage"):
        set_comm_callbacks(exchange, allreduce, HOSTSIM) if not cuda else None
        passage(case_dir, kind, rank, world, cuda, lib)
        dist.barrier()
        dist.destroy_process_group()
        return
    primal_mode = kind in ("channelprimal", "channelcompprimal")
    comp_primal = kind == "channelcompprimal"
    if primal_mode:
        kind = "channel"
    if kind == "nacamrf":
        kind = "naca"  # the same checks on a case with an MRF zone (constant/MRFProperties, cellZones): the zone is cut by the partition
    comp = kind == "nacacomp" or comp_primal
    mesh = cases.naca0012_ogrid(ni=32, nj=16, nk=2) if kind in ("naca", "nacacomp") else cases.channel(nx=12, ny=8, nz=2)
    fn = {"CD": {"type": "force", "source": "patchToFace", "patches": ["wing" if kind in ("naca", "nacacomp") else "walls"],
                 "directionMode": "fixedDirection", "direction": [1.0, 0.0, 0.0], "scale": 1.0}}
    opts = dict(normalizeStates=NORM_STATES, function=fn, adjEqnOption=dict(gmresRelTol=1e-10, gmresMaxIters=500, gmresRestart=250))
    solver_name = "DASimpleFoam -python"
    if comp:
        # DARhoSimpleFoam on the same decomposition: 6 cell states, looser linear-solver tolerance (p ~ 1e5)
        solver_name = "DARhoSimpleFoam -python"
        opts = dict(normalizeStates=dict(U=50.0, p=101325.0, T=300.0, nuTilda=1e-3, phi=1.0), function=fn,
                    adjEqnOption=dict(gmresRelTol=1e-6, gmresMaxIters=900, gmresRestart=900, pcConLevel=3))
    uid = None
    if cuda:
        # product path: NCCL over NVLink; gloo only broadcasts the unique id
        from dafoam_b200.pyDASolvers import nccl_unique_id
        box = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    else:
        set_comm_callbacks(exchange, allreduce, HOSTSIM)
    dev = rank if cuda else 0
    serial = pyDASolvers(solver_name, opts, caseDir=case_dir, device=dev, _lib_path=lib)
    par = pyDASolvers(solver_name, opts, caseDir=case_dir, device=dev, rank=rank, nRanks=world, ncclUniqueId=uid, _lib_path=lib)
    nCg, nFg = mesh.n_cells, mesh.n_faces
    assert par.getNGlobalCells() == nCg and serial.getNLocalCells() == nCg
    idx = par.localStateIndex(nCg, nFg, compressible=comp)
    owned = np.concatenate([np.ones((6 if comp else 5) * par.getNLocalCells(), dtype=bool), par.getLocalToGlobal("faceOwned").astype(bool)])
    n_cells_total = torch.tensor([par.getNLocalCells()])
    dist.all_reduce(n_cells_total)
    assert int(n_cells_total) == nCg

    if primal_mode:
        # solvePrimal on two ranks (ghost exchanges per sweep / CG iteration, all-reduced residual norms, per-rank coarse
        # spaces) against the single-rank SIMPLE: same fixed point
        o2 = dict(opts, primalMinResTol=1e-11, primalMaxIters=4000)
        serial.updateDAOption(o2)
        par.updateDAOption(o2)
        fs, fp = serial.solvePrimal(), par.solvePrimal()
        assert fs == 0 and fp == 0, (fs, fp, serial.primalStats.max_residual, par.primalStats.max_residual)
        Ws, Wp = np.zeros(serial.getNLocalAdjointStates()), np.zeros(idx.size)
        serial.getOFFields(Ws)
        par.getOFFields(Wp)
        nCl = par.getNLocalCells()
        ns_ = 6 if comp else 5
        errs = []
        for a, b in [(0, 3 * nCl)] + [(k * nCl, (k + 1) * nCl) for k in range(3, ns_)]:
            ref = Ws[idx][a:b]
            errs.append(np.linalg.norm(Wp[a:b] - ref) / np.linalg.norm(ref))
        fo = owned[ns_ * nCl:]
        errs.append(np.linalg.norm(Wp[ns_ * nCl:][fo] - Ws[idx][ns_ * nCl:][fo]) / np.linalg.norm(Ws[idx][ns_ * nCl:][fo]))
        assert max(errs) < 1e-7, errs
        print("rank %d ok: primal iterations serial %d, 2 ranks %d, state difference %.1e" % (rank, serial.primalStats.iterations,
                                                                                               par.primalStats.iterations, max(errs)), flush=True)
        dist.barrier()
        dist.destroy_process_group()
        return
    y = np.zeros(nCg)
    serial.getOFField("yWall", "scalar", y)
    Wg = cases.boundary_layer_state(mesh, y, noise=0.01) if kind == "naca" else None
    if comp:
        from oracle.pyoracle import synthetic_state  # a state generator only (no oracle evaluation in this test)
        from tests.common import rel_err  # noqa: F401
        import numpy as _np
        th = cases.default_thermo()
        Cc = _np.zeros(3 * nCg)
        serial.getOFField("C", "vector", Cc) if False else None
        # cell centres and face areas from the mesh points (quad faces)
        Sf, Cf = cases.quad_face_geometry(mesh)
        cc = _np.zeros((nCg, 3)); cnt = _np.zeros(nCg)
        _np.add.at(cc, mesh.owner, Cf); _np.add.at(cnt, mesh.owner, 1.0)
        _np.add.at(cc, mesh.neighbour, Cf[:mesh.n_internal_faces]); _np.add.at(cnt, mesh.neighbour, 1.0)
        cc /= cnt[:, None]
        Wg = synthetic_state(mesh, cc.ravel(), Sf.ravel(), U0=(50.0, 2.0, 0.0), thermo=th)
    if Wg is None:
        Wg = np.zeros(serial.getNLocalAdjointStates())
        serial.getOFFields(Wg)
        Wg *= 1.0 + 0.01 * np.random.default_rng(1).uniform(-1, 1, Wg.size)
    serial.updateOFFields(Wg)
    par.updateOFFields(np.ascontiguousarray(Wg[idx]))

    def check(name, loc, glob, tol):
        err = np.abs(loc[owned] - glob[idx][owned]).max() / max(np.abs(glob).max(), 1e-300)
        assert np.all(loc[~owned] == 0.0), name + ": foreign slots must be structural zeros"
        assert err < tol, (name, err)
        return err

    Rg, Rl = np.zeros(Wg.size), np.zeros(idx.size)
    serial.getResiduals(Rg)
    par.getResiduals(Rl)
    e1 = check("residual", Rl, Rg, 1e-12)
    psi = np.random.default_rng(4321).uniform(-1, 1, Wg.size)
    yg, yl = np.zeros(Wg.size), np.zeros(idx.size)
    serial.calcdRdWTPsiAD(psi, yg)
    pl = np.ascontiguousarray(psi[idx])
    pl[~owned] = 0.0
    par.calcdRdWTPsiAD(pl, yl)
    e2 = check("dRdWTPsi", yl, yg, 1e-12)
    Fs, Fp = serial.calcFunction("CD"), par.calcFunction("CD")
    assert abs(Fs - Fp) <= 1e-12 * abs(Fs)
    dg, dl = np.zeros(Wg.size), np.zeros(idx.size)
    serial.calcJacTVecProduct("s", "stateVar", Wg, "CD", "function", np.array([1.0]), dg)
    par.calcJacTVecProduct("s", "stateVar", np.ascontiguousarray(Wg[idx]), "CD", "function", np.array([1.0]), dl)
    e3 = check("dFdW", dl, dg, 1e-12)
    # adjoint solve: block-Jacobi ILU over the two ranks, GMRES dot products all-reduced
    ks, kp = KSP(), KSP()
    ms, mp_ = Mat(), Mat()
    serial.calcdRdWT(1, ms)
    par.calcdRdWT(1, mp_)
    xs, xp = np.zeros(Wg.size), np.zeros(idx.size)
    fs = serial.solveLinearEqn(ks, dg, xs)
    fp = par.solveLinearEqn(kp, dl, xp)
    assert fs == 0 and fp == 0, (fs, fp, ks.stats.iterations, kp.stats.iterations)
    e4 = np.linalg.norm(xp[owned] - xs[idx][owned]) / np.linalg.norm(xs)
    assert e4 < (1e-3 if comp else 1e-6), e4  # compressible: both solves stop at rtol 1e-6 of an ill-scaled system
    if not comp:
        # two-level preconditioner (restriction all-reduced, coarse GEMV on every rank) with IDR(4) across the two ranks
        par.updateDAOption(dict(adjEqnOption=dict(coarseAggregates=12, kspType="idrs", idrS=4, gmresMaxIters=3000)))
        par.calcdRdWT(1, mp_)
        xq = np.zeros(idx.size)
        fq = par.solveLinearEqn(kp, dl, xq)
        assert fq == 0, (fq, kp.stats.iterations)
        e5 = np.linalg.norm(xq[owned] - xs[idx][owned]) / np.linalg.norm(xs)
        assert e5 < 1e-5, e5
    print("rank %d ok: residual %.1e dRdWTPsi %.1e dFdW %.1e psi %.1e (its serial %d, 2 ranks %d)"
          % (rank, e1, e2, e3, e4, ks.stats.iterations, kp.stats.iterations), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def passage(case_dir, kind, rank, world, cuda, lib):
    """Annular passage with cyclic sides cut into two sub-meshes (the cut crosses the coupled patches, so periodic images travel
    between the ranks and are rotated on the way) against the same passage on one rank, where the images are local copies."""
    comp = "turbo" in kind
    fn = {"CD": {"type": "force", "source": "patchToFace", "patches": ["hub"], "directionMode": "fixedDirection", "direction": [0.0, 0.0, 1.0],
                 "scale": 1.0}}
    if comp:
        name = "DATurboFoam -python"
        opts = dict(normalizeStates=dict(U=50.0, p=101325.0, T=300.0, nuTilda=1e-3, phi=1.0), function=fn,
                    adjEqnOption=dict(gmresRelTol=1e-8, gmresMaxIters=900, gmresRestart=900, pcConLevel=3))
    else:
        name = "DASimpleFoam -python"
        opts = dict(normalizeStates=NORM_STATES, function=fn, adjEqnOption=dict(gmresRelTol=1e-10, gmresMaxIters=600, gmresRestart=300))
    uid = None
    if cuda:
        from dafoam_b200.pyDASolvers import nccl_unique_id
        box = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    dev = rank if cuda else 0
    one = pyDASolvers(name, opts, caseDir=case_dir, device=dev, _lib_path=lib)
    two = pyDASolvers(name, opts, caseDir=case_dir, device=dev, rank=rank, nRanks=world, ncclUniqueId=uid, _lib_path=lib)
    nCg = one.getNGlobalCells()
    assert two.getNGlobalCells() == nCg and one.getNLocalCells() == nCg
    fo1 = one.getLocalToGlobal("faceOwned").astype(bool)
    nFg = int(one.getLocalToGlobal("faces").max()) + 1
    ns = 6 if comp else 5
    n = ns * nCg + nFg

    def maps(sol):
        idx = sol.localStateIndex(nCg, nFg, compressible=comp)
        owned = np.concatenate([np.ones(ns * sol.getNLocalCells(), dtype=bool), sol.getLocalToGlobal("faceOwned").astype(bool)])
        return idx, owned

    i1, o1 = maps(one)
    i2, o2 = maps(two)
    assert fo1.sum() == nFg

    def merged(idx, owned, v):
        out = np.zeros(n)
        out[idx[owned]] = v[owned]
        return out

    if kind.endswith("primal"):
        o = dict(opts, primalMinResTol=1e-10, primalMaxIters=3000)
        one.updateDAOption(o)
        two.updateDAOption(o)
        f1, f2 = one.solvePrimal(), two.solvePrimal()
        assert f1 == 0 and f2 == 0, (f1, f2, one.primalStats.max_residual, two.primalStats.max_residual)
        W1, W2 = np.zeros(i1.size), np.zeros(i2.size)
        one.getOFFields(W1)
        two.getOFFields(W2)
        g1 = merged(i1, o1, W1)
        err = np.abs(W2[o2] - g1[i2][o2]).max() / np.abs(g1).max()
        assert err < 1e-7, err
        # the converged flow must repeat across the coupled patches: something flows through them
        phi_c = g1[ns * nCg:][one.getLocalToGlobal("faces")[~fo1]]
        assert np.abs(phi_c).max() > 0.0
        print("rank %d ok: primal iterations one rank %d, two ranks %d, state difference %.1e" % (rank, one.primalStats.iterations,
                                                                                                two.primalStats.iterations, err), flush=True)
        return
    W1 = np.zeros(i1.size)
    one.getOFFields(W1)
    rng = np.random.default_rng(5)
    Wg = merged(i1, o1, W1) * (1.0 + 0.01 * rng.uniform(-1, 1, n))
    Wg[:3 * nCg] += 0.3 * rng.uniform(-1, 1, 3 * nCg)
    one.updateOFFields(np.ascontiguousarray(Wg[i1]))
    two.updateOFFields(np.ascontiguousarray(Wg[i2]))

    def both(f, tol, what):
        a, b = np.zeros(i1.size), np.zeros(i2.size)
        f(one, a, i1, o1)
        f(two, b, i2, o2)
        g = merged(i1, o1, a)
        assert np.all(b[~o2] == 0.0), what + ": foreign slots must be structural zeros"
        err = np.abs(b[o2] - g[i2][o2]).max() / max(np.abs(g).max(), 1e-300)
        assert err < tol, (what, err)
        return err, g, b

    e1, _, _ = both(lambda s_, out, idx, own: s_.getResiduals(out), 1e-12, "residual")
    psi = rng.uniform(-1, 1, n)

    def prod(s_, out, idx, own):
        x = np.ascontiguousarray(psi[idx])
        x[~own] = 0.0
        s_.calcdRdWTPsiAD(x, out)

    e2, _, _ = both(prod, 1e-12, "dRdWTPsi")
    F1, F2 = one.calcFunction("CD"), two.calcFunction("CD")
    assert abs(F1 - F2) <= 1e-12 * abs(F1), (F1, F2)
    e3, dg, dl = both(lambda s_, out, idx, own: s_.calcJacTVecProduct("s", "stateVar", np.ascontiguousarray(Wg[idx]), "CD", "function",
                                                                         np.array([1.0]), out), 1e-12, "dFdW")
    k1, k2, m1, m2 = KSP(), KSP(), Mat(), Mat()
    one.calcdRdWT(1, m1)
    two.calcdRdWT(1, m2)
    x1, x2 = np.zeros(i1.size), np.zeros(i2.size)
    r1 = np.ascontiguousarray(dg[i1])
    r1[~o1] = 0.0
    f1 = one.solveLinearEqn(k1, r1, x1)
    f2 = two.solveLinearEqn(k2, dl, x2)
    assert f1 == 0 and f2 == 0, (f1, f2, k1.stats.iterations, k2.stats.iterations)
    g = merged(i1, o1, x1)
    e4 = np.linalg.norm(x2[o2] - g[i2][o2]) / np.linalg.norm(g)
    assert e4 < (1e-3 if comp else 1e-6), e4
    print("rank %d ok: residual %.1e dRdWTPsi %.1e dFdW %.1e psi %.1e (its one rank %d, two ranks %d)"
          % (rank, e1, e2, e3, e4, k1.stats.iterations, k2.stats.iterations), flush=True)


if __name__ == "__main__":
    main()
