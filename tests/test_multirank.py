"""world_size-2 domain-decomposed run on CPUs (gloo): partitioner, ghost-cell exchange plan, de-duplicated
cut-face DOFs and the all-reduced GMRES, checked against the single-rank engine on the same case."""
import os
import subprocess
import sys
import tempfile

import pytest

from dafoam_b200 import cases
from tests.common import ROOT


def _write_passage(d, kind):
    if "axial" in kind:
        # a long passage: the RCB cut is axial, so each rank holds parts of BOTH coupled patches (images copied inside the rank) next
        # to an ordinary cut (ghosts from the other rank)
        mesh = cases.annular_passage(nr=4, nt=4, nz=16, lz=0.9, n_sectors=7)
    else:
        mesh = cases.annular_passage(nr=5, nt=6, nz=8, n_sectors=7)
    bcs = cases.default_bcs_passage(Uin=(0.0, 0.0, 60.0 if "turbo" in kind else 10.0))
    kw = {}
    if "turbo" in kind:
        bcs = cases.compressible_bcs(bcs)
        kw = dict(thermo=cases.default_thermo(energy="sensibleEnthalpy"),
                  mrf=dict(cellZone="rotor", cells=list(range(mesh.n_cells)), origin=(0.0, 0.0, 0.0), axis=(0.0, 0.0, 1.0), omega=200.0,
                           nonRotatingPatches=["inlet", "outlet", "shroud"]))
    cases.write_case(d, mesh, bcs, **kw)


def _run(kind, extra, port):
    d = tempfile.mkdtemp(prefix="dab_mp_")
    if kind.startswith("passage"):
        _write_passage(d, kind)
    elif kind == "naca":
        cases.write_case(d, cases.naca0012_ogrid(ni=32, nj=16, nk=2), cases.default_bcs_naca())
    else:
        cases.write_case(d, cases.channel(nx=12, ny=8, nz=2), cases.default_bcs_channel())
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mp_worker.py"), d, kind] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, OMP_NUM_THREADS="1"), cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok: ") == 2, r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("kind,port", [("naca", 29735), ("channelprimal", 29741), ("passageturbo", 29757)])
def test_two_gpus_match_one_gpu_nccl(kind, port):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(kind, ["cuda"], port)


@pytest.mark.parametrize("kind", ["naca", "channel", "nacacomp", "channelprimal", "channelcompprimal", "nacamrf"])
def test_two_ranks_match_one_rank(kind):
    d = tempfile.mkdtemp(prefix="dab_mp_")
    if kind == "naca":
        cases.write_case(d, cases.naca0012_ogrid(ni=32, nj=16, nk=2), cases.default_bcs_naca())
    elif kind == "nacamrf":
        from tests.common import mrf_zone
        mesh = cases.naca0012_ogrid(ni=32, nj=16, nk=2)
        cases.write_case(d, mesh, cases.default_bcs_naca(), mrf=mrf_zone(mesh, omega=2.0))
    elif kind == "nacacomp":
        cases.write_case(d, cases.naca0012_ogrid(ni=32, nj=16, nk=2), cases.compressible_bcs(cases.default_bcs_naca(U0=(50.0, 2.0, 0.0))),
                         thermo=cases.default_thermo())
    elif kind == "channelcompprimal":
        cases.write_case(d, cases.channel(nx=12, ny=8, nz=2), cases.compressible_bcs(cases.default_bcs_channel(U0=(60.0, 0.0, 0.0))),
                         thermo=cases.default_thermo())
    else:
        cases.write_case(d, cases.channel(nx=12, ny=8, nz=2), cases.default_bcs_channel())
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", {"naca": "29731", "channel": "29733", "nacacomp": "29737", "channelprimal": "29739", "channelcompprimal": "29743", "nacamrf": "29745"}[kind], os.path.join(ROOT, "tests", "mp_worker.py"), d, kind]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok: ") == 2, r.stdout


@pytest.mark.parametrize("kind,port,nproc", [("passage", 29751, 2), ("passageturbo", 29753, 2), ("passageprimal", 29755, 2),
                                             ("passageaxial", 29759, 3), ("passageturbo", 29761, 4)])
def test_cyclic_passage_on_two_ranks(kind, port, nproc):
    """cyclic patch pairs across partition cuts (tests/mp_worker.py passage): images rotated inside the pack kernels; 2, 3 and 4 ranks
    (BASELINE config 5 runs on 8)"""
    d = tempfile.mkdtemp(prefix="dab_mp_")
    _write_passage(d, kind)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % nproc, "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mp_worker.py"), d, kind]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, OMP_NUM_THREADS="1"), cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok: ") == nproc, r.stdout
