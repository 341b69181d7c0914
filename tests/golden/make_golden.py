"""Generate the golden fixtures of tests/golden/ from the oracle.  Run from the repo root:
    python tests/golden/make_golden.py
(Committed together with its outputs so that the vectors can be regenerated and audited.)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from dafoam_b200 import cases  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

NORM_STATES = dict(U=10.0, p=50.0, nuTilda=1e-3, phi=1.0)
HERE = os.path.dirname(os.path.abspath(__file__))


def golden_case(kind):
    if kind == "naca":
        mesh, bcs = cases.naca0012_ogrid(ni=24, nj=12, nk=1), cases.default_bcs_naca()
        fpatch, ipatch, name = "wing", "inout", "naca_sa_24x12"
    else:
        mesh, bcs = cases.channel(nx=16, ny=10, nz=1), cases.default_bcs_channel()
        fpatch, ipatch, name = "walls", "inlet", "channel_sa_16x10"
    return mesh, bcs, fpatch, ipatch, name


def state_for(kind, mesh, orc):
    if kind == "naca":
        return cases.boundary_layer_state(mesh, orc.geometry("yWall"))
    from oracle.pyoracle import synthetic_state
    return synthetic_state(mesh, orc.geometry("C"), orc.geometry("Sf"), U0=(10.0, 0.5, 0.0), noise=0.01)


def compute(kind):
    mesh, bcs, fpatch, ipatch, name = golden_case(kind)
    orc = Oracle(mesh, bcs, normalizeStates=NORM_STATES)
    W = state_for(kind, mesh, orc)
    names = [p["name"] for p in mesh.patches]
    fi, ii = names.index(fpatch), names.index(ipatch)
    R = orc.residual(W)
    Rpc = orc.residual(W, 1)
    orc.record(W)
    psi = np.random.default_rng(4321).uniform(-1, 1, orc.ndof)
    jt = orc.jtvec(psi)
    jt_const = orc.jtvec(np.full(orc.ndof, 1e-3))
    d = [1.0, 0.0, 0.0]
    F = orc.force(W, fi, d, 1.0)
    dFdW = orc.dforce_dw(W, fi, d, 1.0)
    return name, dict(W=W, R=R, Rpc=Rpc, psi=psi, jt=jt, jt_const=jt_const, F=F, dFdW=dFdW,
                      norm_jt_const=np.linalg.norm(jt_const))


if __name__ == "__main__":
    for kind in ("naca", "channel"):
        name, data = compute(kind)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
        print(name, {k: (float(np.linalg.norm(v)) if np.ndim(v) else float(v)) for k, v in data.items()})
