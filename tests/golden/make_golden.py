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


NS_COMP = dict(U=50.0, p=101325.0, T=300.0, nuTilda=1e-3, phi=1.0)
NRES_COMP = ("URes", "pRes", "TRes", "nuTildaRes", "phiRes")


def golden_spec(kind):
    """Everything needed to rebuild the case of a golden file: the two original kinds plus the SA-fv3 and DARhoSimpleFoam ones."""
    if kind in ("naca", "channel"):
        mesh, bcs, fpatch, ipatch, name = golden_case(kind)
        return dict(mesh=mesh, bcs=bcs, fpatch=fpatch, name=name, solver="DASimpleFoam", ras="SpalartAllmaras", thermo=None, ns=NORM_STATES,
                    nres=("URes", "pRes", "nuTildaRes", "phiRes"))
    mesh = cases.naca0012_ogrid(ni=24, nj=12, nk=1)
    if kind == "nacafv3":
        return dict(mesh=mesh, bcs=cases.default_bcs_naca(), fpatch="wing", name="naca_safv3_24x12", solver="DASimpleFoam",
                    ras="SpalartAllmarasFv3", thermo=None, ns=NORM_STATES, nres=("URes", "pRes", "nuTildaRes", "phiRes"))
    if kind == "nacacomp":
        return dict(mesh=mesh, bcs=cases.compressible_bcs(cases.default_bcs_naca(U0=(50.0, 2.0, 0.0))), fpatch="wing", name="naca_rhosimple_24x12",
                    solver="DARhoSimpleFoam", ras="SpalartAllmaras", thermo=cases.default_thermo(), ns=NS_COMP, nres=NRES_COMP)
    from tests.common import mrf_zone
    if kind == "nacamrf":
        # DASimpleFoam with an MRF zone around the (rotating) aerofoil wall
        return dict(mesh=mesh, bcs=cases.default_bcs_naca(), fpatch="wing", name="naca_mrf_24x12", solver="DASimpleFoam", ras="SpalartAllmaras",
                    thermo=None, ns=NORM_STATES, nres=("URes", "pRes", "nuTildaRes", "phiRes"), mrf=mrf_zone(mesh, omega=25.0))
    if kind == "nacaturbo":
        # DATurboFoam, enthalpy form (viscous work + p(U - URel) in the energy row), sutherland transport, MRF zone
        return dict(mesh=mesh, bcs=cases.compressible_bcs(cases.default_bcs_naca(U0=(50.0, 2.0, 0.0))), fpatch="wing", name="naca_turbo_h_24x12",
                    solver="DATurboFoam", ras="SpalartAllmaras", thermo=cases.default_thermo(energy="sensibleEnthalpy", transport="sutherland"),
                    ns=NS_COMP, nres=NRES_COMP, mrf=mrf_zone(mesh, omega=25.0))
    if kind == "nacatransonic":
        # DARhoSimpleCFoam (transonic pressure equation, div(phid,p) limitedLinear 1.0), freestream Mach 0.66
        return dict(mesh=mesh, bcs=cases.compressible_bcs(cases.default_bcs_naca(U0=(230.0, 8.0, 0.0))), fpatch="wing",
                    name="naca_rhosimplec_24x12", solver="DARhoSimpleCFoam", ras="SpalartAllmaras", thermo=cases.default_thermo(),
                    ns=dict(NS_COMP, U=230.0), nres=NRES_COMP, transonic=dict(scheme="Gauss limitedLinear 1.0", code=4, k=1.0), U0=(230.0, 8.0, 0.0))
    raise ValueError(kind)


def oracle_of(spec):
    orc = Oracle(spec["mesh"], spec["bcs"], normalizeStates=spec["ns"], normalizeResiduals=spec["nres"], rasModel=spec["ras"], thermo=spec["thermo"])
    if spec["solver"] == "DATurboFoam":
        orc.set_turbo(True)
    if spec.get("mrf"):
        orc.set_mrf(spec["mesh"], spec["mrf"])
    if spec.get("transonic"):
        orc.set_transonic(True, spec["transonic"]["code"], spec["transonic"]["k"], -1)
    return orc


def state_for(kind, mesh, orc):
    if kind == "nacatransonic":
        from oracle.pyoracle import synthetic_state
        return synthetic_state(mesh, orc.geometry("C"), orc.geometry("Sf"), U0=(230.0, 8.0, 0.0), thermo=cases.default_thermo(), noise=0.01)
    if kind in ("nacacomp", "nacaturbo"):
        from oracle.pyoracle import synthetic_state
        return synthetic_state(mesh, orc.geometry("C"), orc.geometry("Sf"), U0=(50.0, 2.0, 0.0), thermo=cases.default_thermo(), noise=0.01)
    if kind in ("nacafv3", "nacamrf"):
        return cases.boundary_layer_state(mesh, orc.geometry("yWall"))
    if kind == "naca":
        return cases.boundary_layer_state(mesh, orc.geometry("yWall"))
    from oracle.pyoracle import synthetic_state
    return synthetic_state(mesh, orc.geometry("C"), orc.geometry("Sf"), U0=(10.0, 0.5, 0.0), noise=0.01)


def compute(kind):
    spec = golden_spec(kind)
    mesh, name = spec["mesh"], spec["name"]
    orc = oracle_of(spec)
    W = state_for(kind, mesh, orc)
    names = [p["name"] for p in mesh.patches]
    fi = names.index(spec["fpatch"])
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
    only = sys.argv[1:] or ("naca", "channel", "nacafv3", "nacacomp", "nacamrf", "nacaturbo", "nacatransonic")
    for kind in only:
        name, data = compute(kind)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
        print(name, {k: (float(np.linalg.norm(v)) if np.ndim(v) else float(v)) for k, v in data.items()})
