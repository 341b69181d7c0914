"""Committed golden vectors (tests/golden/, generated from the oracle by make_golden.py): the oracle must not drift,
and the engine must reproduce them.  The `reference` test is the hook that replays the reference's own known answers
(tests/golden/reference_constants.json) once the ConvergentChannel fixture is supplied -- parity with the reference is
unpinned until then."""
import json
import os
import tempfile

import numpy as np
import pytest

from dafoam_b200 import cases
from dafoam_b200.pyDASolvers import pyDASolvers
from oracle.pyoracle import Oracle
from tests.common import HOSTSIM, NORM_STATES, ROOT, rel_err
from tests.golden.make_golden import golden_spec, oracle_of

GOLD = os.path.join(ROOT, "tests", "golden")


KINDS = ["naca", "channel", "nacafv3", "nacacomp", "nacamrf", "nacaturbo", "nacatransonic"]


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_reproduces_golden(kind):
    spec = golden_spec(kind)
    g = np.load(os.path.join(GOLD, spec["name"] + ".npz"))
    orc = oracle_of(spec)
    assert rel_err(orc.residual(g["W"]), g["R"]) < 1e-13
    assert rel_err(orc.residual(g["W"], 1), g["Rpc"]) < 1e-13
    orc.record(g["W"])
    assert rel_err(orc.jtvec(g["psi"]), g["jt"]) < 1e-13
    assert rel_err(orc.jtvec(np.full(orc.ndof, 1e-3)), g["jt_const"]) < 1e-13


def engine_vs_golden(kind, lib_path, tol=1e-11):
    spec = golden_spec(kind)
    mesh, bcs, fpatch = spec["mesh"], spec["bcs"], spec["fpatch"]
    g = np.load(os.path.join(GOLD, spec["name"] + ".npz"))
    d = tempfile.mkdtemp(prefix="dab_gold_")
    kw = dict(ras_model=spec["ras"])
    if spec["thermo"] is not None:
        kw["thermo"] = spec["thermo"]
    if spec.get("mrf"):
        kw["mrf"] = spec["mrf"]
    if spec.get("transonic"):
        kw["div_phid_p"] = spec["transonic"]["scheme"]
    cases.write_case(d, mesh, bcs, **kw)
    fn = {"F": {"type": "force", "source": "patchToFace", "patches": [fpatch], "directionMode": "fixedDirection",
                "direction": [1.0, 0.0, 0.0], "scale": 1.0}}
    sol = pyDASolvers(spec["solver"] + " -python", dict(normalizeStates=spec["ns"], normalizeResiduals=list(spec["nres"]), function=fn),
                      caseDir=d, _lib_path=lib_path)
    W = np.ascontiguousarray(g["W"])
    n = W.size
    sol.updateOFFields(W)
    R, Rpc, y, yc, dF = (np.zeros(n) for _ in range(5))
    sol.getResiduals(R, 0)
    sol.getResiduals(Rpc, 1)
    sol.calcdRdWTPsiAD(np.ascontiguousarray(g["psi"]), y)
    # the reference's own check: norm of dRdW^T * (0.001 * ones) through calcJacTVecProduct (runUnitTests_DATurbModel.py:56-66)
    sol.calcJacTVecProduct("stateName", "stateVar", W, "residualName", "residual", np.full(n, 1e-3), yc)
    sol.calcJacTVecProduct("stateName", "stateVar", W, "F", "function", np.array([1.0]), dF)
    assert rel_err(R, g["R"]) < tol and rel_err(Rpc, g["Rpc"]) < tol
    assert rel_err(y, g["jt"]) < tol and rel_err(yc, g["jt_const"]) < tol
    assert abs(np.linalg.norm(yc) - float(g["norm_jt_const"])) <= 1e-10 * float(g["norm_jt_const"])
    assert abs(sol.calcFunction("F") - float(g["F"])) <= max(1e-11, tol) * abs(float(g["F"]))
    assert rel_err(dF, g["dFdW"]) < tol


@pytest.mark.parametrize("kind", KINDS)
def test_engine_host_build_reproduces_golden(kind):
    engine_vs_golden(kind, HOSTSIM)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_engine_cuda_reproduces_golden(kind):
    # FMA contraction on the GPU + 8 orders of magnitude of cell volumes on the O-grid: 5e-10 observed
    engine_vs_golden(kind, None, tol=5e-9)


def test_reference_known_answers_hook():
    consts = json.load(open(os.path.join(GOLD, "reference_constants.json")))
    assert abs(consts["DASimpleFoam"]["sa"][2] - 1732.238877108044) < 1e-9  # the number this path must hit
    fixture = os.path.join(ROOT, "tests", "fixtures", "ConvergentChannel")
    if not os.path.isdir(fixture):
        pytest.skip("ConvergentChannel case not available (downloaded by the reference at test time): parity unpinned")
    states = np.load(os.path.join(fixture, "states.npy"))
    sol = pyDASolvers("DASimpleFoam -python", {}, caseDir=fixture, _lib_path=HOSTSIM)
    y = np.zeros(states.size)
    sol.calcJacTVecProduct("stateName", "stateVar", states, "residualName", "residual", np.full(states.size, 1e-3), y)
    ref = consts["DASimpleFoam"]["sa"][2]
    assert abs(np.linalg.norm(y) - ref) <= 1e-8 * ref
