"""Golden vectors of the cyclic rotor passage (tests/test_cyclic.py), generated from the ORACLE on the closed ring of passages:
    python tests/golden/make_golden_cyclic.py
Stored in the engine's merged numbering of the single passage (cases.merged_face_order): the state W, psi, and the passage-0 rows of
R(W) and of [dR/dW]^T psi that the oracle computes on the ring of 5 passages WITHOUT any cyclic patch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests.test_cyclic import Pair  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SPEC = dict(turbulent=True, divU="linearUpwindV", solver="DATurboFoam", energy="sensibleEnthalpy", mrf_omega=300.0)


def main():
    P = Pair(SPEC["turbulent"], SPEC["divU"], solver=SPEC["solver"], energy=SPEC["energy"], mrf_omega=SPEC["mrf_omega"])
    W = P.state()
    Wr = P.to_ring(W)
    R = P.from_ring(P.orc.residual(Wr, 0))
    P.orc.record(Wr)
    psi = np.random.default_rng(2024).uniform(-1, 1, P.n_sec())
    y = P.from_ring(P.orc.jtvec(P.to_ring(psi)))
    np.savez_compressed(os.path.join(HERE, "passage_turbo_cyclic_4x4x6.npz"), W=W, psi=psi, R=R, y=y)
    print("wrote passage_turbo_cyclic_4x4x6.npz", W.size, np.linalg.norm(R), np.linalg.norm(y))


if __name__ == "__main__":
    main()
