"""Generates tests/golden/*.  The reference (Unity C#) cannot be executed in this image (no .NET), so the LQNG vectors
are produced by the CPU oracle *after* it has been cross-checked against the independent numpy restatement
(oracle/np_lqng.py) and the surveyor's smoke value; the game known-answers are the surveyor's independently computed
float32 numbers of SURVEY.md Appendix D, typed in by hand below (they are NOT produced by the oracle).
Run from the repo root: python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hierarchicalkarting_b200 import scenarios as S  # noqa: E402
from oracle import oracle as O, np_lqng  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def lqng():
    out = {}
    for N, track, seed in ((2, S.OVAL, 4242), (4, S.COMPLEX, 4343)):
        p = S.make_problems(track, 8, N, seed=seed)
        A, B, Q, q, R, x0 = S.assemble_dense(p)
        r = O.lqng_solve_batch(A, B, Q, q, R, x0, 3)
        for b in range(8):      # second opinion before freezing
            u, Ps, als, uall = np_lqng.solve_feedback_lqr(list(A[b]), list(B[b]), list(Q[b]), list(q[b]), list(R[b]),
                                                          list(x0[b].reshape(N, 4)), 3)
            assert np.allclose(Ps, r["P"][b], rtol=1e-11, atol=1e-13) and np.allclose(uall, r["u0"][b], rtol=1e-11, atol=1e-13)
        for k, v in zip(("A", "B", "Q", "q", "R", "x0"), (A, B, Q, q, R, x0)):
            out[f"N{N}_{k}"] = v
        for k in ("u0", "P", "alpha", "traj"):
            out[f"N{N}_{k}_out"] = r[k]
    np.savez_compressed(os.path.join(HERE, "lqng_golden.npz"), **out)


# SURVEY.md Appendix D (surveyor's float32 emulation; Acc 7, Brk 16, Top 15, MaxGs 2, MinGs .5, TWF .001, precision 100)
# section = (insideR, length, width, turnDeg, leftTurn); from = (lane, min_v, max_v, tireAge, timeAtSection); action = (min_v, max_v, lane)
GAME_KAT = {
    "max_speed_for_radius_and_wear": [[17.5, 0.25, 15.0], [1.0, 0.3, 3.8994231], [8.25, 0.0, 12.722618], [20.0, 1.4, 15.0]],
    "apply_action": [
        {"section": [0, 10, 10, 0, 0], "from": [2, 0, 2, 2500, 0], "action": [6, 8, 2], "dist": 10.0, "radius": 0.0, "toc": 1.6026193, "time": 160, "tireLoad": 0.099999994, "tireAge": 2500, "infeasible": 0},
        {"section": [0, 10, 10, 0, 0], "from": [2, 0, 2, 2500, 0], "action": [6, 8, 4], "dist": 12.018505, "radius": 0.0, "toc": 1.7848942, "time": 178, "tireLoad": 0.12018505, "tireAge": 2501, "infeasible": 0},
        {"section": [0, 10, 10, 0, 0], "from": [2, 0, 2, 2500, 0], "action": [14, 15, 2], "dist": 10.0, "radius": 0.0, "toc": -1.0, "time": -100, "infeasible": 1},
        {"section": [0, 10, 10, 0, 0], "from": [3, 12, 14, 2600, 350], "action": [14, 15, 3], "dist": 10.0, "radius": 0.0, "toc": 0.6862351, "time": 418, "tireLoad": 0.099999994, "tireAge": 2600, "infeasible": 0},
        {"section": [15, 10, 10, 45, 1], "from": [1, 10, 12, 2700, 500], "action": [10, 12, 1], "dist": 11.780972, "radius": 15.0, "toc": 0.89492196, "time": 589, "tireLoad": 1.1309735, "tireAge": 2711, "infeasible": 0},
        {"section": [15, 10, 10, 45, 1], "from": [4, 12, 14, 2700, 500], "action": [14, 15, 2], "dist": 15.707964, "radius": 20.0, "toc": 1.066766, "time": 606, "tireLoad": 1.767146, "tireAge": 2717, "infeasible": 0},
        {"section": [1, 10, 10, 90, 0], "from": [3, 6, 8, 3000, 900], "action": [6, 8, 4], "dist": 3.5342917, "radius": 2.25, "toc": 0.5299821, "time": 952, "tireLoad": 1.0053097, "tireAge": 3010, "infeasible": 0},
        {"section": [7, 10, 10, 70, 1], "from": [2, 8, 10, 3000, 900], "action": [6, 8, 3], "dist": 13.133602, "radius": 10.75, "toc": 1.1749473, "time": 1017, "tireLoad": 0.7819075, "tireAge": 3007, "infeasible": 0},
        {"section": [0, 10, 10, 0, 0], "from": [1, 14, 15, 2500, 100], "action": [6, 8, 1], "dist": 10.0, "radius": 0.0, "toc": 0.8011905, "time": 180, "tireLoad": 0.099999994, "tireAge": 2500, "infeasible": 0},
    ],
    # SURVEY.md B.7 closed-form examples of the rollout policy's index distribution
    "policy_pmf": {"3": [0.6827, 0.3146, 0.0026], "8": [0.2923, 0.4471, 0.1998, 0.0521, 0.0079, 0.0007],
                   "20": [0.1192, 0.2281, 0.1995, 0.1595, 0.1167, 0.0781, 0.0478, 0.0267]},
    # Random123 known answers for Philox4x32-10 (key, counter) -> output
    "philox4x32_10": [{"seed": 0, "ctr": [0, 0, 0, 0], "out": [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]},
                      {"seed": 0xffffffffffffffff, "ctr": [0xffffffff] * 4, "out": [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]},
                      {"seed": 0x299f31d0a4093822, "ctr": [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], "out": [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]}],
    # SURVEY.md Appendix D LQNG smoke value (surveyor's numpy): me=(10,2,12,.3) opp=(13,3.5,11,.25) targets (20,5,15,.35)/(20,7.5,15,.3)
    "lqng_smoke_u0": [0.00558, 0.076146],
}

if __name__ == "__main__":
    lqng()
    with open(os.path.join(HERE, "game_kat.json"), "w") as f:
        json.dump(GAME_KAT, f, indent=1)
    print("golden written")
