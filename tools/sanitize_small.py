#!/usr/bin/env python
"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck): a few hundred LQNG problems
of each size and output form, the compact entry, rollouts, one short race."""
import sys
sys.path.insert(0, '.')
import numpy as np
from hierarchicalkarting_b200 import abi, lqr, mcts as M, tracks, scenarios as S, race as RC
lib = abi.load_library(); abi.check(lib.hk_init(0))
for N, track in ((1, S.OVAL), (2, S.OVAL), (3, S.COMPLEX), (4, S.COMPLEX)):
    A, B, Q, q, R, x0 = S.assemble_dense(S.make_problems(track, 130, N, seed=3))
    for full in (False, True):
        out = lqr.solve_batch(A, B, Q, q, R, x0, 3, full=full)
        assert np.all(np.isfinite(out["u0"]))
# time-varying operands (2-kart: whole horizon staged by TMA; 3/4-kart: DMMA kernel; 1-kart: general kernel)
rng = np.random.default_rng(1)
for N, track in ((1, S.OVAL), (2, S.OVAL), (3, S.COMPLEX), (4, S.COMPLEX)):
    A, B, Q, q, R, x0 = S.assemble_dense(S.make_problems(track, 70, N, seed=5))
    tv = lambda a, s_: np.ascontiguousarray(np.repeat(a[:, None], 4, axis=1) * (1.0 + s_ * rng.standard_normal((70, 4) + (1,) * (a.ndim - 1))))
    for full in (False, True):
        out = lqr.solve_batch(tv(A, 0.01), tv(B, 0.05), tv(Q, 0.05), tv(q, 0.05), tv(R, 0.05), x0, 3, time_varying=True, full=full)
        assert np.all(np.isfinite(out["u0"]))
prob = S.make_problems(S.OVAL, 20000, 2, seed=4)
keys = ("x0", "target", "tw", "cw", "aw", "otgt", "otw")
arrs = [np.ascontiguousarray(prob[k], dtype=np.float64) for k in keys]
u0 = np.zeros((20000, 4)); st = np.zeros(20000, dtype=np.int32)
abi.check(lib.hk_lqng_assemble_solve_batch(20000, 2, 3, float(prob["dt"]), *[abi.dptr(a) for a in arrs], abi.dptr(u0), abi.iptr(st)))
G = M.Game(tracks.COMPLEX, 2, 2)
leaf = tracks.root_state(tracks.COMPLEX, 3, [2, 3], teams=[0, 1], tire_age=2500, times=[0, 80])
G.rollouts(leaf, 20000, seed=1)
roots = [tracks.root_state(tracks.COMPLEX, s0, [2, 3], teams=[0, 1], tire_age=2500, times=[0, 40]) for s0 in (0, 5, 9, 30)]
out = G.search_batch(roots, 5, 8, seed=3)
assert out["n_nodes"].min() > 1
RS = RC.Races(S.OVAL, RC.race_params(S.OVAL))
karts, plans = RC.start_grid(S.OVAL, 64, seed=1)
RS.run(karts, plans, 0, 5)
prm_m = RC.race_params(S.OVAL, high_mode_mcts=True)
RM = RC.Races(S.OVAL, prm_m)
gm = M.Game(S.OVAL, 2, prm_m.velocityBucketSize)
km, pm = RC.start_grid(S.OVAL, 16, seed=2)
RM.run_mcts(km, pm, gm, 6, 4, 5, 0, 102)
print("sanitize_small ok")
