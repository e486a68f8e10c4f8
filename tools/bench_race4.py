#!/usr/bin/env python
"""race4 leg of bench.py alone (Duos: 8,192 4-kart 2v2 races on Complex): Fixed high level, then the MCTS planner.
HK_RACEN_SPLIT=0 solves every game in the 4-player frame (no 2-kart kernel for the games of one or two players)."""
import sys, time
sys.path.insert(0, '.')
from hierarchicalkarting_b200 import abi, mcts as M, scenarios as S, race as RC
lib = abi.load_library(); abi.check(lib.hk_init(0))
R4 = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
RN = RC.RacesN(S.COMPLEX, RC.race_params(S.COMPLEX), 4)
k4, p4, b4, u4 = RC.start_grid_n(S.COMPLEX, R4, 4, seed=20260007)
RN.plan_fixed(k4, p4)
RN.run_n(k4, p4, b4, u4, 0, 100)
for rep in range(3):
    t0 = time.perf_counter()
    bad = RN.run_n(k4, p4, b4, u4, 100 + 200 * rep, 200)
    el = time.perf_counter() - t0
    npl = RN.recipe_n(k4, p4, b4)["n_players"]
    print(f"race4 fixed: {1e3 * el / 200:.4f} ms/step  {4 * R4 * 200 / el:.4e} agent-steps/s  bad {bad}  N at end {[int((npl == n).sum()) for n in (1, 2, 3, 4)]}", flush=True)
prm = RC.race_params(S.COMPLEX, high_mode_mcts=True)
RNm = RC.RacesN(S.COMPLEX, prm, 4)
game4 = M.Game(S.COMPLEX, 4, prm.velocityBucketSize)
km, pm, bm, um = RC.start_grid_n(S.COMPLEX, R4, 4, seed=20260007)
RNm.run_n(km, pm, bm, um, 0, 100)
for rep in range(2):
    kk, pp, bb, uu = km.copy(), pm.copy(), bm.copy(), um.copy()
    pl = RNm.planner(game4, R4, 256, 20260008, mode=0, reuse_cycles=3, apply_delay=45)
    t0 = time.perf_counter()
    bad = RNm.run_n(kk, pp, bb, uu, 100, 200, planner=pl)
    el = time.perf_counter() - t0
    pl.close()
    print(f"race4 mcts: {1e3 * el:.2f} ms / 200 steps  {4 * R4 * 200 / el:.4e} agent-steps/s  bad {bad}", flush=True)
