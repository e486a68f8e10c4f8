#!/usr/bin/env python
"""race_mcts leg of bench.py alone (BASELINE config 5 with the MCTS high level: 16,384 2-kart Oval races, planner as the reference runs it):
total time of 200 steps with two planning events.  HK_PLANNER_OVERLAP=0 puts the search on the caller's stream (no overlap with the steps)."""
import sys, time
sys.path.insert(0, '.')
from hierarchicalkarting_b200 import abi, mcts as M, scenarios as S, race as RC
lib = abi.load_library(); abi.check(lib.hk_init(0))
RACES = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
prm = RC.race_params(S.OVAL, high_mode_mcts=True)
RM = RC.Races(S.OVAL, prm)
game = M.Game(S.OVAL, 2, prm.velocityBucketSize)
km, pm = RC.start_grid(S.OVAL, RACES, seed=20260004)
RM.run(km, pm, 0, 100)
kw = dict(mode=0, iterations=512, reuse_cycles=3, apply_delay=45)
for rep in range(3):
    kk, pp = km.copy(), pm.copy()
    pl = RC.Planner(game, RACES, seed=20260006, **kw)
    t0 = time.perf_counter()
    RM.run_planned(kk, pp, pl, 100, 200)
    el = time.perf_counter() - t0
    _, _, ts = pl.state()
    pl.close()
    print(f"race_mcts: {1e3 * el:.2f} ms / 200 steps  {2 * RACES * 200 / el:.4e} agent-steps/s  sections {kk['section'].mean():.3f}  out-of-nodes {(ts == 3).sum()}", flush=True)
