#!/usr/bin/env python
"""hk_mcts_forest_search throughput (the faithful sequential search, one thread per tree): n roots (Oval race states after 100 steps)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from hierarchicalkarting_b200 import abi, mcts as M, race as R, scenarios as S
lib = abi.load_library(); abi.check(lib.hk_init(0))
track = S.OVAL
prm = R.race_params(track, high_mode_mcts=True)
G = R.Races(track, prm)
game = M.Game(track, 2, prm.velocityBucketSize)
n_races = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
its = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [64, 128, 256, 512]
karts, plans = R.start_grid(track, n_races, seed=20260004)
G.run(karts, plans, 0, 100)
roots, nearby = R.mcts_root_states_batch(track, prm, karts, plans)
flat = np.ascontiguousarray(roots.reshape(-1))
n = flat.shape[0]
for K in its:
    F = M.Forest(game, n, 1 + K * 16)
    F.search(flat, 2, 1)
    lib.hk_mcts_forest_search  # warm
    t0 = time.perf_counter()
    out = F.search(flat, K, 1)
    el = time.perf_counter() - t0
    print(f"roots {n} iterations {K}: {el * 1e3:.1f} ms  {n / el:.3e} decisions/s  {n * K / el:.3e} playouts/s  mean best states {out['n_best'].mean():.2f}  "
          f"mean nodes {out['n_nodes'].mean():.0f}  status!=0: {(out['status'] != 0).sum()}", flush=True)
    F.close()
