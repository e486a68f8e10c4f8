#!/usr/bin/env python
"""hk_mcts_search_batch throughput: n roots (Oval race states after 100 steps), K iterations, R rollouts per leaf."""
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
karts, plans = R.start_grid(track, n_races, seed=20260004)
G.run(karts, plans, 0, 100)
roots, nearby = R.mcts_root_states_batch(track, prm, karts, plans)
flat = np.ascontiguousarray(roots.reshape(-1))
for K, RPL in ((8, 16), (24, 16), (24, 32), (48, 16)):
    game.search_batch_array(flat[:64], K, RPL, 1)
    t0 = time.perf_counter()
    out = game.search_batch_array(flat, K, RPL, 1)
    el = time.perf_counter() - t0
    eps = out["root_episodes"].sum(axis=1)
    print(f"roots {flat.shape[0]} K {K} R {RPL}: {el * 1e3:.1f} ms  {flat.shape[0] / el:.3e} decisions/s  rollouts {eps.sum():.3e} ({eps.sum() / el:.3e}/s)  "
          f"mean best states {out['n_best'].mean():.2f}  mean nodes {out['n_nodes'].mean():.0f}", flush=True)
t0 = time.perf_counter(); R.mcts_root_states_batch(track, prm, karts, plans); t1 = time.perf_counter()
R.apply_best_states_batch(track, karts, plans, nearby, out["best"].reshape(n_races, 2, abi.HK_MCTS_MAX_SEQ), out["n_best"].reshape(n_races, 2)); t2 = time.perf_counter()
print(f"host: root states {1e3 * (t1 - t0):.1f} ms, hand-off {1e3 * (t2 - t1):.1f} ms")
