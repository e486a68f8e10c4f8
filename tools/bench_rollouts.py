#!/usr/bin/env python
"""MCTS rollouts/s (BASELINE config 4: Complex, 2 karts, depth 8, bucket 2, 10^6 rollouts per decision), host call incl. D2H."""
import sys, time
sys.path.insert(0, '.')
import os
from hierarchicalkarting_b200 import abi, mcts as M, tracks
if os.environ.get("HK_LIB_PATH"): abi.LIB_PATH = os.environ["HK_LIB_PATH"]
lib = abi.load_library(); abi.check(lib.hk_init(0))
for nk, lanes, teams, times in ((2, [2, 3], [0, 1], [0, 80]), (4, [2, 3, 2, 3], [0, 0, 1, 1], [0, 80, 30, 50])):
    G = M.Game(tracks.COMPLEX, nk, 2)
    leaf = tracks.root_state(tracks.COMPLEX, 3, lanes, teams=teams, tire_age=2500, times=times)
    for w in range(3): G.rollouts(leaf, 1_000_000, seed=1)
    t0 = time.perf_counter(); plies = 0
    for r in range(10): plies += G.rollouts(leaf, 1_000_000, seed=2 + r)["plies"]
    el = time.perf_counter() - t0
    print(f"{nk} karts: {1e7 / el:.4e} rollouts/s  {plies / el:.4e} plies/s  {el * 100:.3f} ms per 1e6 rollouts", flush=True)
