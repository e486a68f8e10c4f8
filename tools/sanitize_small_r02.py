#!/usr/bin/env python
"""Small invocations of the round-2 kernels for compute-sanitizer (memcheck): the sequential search's fast path with prefix tables (fresh,
kept and overflowing trees, 2 and 4 karts, bucket 1 and 2), the general kernel, the planned 2-kart loop (fused step + recipe, packed
records), the Duos loop (split of small games, gated 4-kart kernel, planner with team scoring), the device-resident entries, the packed
LQNG entry."""
import sys
sys.path.insert(0, '.')
import os
import numpy as np
import torch
from hierarchicalkarting_b200 import abi, lqr, mcts as M, tracks, scenarios as S, race as RC
lib = abi.load_library(); abi.check(lib.hk_init(0))
rng = np.random.default_rng(0)
for track, nk, bucket, teams in ((tracks.OVAL, 2, 2, [0, 1]), (tracks.COMPLEX, 4, 2, [0, 0, 1, 1]), (tracks.COMPLEX, 3, 1, [0, 1, 1]), (tracks.OVAL, 1, 2, [0])):
    G = M.Game(track, nk, bucket)
    roots = []
    for r in range(40):
        st = tracks.root_state(track, int(rng.integers(0, 2 * track.n_sections)), [int(x) for x in rng.integers(1, 5, nk)], teams=teams,
                               tire_age=2500, times=[0] + [int(x) for x in rng.integers(0, 100, nk - 1)])
        for i in range(nk):
            st.karts[i].max_velocity = bucket
        roots.append(st)
    F = M.Forest(G, 40, 900)
    F.search(roots, 30, 1)
    F.search(roots, 25, 2, fresh=np.array([r % 2 for r in range(40)], np.int32))
    F.search(None, 40, 0, fresh=np.zeros(40, np.int32))                # some slabs fill up
    os.environ["HK_SEQ_FAST"] = "0"
    F2 = M.Forest(G, 40, 900); F2.search(roots, 12, 1)
    os.environ.pop("HK_SEQ_FAST")
    G.search_seq_batch(M._states_array(roots), 10, 3)
prob = S.make_problems(S.OVAL, 3000, 2, seed=4)
lqr.assemble_solve_packed(lqr.pack_records(prob), 2, 3, float(prob["dt"]))
prm_m = RC.race_params(S.OVAL, high_mode_mcts=True); prm_m.planEvery = 20
RM = RC.Races(S.OVAL, prm_m)
gm = M.Game(S.OVAL, 2, prm_m.velocityBucketSize)
km, pm = RC.start_grid(S.OVAL, 24, seed=2)
pl = RC.Planner(gm, 24, 12, seed=3, mode=0, first_iterations=10, reuse_cycles=3, apply_delay=7, max_tree_nodes=4000)
RM.run_planned(km, pm, pl, 0, 70)
dk, dp = RC.device_state(km, pm)
RM.run_device(dk, dp, 70, 33, planner=pl)
for K, teams in ((4, [0, 0, 1, 1]), (3, [0, 1, 1])):
    prm4 = RC.race_params(S.COMPLEX, high_mode_mcts=True); prm4.planEvery = 20
    RN = RC.RacesN(S.COMPLEX, prm4, K)
    g4 = M.Game(S.COMPLEX, K, prm4.velocityBucketSize)
    k4, p4, b4, u4 = RC.start_grid_n(S.COMPLEX, 20, K, seed=5, teams=teams)
    pl4 = RN.planner(g4, 20, 10, 7, mode=0, first_iterations=8, reuse_cycles=3, apply_delay=5, max_tree_nodes=6000)
    RN.run_n(k4, p4, b4, u4, 0, 50, planner=pl4)
    dk, dp, db = RC.device_state(k4, p4, b4)
    du = torch.zeros((20 * K, 8), dtype=torch.float64, device="cuda:0")
    RN.run_n_device(dk, dp, db, du, 50, 30, planner=pl4)
RNf = RC.RacesN(S.COMPLEX, RC.race_params(S.COMPLEX), 4)
k4, p4, b4, u4 = RC.start_grid_n(S.COMPLEX, 50, 4, seed=6)
RNf.plan_fixed(k4, p4); RNf.run_n(k4, p4, b4, u4, 0, 120)
print("sanitize_small_r02 ok")
