"""One launch of seq_search_kernel (the faithful sequential tree search, one thread per tree) for ncu: 32,768 Oval race roots, 512 iterations."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from hierarchicalkarting_b200 import abi, mcts as M, race as R, scenarios as S
lib = abi.load_library(); abi.check(lib.hk_init(0))
track = S.OVAL
prm = R.race_params(track, high_mode_mcts=True)
G = R.Races(track, prm)
game = M.Game(track, 2, prm.velocityBucketSize)
n_races = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
its = int(sys.argv[2]) if len(sys.argv) > 2 else 512
karts, plans = R.start_grid(track, n_races, seed=20260004)
G.run(karts, plans, 0, 100)
roots, nearby = R.mcts_root_states_batch(track, prm, karts, plans)
flat = np.ascontiguousarray(roots.reshape(-1))
F = M.Forest(game, flat.shape[0], 1 + its * 16)
out = F.search(flat, its, 1)
print("nodes mean", out["n_nodes"].mean(), "best", out["n_best"].mean())
