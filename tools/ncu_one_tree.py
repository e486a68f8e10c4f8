"""One warm process for `ncu -k regex:tree_search`: 4,096 Oval race roots, 24 iterations x 16 rollouts per leaf."""
import sys
sys.path.insert(0, '.')
import numpy as np
from hierarchicalkarting_b200 import abi, mcts as M, race as R, scenarios as S
lib = abi.load_library(); abi.check(lib.hk_init(0))
track = S.OVAL
prm = R.race_params(track, high_mode_mcts=True)
G = R.Races(track, prm)
game = M.Game(track, 2, prm.velocityBucketSize)
karts, plans = R.start_grid(track, 2048, seed=20260004)
G.run(karts, plans, 0, 100)
roots, nearby = R.mcts_root_states_batch(track, prm, karts, plans)
flat = np.ascontiguousarray(roots.reshape(-1))
for k in range(3):
    game.search_batch_array(flat, 24, 16, 1 + k)
