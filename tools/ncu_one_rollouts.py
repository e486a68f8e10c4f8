"""One warm process for `ncu -k regex:rollouts_kernel`: BASELINE config 4 leaf, 10^6 rollouts per launch."""
import sys
sys.path.insert(0, '.')
from hierarchicalkarting_b200 import abi, mcts as M, tracks
lib = abi.load_library(); abi.check(lib.hk_init(0))
G = M.Game(tracks.COMPLEX, 2, 2)
leaf = tracks.root_state(tracks.COMPLEX, 3, [2, 3], teams=[0, 1], tire_age=2500, times=[0, 80])
for w in range(4): G.rollouts(leaf, 1_000_000, seed=1 + w)
