"""One planning event of the Duos loop with the MCTS planner for ncu: 8,192 4-kart Complex races, steps 100..100 (+ a few)."""
import sys
sys.path.insert(0, '.')
from hierarchicalkarting_b200 import abi, mcts as M, scenarios as S, race as RC
lib = abi.load_library(); abi.check(lib.hk_init(0))
R4 = 8192
prm = RC.race_params(S.COMPLEX, high_mode_mcts=True)
RNm = RC.RacesN(S.COMPLEX, prm, 4)
game4 = M.Game(S.COMPLEX, 4, prm.velocityBucketSize)
km, pm, bm, um = RC.start_grid_n(S.COMPLEX, R4, 4, seed=20260007)
RNm.run_n(km, pm, bm, um, 0, 100)
pl = RNm.planner(game4, R4, 256, 20260008, mode=0, reuse_cycles=3, apply_delay=45)
RNm.run_n(km, pm, bm, um, 100, 2, planner=pl)
pl.close()
