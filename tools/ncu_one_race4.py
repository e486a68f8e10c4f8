"""One warm process for ncu over the Duos race loop: 8,192 4-kart Complex races, a few steps."""
import sys
sys.path.insert(0, '.')
from hierarchicalkarting_b200 import abi, scenarios as S, race as RC
lib = abi.load_library(); abi.check(lib.hk_init(0))
RN = RC.RacesN(S.COMPLEX, RC.race_params(S.COMPLEX), 4)
k4, p4, b4, u4 = RC.start_grid_n(S.COMPLEX, 8192, 4, seed=20260007)
RN.plan_fixed(k4, p4)
RN.run_n(k4, p4, b4, u4, 0, 400)
RN.run_n(k4, p4, b4, u4, 400, 8)
