"""One warm process for ncu over the race loop: 16,384 2-kart Oval races, a few steps."""
import sys
sys.path.insert(0, '.')
from hierarchicalkarting_b200 import abi, scenarios as S, race as RC
lib = abi.load_library(); abi.check(lib.hk_init(0))
RS = RC.Races(S.OVAL, RC.race_params(S.OVAL))
karts, plans = RC.start_grid(S.OVAL, 16384, seed=20260004)
RS.run(karts, plans, 0, 100)
RS.run(karts, plans, 100, 6)
