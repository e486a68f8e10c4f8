#!/usr/bin/env python
"""Race leg of bench.py alone (BASELINE config 5: 16,384 2-kart Oval races, 200 steps): ms per step.  HK_LIB_PATH selects another build."""
import os, sys, time
sys.path.insert(0, '.')
from hierarchicalkarting_b200 import abi, scenarios as S, race as RC
if os.environ.get("HK_LIB_PATH"): abi.LIB_PATH = os.environ["HK_LIB_PATH"]
lib = abi.load_library(); abi.check(lib.hk_init(0))
RS = RC.Races(S.OVAL, RC.race_params(S.OVAL))
karts, plans = RC.start_grid(S.OVAL, 16384, seed=20260004)
RS.run(karts, plans, 0, 100)
for rep in range(3):
    t0 = time.perf_counter()
    RS.run(karts, plans, 100 + 200 * rep, 200)
    el = time.perf_counter() - t0
    print(f"race: {1e3 * el / 200:.4f} ms/step  {2 * 16384 * 200 / el:.4e} agent-steps/s", flush=True)
