"""One forest search of 4-kart Complex games (Duos) for ncu: 32,768 trees x 256 iterations from race states after 100 steps."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from hierarchicalkarting_b200 import abi, mcts as M, race as RC, scenarios as S, tracks
lib = abi.load_library(); abi.check(lib.hk_init(0))
track = S.COMPLEX
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
its = int(sys.argv[2]) if len(sys.argv) > 2 else 256
game = M.Game(track, 4, 2)
rng = np.random.default_rng(1)
roots = []
for r in range(256):
    sec = int(rng.integers(0, 2 * track.n_sections))
    st = tracks.root_state(track, sec, [int(x) for x in rng.integers(1, 5, 4)], teams=[0, 0, 1, 1], tire_age=2500, times=[0] + [int(x) for x in rng.integers(0, 120, 3)])
    for i in range(4):
        st.karts[i].max_velocity = 2
    roots.append(st)
arr = np.tile(M._states_array(roots), n // 256)
F = M.Forest(game, n, 1 + its * 32)
import time
F.search(arr[:n], 2, 1)
t0 = time.perf_counter()
out = F.search(arr[:n], its, 1)
print("ms", 1e3 * (time.perf_counter() - t0), "nodes mean", out["n_nodes"].mean(), "best", out["n_best"].mean(), "status", (out["status"] != 0).sum())
