"""Generates tests/golden/*.  The reference (Unity C#) cannot be executed in this image (no .NET), so the LQNG vectors
are produced by the CPU oracle *after* it has been cross-checked against the independent numpy restatement
(oracle/np_lqng.py) and the surveyor's smoke value; the game known-answers are the surveyor's independently computed
float32 numbers of SURVEY.md Appendix D, typed in by hand below (they are NOT produced by the oracle).
Run from the repo root: python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hierarchicalkarting_b200 import scenarios as S  # noqa: E402
from oracle import oracle as O, np_lqng  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def lqng():
    out = {}
    for N, track, seed in ((2, S.OVAL, 4242), (4, S.COMPLEX, 4343)):
        p = S.make_problems(track, 8, N, seed=seed)
        A, B, Q, q, R, x0 = S.assemble_dense(p)
        r = O.lqng_solve_batch(A, B, Q, q, R, x0, 3)
        for b in range(8):      # second opinion before freezing
            u, Ps, als, uall = np_lqng.solve_feedback_lqr(list(A[b]), list(B[b]), list(Q[b]), list(q[b]), list(R[b]),
                                                          list(x0[b].reshape(N, 4)), 3)
            assert np.allclose(Ps, r["P"][b], rtol=1e-11, atol=1e-13) and np.allclose(uall, r["u0"][b], rtol=1e-11, atol=1e-13)
        for k, v in zip(("A", "B", "Q", "q", "R", "x0"), (A, B, Q, q, R, x0)):
            out[f"N{N}_{k}"] = v
        for k in ("u0", "P", "alpha", "traj"):
            out[f"N{N}_{k}_out"] = r[k]
    np.savez_compressed(os.path.join(HERE, "lqng_golden.npz"), **out)


# SURVEY.md Appendix D (surveyor's float32 emulation; Acc 7, Brk 16, Top 15, MaxGs 2, MinGs .5, TWF .001, precision 100)
# section = (insideR, length, width, turnDeg, leftTurn); from = (lane, min_v, max_v, tireAge, timeAtSection); action = (min_v, max_v, lane)
GAME_KAT = {
    "max_speed_for_radius_and_wear": [[17.5, 0.25, 15.0], [1.0, 0.3, 3.8994231], [8.25, 0.0, 12.722618], [20.0, 1.4, 15.0]],
    "apply_action": [
        {"section": [0, 10, 10, 0, 0], "from": [2, 0, 2, 2500, 0], "action": [6, 8, 2], "dist": 10.0, "radius": 0.0, "toc": 1.6026193, "time": 160, "tireLoad": 0.099999994, "tireAge": 2500, "infeasible": 0},
        {"section": [0, 10, 10, 0, 0], "from": [2, 0, 2, 2500, 0], "action": [6, 8, 4], "dist": 12.018505, "radius": 0.0, "toc": 1.7848942, "time": 178, "tireLoad": 0.12018505, "tireAge": 2501, "infeasible": 0},
        {"section": [0, 10, 10, 0, 0], "from": [2, 0, 2, 2500, 0], "action": [14, 15, 2], "dist": 10.0, "radius": 0.0, "toc": -1.0, "time": -100, "infeasible": 1},
        {"section": [0, 10, 10, 0, 0], "from": [3, 12, 14, 2600, 350], "action": [14, 15, 3], "dist": 10.0, "radius": 0.0, "toc": 0.6862351, "time": 418, "tireLoad": 0.099999994, "tireAge": 2600, "infeasible": 0},
        {"section": [15, 10, 10, 45, 1], "from": [1, 10, 12, 2700, 500], "action": [10, 12, 1], "dist": 11.780972, "radius": 15.0, "toc": 0.89492196, "time": 589, "tireLoad": 1.1309735, "tireAge": 2711, "infeasible": 0},
        {"section": [15, 10, 10, 45, 1], "from": [4, 12, 14, 2700, 500], "action": [14, 15, 2], "dist": 15.707964, "radius": 20.0, "toc": 1.066766, "time": 606, "tireLoad": 1.767146, "tireAge": 2717, "infeasible": 0},
        {"section": [1, 10, 10, 90, 0], "from": [3, 6, 8, 3000, 900], "action": [6, 8, 4], "dist": 3.5342917, "radius": 2.25, "toc": 0.5299821, "time": 952, "tireLoad": 1.0053097, "tireAge": 3010, "infeasible": 0},
        {"section": [7, 10, 10, 70, 1], "from": [2, 8, 10, 3000, 900], "action": [6, 8, 3], "dist": 13.133602, "radius": 10.75, "toc": 1.1749473, "time": 1017, "tireLoad": 0.7819075, "tireAge": 3007, "infeasible": 0},
        {"section": [0, 10, 10, 0, 0], "from": [1, 14, 15, 2500, 100], "action": [6, 8, 1], "dist": 10.0, "radius": 0.0, "toc": 0.8011905, "time": 180, "tireLoad": 0.099999994, "tireAge": 2500, "infeasible": 0},
    ],
    # SURVEY.md B.7 closed-form examples of the rollout policy's index distribution
    "policy_pmf": {"3": [0.6827, 0.3146, 0.0026], "8": [0.2923, 0.4471, 0.1998, 0.0521, 0.0079, 0.0007],
                   "20": [0.1192, 0.2281, 0.1995, 0.1595, 0.1167, 0.0781, 0.0478, 0.0267]},
    # Random123 known answers for Philox4x32-10 (key, counter) -> output
    "philox4x32_10": [{"seed": 0, "ctr": [0, 0, 0, 0], "out": [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]},
                      {"seed": 0xffffffffffffffff, "ctr": [0xffffffff] * 4, "out": [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]},
                      {"seed": 0x299f31d0a4093822, "ctr": [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], "out": [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]}],
    # SURVEY.md Appendix D LQNG smoke value (surveyor's numpy): me=(10,2,12,.3) opp=(13,3.5,11,.25) targets (20,5,15,.35)/(20,7.5,15,.3)
    "lqng_smoke_u0": [0.00558, 0.076146],
}

def game():
    """tests/golden/game_golden.npz: 400 (kart state, action) -> kart state transitions and 24 sequential tree searches, produced by the
    C oracle AFTER the two independent restatements (oracle/np_game.py for the arithmetic, oracle/np_mcts_seq.py for the tree) agreed
    with it on exactly these inputs."""
    from hierarchicalkarting_b200 import tracks
    from oracle import np_game, np_mcts_seq, structs as OS
    out = {}
    rng = np.random.default_rng(20260400)
    for name, track, bucket in (("oval2", tracks.OVAL, 2), ("complex1", tracks.COMPLEX, 1)):
        karts = tracks.kart_array(2)
        params = tracks.game_params(track, bucket)
        g = O.Game(track.sections_array(), track.n_sections, karts, 2, params)
        ng = np_game.NpGame(track.sections_array(), track.n_sections, karts, 2, params)
        n = 200
        ks = np.zeros(n, dtype=OS.np_dtype(OS.hk_kart_state))
        ks["section"] = rng.integers(0, 3 * track.n_sections, n)
        ks["lane"] = rng.integers(1, 5, n)
        vmin = rng.choice([0] + list(range(6, 15, bucket)), n)
        ks["min_velocity"] = vmin
        ks["max_velocity"] = np.where(vmin == 0, bucket, np.minimum(vmin + bucket, 15))
        ks["tireAge"] = rng.choice([0, 2500, 6000, 9999, 13400, 20000], n)
        ks["laneChanges"] = rng.integers(0, 4, n)
        ks["timeAtSection"] = rng.integers(0, 2000, n)
        a_min = rng.choice(list(range(6, 15, bucket)), n).astype(np.int32)
        acts = np.stack([a_min, np.minimum(a_min + bucket, 15), rng.integers(1, 5, n)], axis=1).astype(np.int32)
        new = np.zeros_like(ks)
        for i in range(n):
            ref = g.apply_action(OS.hk_kart_state.from_buffer_copy(ks[i].tobytes()), tuple(int(v) for v in acts[i]))
            new[i] = np.frombuffer(bytes(ref), dtype=ks.dtype)[0]
        assert new.tobytes() == ng.apply_actions(ks, acts[:, 0], acts[:, 1], acts[:, 2]).tobytes()      # second opinion before freezing
        out[f"{name}_kart_states"], out[f"{name}_actions"], out[f"{name}_new_states"] = ks, acts, new
        roots = np.zeros(12, dtype=OS.GAME_STATE_DTYPE)
        for r in range(12):
            st = tracks.root_state(track, int(rng.integers(0, 2 * track.n_sections)), [int(x) for x in rng.integers(1, 5, 2)], teams=[0, 1],
                                   tire_age=int(rng.choice([0, 2500, 9900])), times=[0, int(rng.integers(0, 120))])
            for i in range(2):
                st.karts[i].max_velocity = bucket
            roots[r] = np.frombuffer(bytes(st), dtype=OS.GAME_STATE_DTYPE)[0]
        res = O.tree_search_batch(g, roots, 64, seed=20260401, mode=0, threads=1)
        for r in (0, 5):                                                                                # second opinion on two of them
            s = np_mcts_seq.SequentialSearch(ng, 20260401 + r)
            root = s.constructSearchTree(OS.hk_game_state.from_buffer_copy(roots[r].tobytes()), 64)
            best = s.getBestStatesSequence(root)
            assert len(best) == int(res["n_best"][r]) and all(bytes(b) == res["best"][r, k].tobytes() for k, b in enumerate(best))
        out[f"{name}_roots"] = roots
        for k in ("best", "n_best", "root_gen", "root_episodes", "root_values", "n_nodes"):
            out[f"{name}_search_{k}"] = res[k]
    np.savez_compressed(os.path.join(HERE, "game_golden.npz"), **out)


def recipe():
    """tests/golden/recipe_golden.npz: race states of 40 Duos races (4 karts, teams [0, 0, 1, 1], Complex, both high-level modes) and 60
    2-kart Oval races, and every agent's SolveLQR problem in the 4-player layout, produced by the C oracle AFTER the Python restatement
    (oracle/np_recipe.py) agreed with it on exactly these inputs."""
    sys.path.insert(0, os.path.dirname(HERE))
    from hierarchicalkarting_b200 import race as R
    from oracle import np_race
    from test_race_cpu import _oracle_races, _track_tables
    out = {}
    for name, track, K, teams, n in (("duos", S.COMPLEX, 4, [0, 0, 1, 1], 40), ("pair", S.OVAL, 2, [0, 1], 60)):
        rng = np.random.default_rng(20260500 + K)
        karts, plans, beliefs, _ = R.start_grid_n(track, n, K, seed=3, teams=teams)
        L = track.n_sections
        lanes_xy, head = track.lane_table(), track.heading_table()
        base = rng.integers(0, 2 * L, size=n)
        for e in range(K):
            sec = np.maximum(base + rng.integers(-1, 2, size=n), 0)
            s0 = sec % L
            ln = rng.integers(1, 5, size=n)
            p0, p1 = lanes_xy[s0, ln - 1], lanes_xy[(s0 + 1) % L, ln - 1]
            pos = p0 + (p1 - p0) * rng.random(n)[:, None] + rng.normal(0, 5.0 * rng.random(n)[:, None], size=(n, 2))
            karts["x"][:, e], karts["z"][:, e] = pos[:, 0], pos[:, 1]
            karts["v"][:, e] = np.where(rng.random(n) < 0.15, rng.uniform(0, 5, n), rng.uniform(5, 15, n))
            karts["h"][:, e] = np.mod(head[s0] + rng.normal(0, 0.2, n), 2 * np.pi)
            karts["section"][:, e], karts["lane"][:, e] = sec, ln
        karts["active"][::13, K - 1] = 0
        for arr in (plans, beliefs):
            on = rng.random(arr["lane"].shape) < 0.6
            arr["lane"][:] = np.where(on, rng.integers(1, 5, size=on.shape), 0)
            arr["vel"][:] = np.where(on, rng.choice([8, 10, 12, 14, 15], size=on.shape), 0)
        out[f"{name}_karts"], out[f"{name}_plans"], out[f"{name}_beliefs"] = karts, plans, beliefs
        for mcts in (False, True):
            OR, prm = _oracle_races(O, track, high_mode_mcts=mcts)
            tt = _track_tables(track)
            res = dict(n_players=np.zeros((n, K), np.int32), players=np.full((n, K, 4), -1, np.int32), x0=np.zeros((n, K, 4, 4)),
                       target=np.zeros((n, K, 4, 4)), tw=np.zeros((n, K, 4, 4)), cw=np.ones((n, K, 4)), aw=np.zeros((n, K, 4, 3, 2)),
                       otgt=np.zeros((n, K, 4, 3, 4)), otw=np.zeros((n, K, 4, 3, 3)))
            for r in range(n):
                for e in range(K):
                    c = OR.recipe_n_one(K, karts[r], plans[r], beliefs[r, e], e)
                    a = np_race.recipe_agent(tt, prm, karts[r], plans[r], beliefs[r], e)                # second opinion before freezing
                    N = len(c["players"])
                    assert a["players"] == c["players"]
                    for k in ("x0", "tw", "cw", "aw", "otgt", "otw"):
                        assert np.array_equal(np.asarray(a[k]).reshape(c[k].shape), c[k]), (k, r, e)
                    assert np.max(np.abs(np.asarray(a["target"]) - c["target"])) <= 1e-13
                    res["n_players"][r, e] = N
                    res["players"][r, e, :N] = c["players"]
                    for k in ("x0", "target", "tw"):
                        res[k][r, e, :N] = c[k]
                    res["cw"][r, e, :N] = c["cw"]
                    for k in ("aw", "otgt", "otw"):
                        res[k][r, e, :N, :max(N - 1, 0)] = c[k]
            for k, v in res.items():
                out[f"{name}_{'mcts' if mcts else 'fixed'}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "recipe_golden.npz"), **out)


if __name__ == "__main__":
    lqng()
    game()
    recipe()
    with open(os.path.join(HERE, "game_kat.json"), "w") as f:
        json.dump(GAME_KAT, f, indent=1)
    print("golden written")
