"""CPU tests of the closed-loop ORACLE (oracle/hk_oracle_race.c): the C restatement of the SolveLQR problem recipe is pinned
by the independently written numpy recipe of scenarios.make_problems; planFixed / plant / checkpoint bookkeeping are checked on
hand-built cases; the loop is checked through size-independent properties (monotone sections, finish, freeze)."""
import numpy as np
import pytest

from hierarchicalkarting_b200 import abi, race as R, scenarios as S


def _oracle_races(oracle, track, **kw):
    prm = R.race_params(track, **kw)
    sec, trig, fwd, lane = R.geometry(track)
    return oracle.Races(sec, trig, fwd, lane, track.n_sections, prm), prm


def karts_from_problems(p, track):
    """Race state whose (race, ego 0) recipe must reproduce make_problems: same states, the sampled lanes / velocities as the
    ego's own plan and as its belief about the other kart."""
    sm, batch, L = p["sampled"], p["x0"].shape[0], track.n_sections
    karts = np.zeros((batch, 2), dtype=abi.RACE_KART_DTYPE)
    plans = np.zeros((batch, 2), dtype=abi.RACE_PLAN_DTYPE)
    for i in range(2):
        for c, k in enumerate(("x", "z", "v", "h")):
            karts[k][:, i] = p["x0"][:, i, c]
        karts["section"][:, i] = sm["section"]
    karts["active"], karts["steer"], karts["lane"] = 1, 3.25, 2
    s1, s2, rows = (sm["section"] + 1) % L, (sm["section"] + 2) % L, np.arange(batch)
    for own, lk, vk in ((0, "lane", "vel"), (1, "oppLane", "oppVel")):
        plans[lk][rows, 0, s1] = sm["tgt_lane"][:, own]; plans[vk][rows, 0, s1] = sm["bucket_max"][:, own]
        plans[lk][rows, 0, s2] = sm["nxt_lane"][:, own]; plans[vk][rows, 0, s2] = sm["nbucket_max"][:, own]
    return karts, plans


@pytest.mark.parametrize("track", [S.OVAL, S.COMPLEX])
@pytest.mark.parametrize("mcts", [True, False])
def test_oracle_recipe_equals_numpy_recipe(oracle, track, mcts):
    p = S.make_problems(track, 3000, 2, seed=7, high_mode_mcts=mcts)
    karts, plans = karts_from_problems(p, track)
    OR, _ = _oracle_races(oracle, track, high_mode_mcts=mcts)
    out = OR.recipe(karts, plans)
    for k in ("x0", "tw", "cw", "aw", "otgt", "otw"):
        assert np.array_equal(out[k][0::2], p[k]), k
    assert np.max(np.abs(out["target"][0::2] - p["target"])) <= 1e-14      # libm vs numpy sin/cos in AngleDifference
    # all three target-heading branches and the stopped branch are exercised
    d_t = np.linalg.norm((p["target"][:, 0, :2] - p["x0"][:, 0, :2]), axis=-1)
    assert (p["x0"][:, :, 2] <= 5).any() and (d_t > 10.5).any() and (d_t < 7.5).any()


def test_plan_fixed_and_absent_plan_targets(oracle):
    track = S.OVAL
    OR, prm = _oracle_races(oracle, track)
    karts, plans = R.start_grid(track, 4, seed=1)
    out = OR.recipe(karts, plans)
    assert np.allclose(out["target"][:, 0, :2], track.trigger_table()[1])          # no plan: Trigger of section+1 (HKA:762-766)
    assert np.all(out["target"][:, 0, 2] == 0.0)                                   # v <= 5: target speed 0 (HKA:808)
    karts["section"][0, 0] = 20
    plans["lane"][0, 0, 22] = 1                                                    # an existing key is kept
    OR.plan_fixed(karts, plans)
    want = [track.rows[(i - 1) % 24][6] for i in range(21, 29)]
    got = [int(plans["lane"][0, 0, i % 24]) for i in range(21, 29)]
    want[1] = 1
    assert got == want and np.all(plans["vel"][0, 0, [i % 24 for i in range(21, 29) if i != 22]] == 15.0)
    assert plans["lane"][0, 0, 5] == 0 and plans["lane"][0, 0, 20] == 0
    assert [int(plans["lane"][1, 1, i]) for i in range(1, 9)] == [track.rows[i - 1][6] for i in range(1, 9)]


def test_step_bookkeeping_hand_cases(oracle):
    track = S.OVAL
    OR, prm = _oracle_races(oracle, track)
    karts, plans = R.start_grid(track, 1, seed=1)
    k = karts[0]
    # kart 0: 0.1 m before gate 1 (z = 7.9) near lane 4, driving up at 10 m/s; kart 1: stays far from any gate
    k["x"][0], k["z"][0], k["v"][0], k["h"][0], k["lane"][0], k["laneChanges"][0] = 19.3, 7.8, 10.0, np.pi / 2, 2, 2
    plans["lane"][0, 0, 1], plans["vel"][0, 0, 1] = 4, 12.0
    before1 = karts[0, 1].copy()
    OR.step(karts, plans, np.array([[1.0, 0.0], [0.0, 0.0]]), 77)
    assert k["section"][0] == 1 and k["lane"][0] == 4 and k["sectionStep"][0] == 77
    assert k["illegalLaneChanges"][0] == 1 and k["laneChanges"][0] == 4            # 2 + |2-4| > 3 on a straight (HKA:638-650)
    assert plans["lane"][0, 0, 1] == 0                                              # key consumed (HKA:631-632)
    assert abs(k["v"][0] - (10.0 + prm.dt * 7.0)) < 1e-12 and abs(k["z"][0] - (7.8 + prm.dt * 10.0)) < 1e-12
    # u0 == 0: coasting and no steering (HKA:1216-1221); kart 1 only coasts
    assert karts[0, 1]["section"] == 0 and karts[0, 1]["h"] == before1["h"]
    assert abs(karts[0, 1]["v"] - max(0.0, before1["v"] - prm.dt * 5.0)) < 1e-12
    # steering: positive angular-velocity command turns clockwise (Unity yaw), clamped to 0.4 steer
    h0 = float(k["h"][0])
    OR.step(karts, plans, np.array([[-1.0, 9.0], [0.0, 0.0]]), 78)
    assert abs(k["h"][0] - (h0 - prm.dt * float(np.float32(0.4) * np.float32(3.25)))) < 1e-7 and k["v"][0] < 10.14
    # straight -> curve resets the counter (HKA:643-646): gate 4 of the Oval starts the first curve
    k["section"][0], k["lane"][0], k["laneChanges"][0] = 3, 3, 3
    k["x"][0], k["z"][0], k["h"][0] = 17.1, 38.1, np.pi / 2
    OR.step(karts, plans, np.array([[1.0, 0.0], [0.0, 0.0]]), 79)
    assert k["section"][0] == 4 and k["laneChanges"][0] == 0


@pytest.mark.parametrize("track", [S.OVAL, S.COMPLEX])
def test_loop_properties(oracle, track):
    OR, prm = _oracle_races(oracle, track, laps=1)
    karts, plans = R.start_grid(track, 32, seed=11)
    last = karts["section"].copy()
    for blk in range(16):
        _, bad = OR.run(karts, plans, blk * 200, 200)
        assert bad == 0 and np.all(karts["section"] >= last) and np.all(np.isfinite(karts["x"]))
        assert np.all((karts["v"] >= 0) & (karts["v"] <= 15.0)) and np.all((karts["h"] >= 0) & (karts["h"] < 2 * np.pi))
        last = karts["section"].copy()
    assert np.all(karts["active"] == 0) and np.all(karts["section"] == prm.goalSection)    # everybody finished one lap
    frozen = karts.copy()
    OR.run(karts, plans, 3200, 50)
    assert np.array_equal(frozen, karts)                                                   # finished karts do not move


def test_telemetry_log_format_and_reference_parser(oracle, tmp_path, capsys):
    """The experiment log of a headless batch has the reference's line format
    (TelemetryViewer.cs:90-104); when the reference checkout is present (this container only) its own
    experiment_log_parser.py must score it."""
    import importlib.util
    import os
    from hierarchicalkarting_b200 import telemetry as T
    track = S.OVAL
    OR, prm = _oracle_races(oracle, track, laps=2)
    karts, plans = R.start_grid(track, 6, seed=3)
    OR.run(karts, plans, 0, 2400)
    logs = tmp_path / "ExperimentLogs"
    logs.mkdir()
    names = ["Fixed-LQR(A)", "Fixed-LQR(B)"]
    T.write_experiment_log(str(logs / "GPU_Fixed.txt"), names, karts, plans, track.n_sections, 2, 2400)
    text = (logs / "GPU_Fixed.txt").read_text().splitlines()
    assert text[0] == "Experiment 0" and text[1].startswith("Fixed-LQR(A) Speed: ") and text[5] == "Fixed-LQR(A) Laps Completed: 2/2"
    assert text[19].startswith("Winner: Fixed-LQR(") and text[20] == "" and text[21] == "Experiment 1"
    laps = [float(l.rsplit(" ", 1)[1]) for l in text if " Best Lap: " in l]
    assert len(laps) == 12 and all(17.0 < t < 23.0 for t in laps)       # the reference's own Oval logs: 19.3 - 20.5 s per lap
    parser = "/root/reference/experiment_log_parser.py"
    if not os.path.exists(parser):
        pytest.skip("reference checkout not present (GPU box)")
    import ast
    tree = ast.parse(open(parser).read())
    keep = [n for n in tree.body if isinstance(n, (ast.Import, ast.FunctionDef))
            or (isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") in ("logs_dir", "points_per_position"))]
    src = ast.Module(body=keep, type_ignores=[])
    ns = {}
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        exec(compile(src, parser, "exec"), ns)                              # defines the functions only
        ns["summarize_experiment"]("GPU_Fixed")
    finally:
        os.chdir(cwd)
    out = capsys.readouterr().out
    assert "Wins {'Fixed-LQR': 6}" in out and "DNFs {}" in out and "Avg Collisions {'Fixed-LQR': 0.0}" in out


def test_batched_mcts_root_states_and_handoff_equal_per_agent_versions():
    """race.mcts_root_states_batch / apply_best_states_batch (what plan_mcts_batch feeds hk_mcts_search_batch with and applies) are the
    vectorised forms of mcts_root_state / apply_best_states (HierarchicalKartAgent.cs:180-245, 366-402)."""
    import ctypes as C
    from hierarchicalkarting_b200 import abi, race as R, tracks
    rng = np.random.default_rng(11)
    track = tracks.COMPLEX
    prm = R.race_params(S.COMPLEX, high_mode_mcts=True)
    n = 300
    karts = np.zeros((n, 2), dtype=abi.RACE_KART_DTYPE)
    plans = np.zeros((n, 2), dtype=abi.RACE_PLAN_DTYPE)
    karts["section"] = rng.integers(0, 90, size=(n, 1)) + rng.integers(-3, 4, size=(n, 2)).clip(0)
    karts["section"][:20] = 0
    karts["lane"] = rng.integers(1, 5, size=(n, 2))
    karts["laneChanges"] = rng.integers(0, 4, size=(n, 2))
    karts["steer"] = rng.uniform(1.0, 4.0, size=(n, 2)).astype(np.float32)
    plans["sectionTimes"] = rng.integers(0, 5000, size=(n, 2, abi.HK_MAX_SECTIONS))
    roots, nearby = R.mcts_root_states_batch(track, prm, karts, plans)
    for r in range(n):
        for e in (0, 1):
            st, nb = R.mcts_root_state(track, prm, karts[r], plans[r], e)
            assert bytes(st) == roots[r, e].tobytes(), (r, e)
            assert list(nearby[r, e, :len(nb)]) == nb and all(v == -1 for v in nearby[r, e, len(nb):])
    # hand-off: synthetic best sequences (sections ahead of the karts, random lanes / buckets)
    best = np.zeros((n, 2, abi.HK_MCTS_MAX_SEQ), dtype=abi.GAME_STATE_DTYPE)
    n_best = rng.integers(0, 9, size=(n, 2)).astype(np.int32)
    for k in range(abi.HK_MCTS_MAX_SEQ):
        best[:, :, k] = roots
        for slot in (0, 1):
            best["karts"][:, :, k, slot]["section"] = roots["initialSection"] + k + rng.integers(0, 2, size=(n, 2))
            best["karts"][:, :, k, slot]["lane"] = rng.integers(1, 5, size=(n, 2))
            best["karts"][:, :, k, slot]["max_velocity"] = rng.integers(8, 16, size=(n, 2))
    pa, pb = plans.copy(), plans.copy()
    R.apply_best_states_batch(track, karts, pa, nearby, best, n_best)

    class _GS:
        def __init__(self, rec):
            self.state = abi.hk_game_state.from_buffer_copy(rec.tobytes())
    for r in range(n):
        for e in (0, 1):
            nb = [int(v) for v in nearby[r, e] if v >= 0]
            R.apply_best_states(track, karts[r], pb[r], e, nb, [_GS(best[r, e, k]) for k in range(int(n_best[r, e]))])
    for f in ("lane", "vel", "oppLane", "oppVel"):
        assert np.array_equal(pa[f], pb[f]), f


def _track_tables(track):
    sec, trig, fwd, lane = R.geometry(track)
    return dict(trig=trig, lane=lane, straight=[sec[i].insideR == 0.0 for i in range(track.n_sections)])


@pytest.mark.parametrize("track", [S.OVAL, S.COMPLEX])
@pytest.mark.parametrize("mcts", [True, False])
def test_oracle_recipe_equals_independent_restatement(oracle, track, mcts):
    """hk_oracle_race.c's recipe against oracle/np_recipe.py — a restatement of SolveLQR's problem construction written from the C#
    text (HierarchicalKartAgent.cs:699-1201) that lives under oracle/, shares no code with the C file and none with the product
    package: every compact field of both agents' problems on 1,500 race states (start grids driven forward, plans from planFixed and
    hand-written beliefs, a finished kart), integers of the weights bit for bit, headings to 1e-14."""
    from oracle import np_recipe
    OR, prm = _oracle_races(oracle, track, high_mode_mcts=mcts)
    tt = _track_tables(track)
    rng = np.random.default_rng(77)
    n = 750
    karts, plans = R.start_grid(track, n, seed=5)
    L = track.n_sections
    sec = rng.integers(0, 2 * L, size=(n, 2))
    sec[:, 1] = sec[:, 0] + rng.integers(-1, 2, size=n)
    sec = np.maximum(sec, 0)
    lanes_xy, head = track.lane_table(), track.heading_table()
    for e in range(2):
        s0 = sec[:, e] % L
        ln = rng.integers(1, 5, size=n)
        p0, p1 = lanes_xy[s0, ln - 1], lanes_xy[(s0 + 1) % L, ln - 1]
        fr = rng.random(n)[:, None]
        pos = (p0 + (p1 - p0) * fr + rng.normal(0, 0.5, size=(n, 2))).astype(np.float32)
        karts["x"][:, e], karts["z"][:, e] = pos[:, 0], pos[:, 1]
        karts["v"][:, e] = np.where(rng.random(n) < 0.15, rng.uniform(0, 5, n), rng.uniform(5, 15, n)).astype(np.float32)
        karts["h"][:, e] = np.mod(head[s0] + rng.normal(0, 0.2, n), 2 * np.pi)
        karts["section"][:, e] = sec[:, e]
        karts["lane"][:, e] = ln
    karts["active"][::97, 1] = 0
    for key, vk in (("lane", "vel"), ("oppLane", "oppVel")):
        on = rng.random((n, 2, abi.HK_MAX_SECTIONS)) < 0.6
        plans[key][:] = np.where(on, rng.integers(1, 5, size=on.shape), 0)
        plans[vk][:] = np.where(on, rng.choice([8, 10, 12, 14, 15], size=on.shape), 0)
    ref = OR.recipe(karts, plans)
    for r in range(n):
        for e in range(2):
            got = np_recipe.race_recipe_2kart(tt, prm, karts[r], plans[r], e)
            b = 2 * r + e
            assert got["players"] == [e, 1 - e]
            for k in ("x0", "tw", "cw", "aw", "otgt", "otw"):
                assert np.array_equal(np.asarray(got[k]).reshape(ref[k][b].shape), ref[k][b]), (k, r, e, got[k], ref[k][b])
            assert np.max(np.abs(got["target"] - ref["target"][b])) <= 1e-14


def test_fixed_mode_avoid_weight_known_answer(oracle):
    """the ego's avoid-weight multiplier is `HighMode == Fixed ? 0.45f : 1.0f` (HierarchicalKartAgent.cs:999-1002), the
    other player's 1.3f.  Known answers worked out by hand from :1019 for two karts 4 m apart (d^1.5 = 8 exactly):
      Fixed, ego:  0.45f = 0.449999988079071044921875, x 8 = 3.599999904632568359375 (exact), 1 / that = 0.277777785136... ->
                   nearest float32 0.2777777910232543945 (the 1.0f multiplier of round 1 gave 0.125: 2.2x too small a weight)
      MCTS,  ego:  1 / (8 x 1.0f) = 0.125
      other:       1.3f = 1.2999999523162841796875, x 8 = 10.3999996185302734375, 1 / that = 0.0961538496... -> float32 0.09615384787321091"""
    track = S.OVAL
    for mcts, want_ego in ((False, 0.2777777910232544), (True, 0.125)):
        OR, prm = _oracle_races(oracle, track, high_mode_mcts=mcts)
        karts, plans = R.start_grid(track, 1, seed=1)
        karts["x"][0], karts["z"][0] = [14.0, 14.0], [0.0, 4.0]
        out = OR.recipe(karts, plans)
        assert out["aw"][0, 0, 0, 0] == out["aw"][0, 0, 0, 1] == want_ego
        assert out["aw"][0, 1, 0, 0] == 0.09615384787321091
        assert out["aw"][1, 0, 0, 0] == want_ego and out["aw"][1, 1, 0, 0] == 0.09615384787321091     # the other agent's own problem


def test_duos_recipe_known_answers():
    """The more-than-two-agents branches of SolveLQR worked out by hand from the C# text for one configuration — four karts, teams [0, 0, 1, 1],
    all within 8 m of the ego (kart 0 at (14, 0); its teammate 4 m away at (14, 4); the opponents 4 m away at (14, -4) and (18, 0)), 10 m/s:
      players [this] + teamAgents + otherAgents = [0, 1, 2, 3] (:702), nearbyAgents = -1 + 4 = 3 (:708-721)
      target weights (:928-947): h = (Fixed ? 2.5 : 3.5) * 3 = 7.5 / 10.5; x = z = 3 * 0.3 * 3.1 / 10 = 0.279; v = 3 * 5e-4 = 0.0015
      control cost (:1192-1196): Fixed 0.135, MCTS 0.25
      avoid weights 1f / (Mathf.Pow(d, 1.5f) * multiplier) (:1019), d = 4 -> d^1.5 = 8:
        ego, opponents:   multiplier (Fixed ? 0.55f : 1.0f) / 3 (:982-985)     -> Fixed 0.6818181872367859, MCTS 0.375
        ego, teammate:    multiplier / 2 (:1113)                               -> Fixed 1.3636363744735718, MCTS 0.75
        another player k: multiplier 1.7f / 3; kart 1's teammate (the ego, 4 m) -> 0.44117647409439087 in its private slot 2 (opponents first)
      opponent-target weights of the ego (:1083-1085): x = z = (Fixed ? 0.1 : 0.2) / (10 * 3), v = 0.08 / 3; teammate-target weights (:1178-1180):
        x = z = -(Fixed ? 0 : 3e-5) / (10 * 3), v = 0
    against oracle/np_recipe.py (the restatement the device recipe is checked against in tests/test_raceN_gpu.py)."""
    from oracle import np_race
    track = S.OVAL
    tt = _track_tables(track)
    for mcts, w_h, cw, a_opp, a_mate, o_xz, m_xz in ((False, 7.5, 0.135, 0.6818181872367859, 1.3636363744735718, 0.1 / 30, -0.0),
                                                     (True, 10.5, 0.25, 0.375, 0.75, 0.2 / 30, -3e-5 / 30)):
        prm = R.race_params(track, high_mode_mcts=mcts)
        karts, plans, beliefs, _ = R.start_grid_n(track, 1, 4, seed=1, teams=[0, 0, 1, 1])
        karts["x"][0], karts["z"][0], karts["v"][0] = [14.0, 14.0, 14.0, 18.0], [0.0, 4.0, -4.0, 0.0], [10.0] * 4
        rec = np_race.recipe_agent(tt, prm, karts[0], plans[0], beliefs[0], 0)
        assert rec["players"] == [0, 1, 2, 3]
        assert np.array_equal(rec["tw"][0], [3 * 0.3 * 3.1 / 10, 3 * 0.3 * 3.1 / 10, 3 * 5e-4, w_h])
        assert np.all(np.asarray(rec["cw"]) == cw)
        assert np.array_equal(rec["aw"][0], [[a_opp, a_opp], [a_opp, a_opp], [a_mate, a_mate]])         # the ego's opponents (karts 2, 3), then its teammate
        assert rec["aw"][1][2][0] == rec["aw"][1][2][1] == 0.44117647409439087                          # kart 1 about its teammate, the ego
        assert np.array_equal(rec["otw"][0][:2], [[o_xz, o_xz, 0.08 / 3]] * 2)
        assert np.array_equal(rec["otw"][0][2], [m_xz, m_xz, 0.0])
    # a kart further than 8 m away drops out of the game (:713) and nearbyAgents falls to Math.Max(1, 1)
    karts["x"][0][3] = 23.0
    rec = np_race.recipe_agent(tt, prm, karts[0], plans[0], beliefs[0], 0)
    assert rec["players"] == [0, 1, 2] and rec["tw"][0][3] == 3.5 * 2


@pytest.mark.parametrize("track", [S.OVAL, S.COMPLEX])
@pytest.mark.parametrize("mcts", [True, False])
@pytest.mark.parametrize("K,teams", [(4, [0, 0, 1, 1]), (4, [0, 1, 0, 1]), (3, [0, 1, 1]), (2, [0, 1])])
def test_duos_recipe_c_restatement_equals_python_restatement(oracle, track, mcts, K, teams):
    """The recipe for any number of agents twice, both from the C# text: oracle/hk_oracle_race.c (hk_oracle_raceN_recipe_one) and
    oracle/np_recipe.py — every field of every real player of 300 x K problems on scattered race states (8 m filter keeping 1..K karts,
    stopped karts, finished karts, random plans and beliefs): weights bit for bit, target headings to 1e-13."""
    from oracle import np_race
    OR, prm = _oracle_races(oracle, track, high_mode_mcts=mcts)
    tt = _track_tables(track)
    rng = np.random.default_rng(100 + K)
    n = 300
    karts, plans, beliefs, _ = R.start_grid_n(track, n, K, seed=9, teams=teams)
    L = track.n_sections
    lanes_xy, head = track.lane_table(), track.heading_table()
    base = rng.integers(0, 2 * L, size=n)
    for e in range(K):
        sec = np.maximum(base + rng.integers(-1, 2, size=n), 0)
        s0 = sec % L
        ln = rng.integers(1, 5, size=n)
        p0, p1 = lanes_xy[s0, ln - 1], lanes_xy[(s0 + 1) % L, ln - 1]
        fr = rng.random(n)[:, None]
        pos = p0 + (p1 - p0) * fr + rng.normal(0, 6.0 * rng.random(n)[:, None], size=(n, 2))
        karts["x"][:, e], karts["z"][:, e] = pos[:, 0], pos[:, 1]
        karts["v"][:, e] = np.where(rng.random(n) < 0.15, rng.uniform(0, 5, n), rng.uniform(5, 15, n))
        karts["h"][:, e] = np.mod(head[s0] + rng.normal(0, 0.2, n), 2 * np.pi)
        karts["section"][:, e], karts["lane"][:, e] = sec, ln
    karts["active"][::41, K - 1] = 0
    for arr in (plans, beliefs):
        on = rng.random(arr["lane"].shape) < 0.6
        arr["lane"][:] = np.where(on, rng.integers(1, 5, size=on.shape), 0)
        arr["vel"][:] = np.where(on, rng.choice([8, 10, 12, 14, 15], size=on.shape), 0)
    seen = set()
    for r in range(n):
        for e in range(K):
            a = np_race.recipe_agent(tt, prm, karts[r], plans[r], beliefs[r], e)
            b = OR.recipe_n_one(K, karts[r], plans[r], beliefs[r, e], e)
            assert a["players"] == b["players"], (r, e)
            N = len(a["players"])
            seen.add(N)
            for k in ("x0", "tw", "cw"):
                assert np.array_equal(np.asarray(a[k]), b[k]), (k, r, e)
            assert np.max(np.abs(np.asarray(a["target"]) - b["target"])) <= 1e-13
            for k in ("aw", "otgt", "otw"):
                assert np.array_equal(np.asarray(a[k]).reshape(b[k].shape), b[k]), (k, r, e, a[k], b[k])
    assert seen == set(range(1, K + 1)) if K > 2 else seen == {2}


def test_recipe_golden_fixture_matches_oracle(oracle):
    """tests/golden/recipe_golden.npz (frozen after the C and the Python restatement agreed, tests/golden/make_golden.py) against the C
    oracle as it is now: a change of either restatement shows up here."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "recipe_golden.npz"))
    for name, track, K in (("duos", S.COMPLEX, 4), ("pair", S.OVAL, 2)):
        karts, plans, beliefs = g[f"{name}_karts"], g[f"{name}_plans"], g[f"{name}_beliefs"]
        for mcts in (False, True):
            OR, _ = _oracle_races(oracle, track, high_mode_mcts=mcts)
            tag = f"{name}_{'mcts' if mcts else 'fixed'}"
            for r in range(karts.shape[0]):
                for e in range(K):
                    c = OR.recipe_n_one(K, karts[r], plans[r], beliefs[r, e], e)
                    N = len(c["players"])
                    assert N == int(g[f"{tag}_n_players"][r, e]) and c["players"] == list(g[f"{tag}_players"][r, e, :N])
                    for k in ("x0", "target", "tw", "cw"):
                        assert np.array_equal(c[k], g[f"{tag}_{k}"][r, e, :N]), (k, r, e)
                    for k in ("aw", "otgt", "otw"):
                        assert np.array_equal(c[k], g[f"{tag}_{k}"][r, e, :N, :max(N - 1, 0)]), (k, r, e)
