"""GPU parity of the closed loop without PhysX (hk_race.cu) against the CPU oracle (oracle/hk_oracle_race.c), through the
C-ABI.  Integers (section, lane, counters, plan keys) bit-exact; doubles within 1e-9 (BASELINE.json north_star tolerance),
in practice ~1e-15 (device vs host libm)."""
import numpy as np
import pytest

from conftest import rel_err
from hierarchicalkarting_b200 import abi, race as R, scenarios as S
from test_race_cpu import _oracle_races, karts_from_problems

pytestmark = pytest.mark.gpu
TOL = 1e-9
INT_FIELDS = ("section", "lane", "laneChanges", "illegalLaneChanges", "sectionStep", "active")
DBL_FIELDS = ("x", "z", "v", "h")


def _same(gk, gp, ok, op, tol=TOL):
    for f in INT_FIELDS:
        assert np.array_equal(gk[f], ok[f]), f
    for f in DBL_FIELDS:
        assert np.max(np.abs(gk[f] - ok[f])) <= tol * max(1.0, float(np.max(np.abs(ok[f])))), f
    for f in ("lane", "vel", "oppLane", "oppVel", "sectionTimes", "lapStep"):
        assert np.array_equal(gp[f], op[f]), f
    for f in ("avgLaneDiff", "avgVelDiff"):
        assert np.max(np.abs(gp[f] - op[f])) <= 1e-6 * max(1.0, float(np.max(np.abs(op[f])))), f


@pytest.mark.parametrize("track", [S.OVAL, S.COMPLEX])
@pytest.mark.parametrize("mcts", [True, False])
def test_recipe_parity(hk, oracle, track, mcts):
    p = S.make_problems(track, 2001, 2, seed=17, high_mode_mcts=mcts)          # ragged size
    karts, plans = karts_from_problems(p, track)
    plans[:, 1] = plans[:, 0]                                                   # give the second agent a plan too
    OR, prm = _oracle_races(oracle, track, high_mode_mcts=mcts)
    G = R.Races(track, prm)
    ref, got = OR.recipe(karts, plans), G.recipe(karts, plans)
    for k in ("x0", "tw", "cw", "aw", "otgt", "otw"):
        assert np.array_equal(got[k], ref[k]), k
    assert np.max(np.abs(got["target"] - ref["target"])) <= 1e-13
    # and the assembled problems solve to the same controls
    from hierarchicalkarting_b200 import lqr
    u_g = lqr.assemble_solve_batch(got, 3)["u0"]
    u_o = oracle.lqng_solve_batch(*S.assemble_dense(ref), 3, full=False)["u0"]
    for b in range(0, u_o.shape[0], 7):
        assert rel_err(u_g[b], u_o[b]) <= TOL


def test_plan_fixed_and_step_parity(hk, oracle):
    rng = np.random.default_rng(5)
    for track in (S.OVAL, S.COMPLEX):
        OR, prm = _oracle_races(oracle, track)
        G = R.Races(track, prm)
        n = 4097
        karts, plans = R.start_grid(track, n, seed=9)
        L = track.n_sections
        # scatter the karts just before / after / far from their next gate, with every kind of control
        sec = rng.integers(0, 3 * L, size=(n, 2))
        c = (sec + 1) % L
        trig, head = track.trigger_table(), track.heading_table()
        fwd = np.stack([np.cos(head), np.sin(head)], axis=-1)
        along = rng.choice([-0.05, -0.15, -0.4, 0.1, -3.0], size=(n, 2))
        lat = rng.uniform(-12.0, 12.0, size=(n, 2))
        karts["x"] = trig[c, 0] + fwd[c, 0] * along - fwd[c, 1] * lat
        karts["z"] = trig[c, 1] + fwd[c, 1] * along + fwd[c, 0] * lat
        karts["v"] = rng.uniform(0.0, 15.0, size=(n, 2))
        karts["h"] = np.mod(head[c] + rng.normal(0, 0.3, size=(n, 2)), 2 * np.pi)
        karts["section"], karts["lane"] = sec, rng.integers(1, 5, size=(n, 2))
        karts["laneChanges"] = rng.integers(0, 5, size=(n, 2))
        karts["active"] = rng.random((n, 2)) < 0.95
        karts["section"][0, 0] = prm.goalSection - 1                            # one kart finishes
        plans["lane"] = rng.integers(0, 5, size=plans["lane"].shape)
        plans["vel"] = rng.integers(6, 16, size=plans["vel"].shape)
        u = rng.normal(0, 3, size=(n, 2, 2))
        u[rng.random((n, 2)) < 0.2, 0] = 0.0                                    # coasting branch
        gk, gp, ok, op = karts.copy(), plans.copy(), karts.copy(), plans.copy()
        G.plan_fixed(gk, gp); OR.plan_fixed(ok, op)
        _same(gk, gp, ok, op)
        assert not np.array_equal(gp["lane"], plans["lane"])
        G.step(gk, gp, u, 123); OR.step(ok, op, u, 123)
        _same(gk, gp, ok, op, tol=1e-13)
        assert (ok["section"] != karts["section"]).mean() > 0.1                 # many crossings happened
        assert (ok["illegalLaneChanges"] > 0).any() and (ok["laneChanges"] != karts["laneChanges"]).any()


@pytest.mark.parametrize("track", [S.OVAL, S.COMPLEX])
def test_run_parity(hk, oracle, track):
    """The whole loop (replan, recipe, assemble, solve, actuator, plant, bookkeeping) for 96 races x 600 steps, re-synchronised
    with the oracle every 100 steps so that rounding differences cannot move a checkpoint crossing to another step."""
    OR, prm = _oracle_races(oracle, track)
    G = R.Races(track, prm)
    karts, plans = R.start_grid(track, 96, seed=23)
    for blk in range(6):
        gk, gp = karts.copy(), plans.copy()
        u_g, bad_g = G.run(gk, gp, blk * 100, 100)
        u_o, bad_o = OR.run(karts, plans, blk * 100, 100)
        assert bad_g == bad_o == 0
        _same(gk, gp, karts, plans, tol=1e-9)
        assert np.max(np.abs(u_g - u_o)) <= 1e-8 * max(1.0, float(np.max(np.abs(u_o))))
    assert karts["section"].min() >= 8


def test_full_size_properties(hk):
    """BASELINE config 5 size (16,384 Oval races): determinism, monotone progress, bounded states, every solve regular."""
    track = S.OVAL
    G = R.Races(track, R.race_params(track))
    karts, plans = R.start_grid(track, 16384, seed=20260004)
    a_k, a_p = karts.copy(), plans.copy()
    _, bad = G.run(a_k, a_p, 0, 300)
    b_k, b_p = karts.copy(), plans.copy()
    G.run(b_k, b_p, 0, 150)
    G.run(b_k, b_p, 150, 150)                                                   # split runs give the same races
    assert bad == 0 and np.array_equal(a_k, b_k) and np.array_equal(a_p, b_p)
    assert np.all(a_k["section"] >= 3) and np.all(a_k["section"] <= 8)
    assert np.all(np.isfinite(a_k["x"])) and np.all((a_k["v"] >= 0) & (a_k["v"] <= 15)) and np.all((a_k["h"] >= 0) & (a_k["h"] < 2 * np.pi))
    perm = np.random.default_rng(0).permutation(16384)
    c_k, c_p = np.ascontiguousarray(karts[perm]), np.ascontiguousarray(plans[perm])
    G.run(c_k, c_p, 0, 300)
    assert np.array_equal(c_k, a_k[perm])                                       # races are independent


def test_mcts_root_and_waypoint_handoff(hk):
    """planWithMCTS's root state (HKA:180-245), the host tree policy over GPU leaf statistics and the
    hand-off of getBestStatesSequence to the LQNG targets (HKA:366-402), closing the MCTS -> LQNG loop."""
    from hierarchicalkarting_b200 import mcts as M
    track = S.OVAL
    prm = R.race_params(track, high_mode_mcts=True)
    G = R.Races(track, prm)
    game = M.Game(track, 2, prm.velocityBucketSize)
    karts, plans = R.start_grid(track, 2, seed=31)
    G.run(karts, plans, 0, 99)                                      # no plan yet: both karts steer for the Trigger centres
    assert np.all(plans["lane"] == 0) and karts["section"].min() >= 1
    # root state of race 0, ego 0
    st, nearby = R.mcts_root_state(track, prm, karts[0], plans[0], 0)
    assert nearby == [0, 1] and st.n_karts == 2 and st.initialSection == int(karts[0]["section"].max())
    assert st.finalSection == st.initialSection + 8 and st.karts[0].tireAge == 2500 and st.karts[0].max_velocity == 2
    lead, trail = int(np.argmax(karts[0]["section"])), int(np.argmin(karts[0]["section"]))
    if karts[0]["section"][lead] != karts[0]["section"][trail]:
        d = int(plans[0]["sectionTimes"][trail][karts[0]["section"][trail] % 24]) - int(plans[0]["sectionTimes"][lead][karts[0]["section"][trail] % 24])
        assert st.karts[trail].timeAtSection == int(np.float32(np.float32(d) * np.float32(0.02)) * np.float32(100)) > 0
    M.KartMCTS.rollouts_per_leaf = 256
    for r in range(2):
        for ego in range(2):
            root, best = R.plan_with_mcts(track, prm, game, karts[r], plans[r], ego, max_iterations=60, seed=100 + 2 * r + ego)
            assert root.numEpisodes > 0 and len(best) >= 1
            sec = int(karts[r]["section"][ego])
            keys = [int(b.state.karts[0].section) for b in best]
            mine = [k for k in keys if k > sec + (0 if sec == 0 else 1)]
            for k in mine:
                assert 1 <= plans[r]["lane"][ego][k % 24] <= 4 and plans[r]["vel"][ego][k % 24] in (8, 10, 12, 14, 15)
            assert (plans[r]["oppLane"][ego] != 0).sum() >= len(best) - 1
            assert plans[r]["lane"][ego][(sec + 1) % 24] == 0           # the next checkpoint keeps its Trigger target (:369)
    before = karts["section"].copy()
    _, bad = G.run(karts, plans, 99, 300)                           # LQNG now tracks the MCTS waypoints (+2 buckets of speed, :757)
    assert bad == 0 and np.all(karts["section"] >= before + 4)


def test_batched_mcts_planning_equals_per_agent_host_search(hk):
    """BASELINE config 5's high level at scale: race.plan_mcts_batch (every agent's tree in one hk_mcts_search_batch launch, one thread
    block per tree) leaves the plan tables the per-agent host flow (planWithMCTS over the KartMCTS mirror) leaves when both are
    driven by the same Philox streams; the races then run on those waypoints."""
    from hierarchicalkarting_b200 import mcts as M
    track = S.OVAL
    prm = R.race_params(track, high_mode_mcts=True)
    G = R.Races(track, prm)
    game = M.Game(track, 2, prm.velocityBucketSize)
    n_races, K, RPL, seed = 6, 24, 32, 777
    karts, plans = R.start_grid(track, n_races, seed=41)
    G.run(karts, plans, 0, 100)
    pa, pb = plans.copy(), plans.copy()
    out = R.plan_mcts_batch(track, prm, game, karts, pa, K, RPL, seed)
    assert out["n_best"].max() >= 1
    old_random, old_R = M.KartMCTS.random, M.KartMCTS.rollouts_per_leaf
    try:
        M.KartMCTS.rollouts_per_leaf = RPL
        for r in range(n_races):
            for ego in range(2):
                M.KartMCTS.random = M.PhiloxPicks(seed + 2 * r + ego)
                root, best = R.plan_with_mcts(track, prm, game, karts[r], pb[r], ego, T=1e9, max_iterations=K, seed=seed + 2 * r + ego, parallel=True)
                assert len(best) == int(out["n_best"][2 * r + ego])
    finally:
        M.KartMCTS.random, M.KartMCTS.rollouts_per_leaf = old_random, old_R
    for f in ("lane", "vel", "oppLane", "oppVel"):
        assert np.array_equal(pa[f], pb[f]), f
    before = karts["section"].copy()
    _, bad = G.run(karts, pa, 100, 300)                             # MCTS mode: hk_race_run does not replan with planFixed
    assert bad == 0 and np.all(karts["section"] >= before + 4)


def test_device_resident_mcts_loop_equals_host_composition(hk):
    """hk_race_run_mcts (root states, tree search and waypoint hand-off as kernels between two steps) leaves exactly the karts and plan
    tables of the host composition run / plan_mcts_batch / run / ... with the seeds shifted per planning event."""
    from hierarchicalkarting_b200 import mcts as M
    track = S.OVAL
    prm = R.race_params(track, high_mode_mcts=True)
    G = R.Races(track, prm)
    game = M.Game(track, 2, prm.velocityBucketSize)
    n_races, K, RPL, seed = 50, 24, 16, 4242
    ka, pa = R.start_grid(track, n_races, seed=43)
    kb, pb = ka.copy(), pa.copy()
    _, bad = G.run_mcts(ka, pa, game, K, RPL, seed, 0, 300)
    assert bad == 0
    for ev in range(3):
        G.run(kb, pb, 100 * ev, 100)
        if ev < 2:
            R.plan_mcts_batch(track, prm, game, kb, pb, K, RPL, seed + (ev + 1) * 2 * n_races)   # event of step 100 (ev + 1): key = seed + (step / planEvery) n_agents + agent
    for f in ka.dtype.names:
        assert np.array_equal(ka[f], kb[f]), f
    for f in pa.dtype.names:
        assert np.array_equal(pa[f], pb[f]), f
    assert ((pa["lane"] != 0) | (pa["oppLane"] != 0)).any() and ka["section"].min() >= 4     # waypoints were handed off, the karts drove on


def test_mcts_loop_parity_against_cpu_oracle(hk, oracle):
    """BASELINE config 5 with the LEAF-PARALLEL search — MCTS waypoints -> LQNG -> dynamics — against a loop in which nothing comes from
    the CUDA library or its host mirrors: the C oracle's race loop (recipe, LQNG, plant, bookkeeping), the C oracle's restatement of
    planWithMCTS's root state and of the hand-off (hk_oracle_race_mcts_root / _apply_best) and oracle/np_mcts.py's tree search over the
    C oracle's game.  Re-synchronised every 100 steps like test_run_parity, so that rounding differences cannot move a checkpoint
    crossing to another step."""
    from hierarchicalkarting_b200 import mcts as M, tracks
    from oracle import np_mcts
    track = S.OVAL
    OR, prm = _oracle_races(oracle, track, high_mode_mcts=True)
    G = R.Races(track, prm)
    game = M.Game(track, 2, prm.velocityBucketSize)
    gparams = tracks.game_params(track, bucket=prm.velocityBucketSize)
    OG = oracle.Game(track.sections_array(), track.n_sections, tracks.kart_array(2), 2, gparams)
    n_races, K, RPL, seed = 5, 24, 16, 9001
    karts, plans = R.start_grid(track, n_races, seed=29)
    waypoints = 0
    for blk in range(4):
        gk, gp = karts.copy(), plans.copy()
        _, bad_g = G.run_mcts(gk, gp, game, K, RPL, seed, blk * 100, 100)                    # plans at its first step when blk > 0
        if blk > 0:
            for r in range(n_races):
                snapshot = karts[r].copy()                                                   # both agents plan from the same race state
                for ego in range(2):
                    st, nearby = OR.mcts_root(gparams, snapshot, plans[r], ego)
                    _, best, _ = np_mcts.TreeSearch(OG, seed + blk * 2 * n_races + 2 * r + ego).search(st, K, RPL)
                    OR.apply_best(snapshot, plans[r], ego, nearby, best)
                    waypoints += len(best)
        _, bad_o = OR.run(karts, plans, blk * 100, 100)
        assert bad_g == bad_o == 0
        _same(gk, gp, karts, plans, tol=1e-9)
    assert waypoints > 0 and karts["section"].min() >= 5


@pytest.mark.parametrize("plan_every,delay,first,reuse", [(100, 0, 40, 3), (100, 45, 40, 3), (20, 0, 0, 3), (20, 7, 30, 2), (25, 0, 16, 0)])
def test_planned_loop_parity_against_cpu_oracle(hk, oracle, plan_every, delay, first, reuse):
    """BASELINE config 5 as the reference runs it: the FAITHFUL sequential search with the planner's whole schedule — plan at episode
    start (T = 1.5), a new tree or a continued one while CyclesRootProcessed < 3 or no plan at all, trees dropped at checkpoint crossings,
    results landing `delay` steps after the search started (HierarchicalKartAgent.cs:85-93, 172-283, 331-353, 366-402, 660-661) — on the
    device (hk_race_run_planned) against the C oracle's planned loop (hk_oracle_race_run_planned: its own root states, hand-off, schedule
    and tree search).  Both sides keep their planners across the blocks; karts / plans are re-synchronised every block.  planEvery = 20
    makes several planning events fall between two checkpoint crossings, so trees ARE continued and the 'reused enough' branch is taken."""
    from hierarchicalkarting_b200 import mcts as M, tracks
    track = S.OVAL
    OR, prm = _oracle_races(oracle, track, high_mode_mcts=True)
    prm.planEvery = plan_every
    G = R.Races(track, prm)
    game = M.Game(track, 2, prm.velocityBucketSize)
    gparams = tracks.game_params(track, bucket=prm.velocityBucketSize)
    OG = oracle.Game(track.sections_array(), track.n_sections, tracks.kart_array(2), 2, gparams)
    n_races, iters, seed = 7, 24, 555
    karts, plans = R.start_grid(track, n_races, seed=31)
    karts["active"][3, 1] = 0                                                                # an inactive agent never plans (:331)
    gpl = R.Planner(game, n_races, iters, seed, mode=0, first_iterations=first, reuse_cycles=reuse, apply_delay=delay,
                    max_tree_nodes=1 + 16 * (first + 14 * iters))   # a crossing during a continued search resets CyclesRootProcessed: no bound short of every event
    opl = OR.planner(OG, gparams, n_races, iters, seed, first_iterations=first, reuse_cycles=reuse, apply_delay=delay)
    continued = stalled = 0
    block = 50
    for blk in range(6):
        gk, gp = karts.copy(), plans.copy()
        _, bad_g = G.run_planned(gk, gp, gpl, blk * block, block)
        _, bad_o = opl.run(karts, plans, blk * block, block)
        assert bad_g == bad_o == 0
        _same(gk, gp, karts, plans, tol=1e-9)
        (rv_g, cy_g, ts_g), (rv_o, cy_o) = gpl.state(), opl.state()
        assert np.array_equal(rv_g, rv_o) and np.array_equal(cy_g, cy_o) and not ts_g.any()
        continued += int((cy_o >= 2).sum())
        stalled += int((cy_o >= max(reuse, 1)).sum())
    assert ((plans["lane"] != 0) | (plans["oppLane"] != 0)).any() and karts["section"][karts["active"] == 1].min() >= 2
    if plan_every == 20 and reuse > 0:
        assert continued > 0                                                                 # constructSearchTree(currentRoot) was exercised


def test_device_resident_entries_equal_the_host_pointer_entries(hk):
    """hk_race_run_device / hk_raceN_run_device (race states that stay in device memory, advanced in blocks, on a caller's stream) against
    hk_race_run / hk_race_run_planned / hk_raceN_run on the same states: identical bytes, the planner's state included."""
    import torch
    from hierarchicalkarting_b200 import mcts as M
    track = S.OVAL
    prm = R.race_params(track)
    G = R.Races(track, prm)
    n = 500
    karts, plans = R.start_grid(track, n, seed=41)
    G.plan_fixed(karts, plans)
    hk_, hp_ = karts.copy(), plans.copy()
    u_host, bad_h = G.run(hk_, hp_, 0, 230)
    dk, dp = R.device_state(karts, plans)
    du = torch.zeros((2 * n, 4), dtype=torch.float64, device="cuda:0")
    stream = torch.cuda.Stream()
    bad_d = 0
    for a, b in ((0, 70), (70, 130), (200, 30)):                       # advanced in blocks; plans at steps 100 and 200 fall inside / on a block edge
        bad_d += G.run_device(dk, dp, a, b, d_u=du, stream=stream.cuda_stream)
    assert bad_d == bad_h == 0
    assert R.host_state(dk, karts).tobytes() == hk_.tobytes() and R.host_state(dp, plans).tobytes() == hp_.tobytes()
    assert np.array_equal(du.cpu().numpy()[:, :2].reshape(n, 2, 2), u_host)
    # MCTS planner: device-resident blocks against one host-pointer call
    prm_m = R.race_params(track, high_mode_mcts=True)
    prm_m.planEvery = 40
    Gm = R.Races(track, prm_m)
    game = M.Game(track, 2, prm_m.velocityBucketSize)
    karts, plans = R.start_grid(track, 64, seed=42)
    kw = dict(mode=0, first_iterations=24, reuse_cycles=3, apply_delay=9, max_tree_nodes=1 + 16 * (24 + 12 * 30))
    p1, p2 = R.Planner(game, 64, 30, seed=5, **kw), R.Planner(game, 64, 30, seed=5, **kw)
    hk_, hp_ = karts.copy(), plans.copy()
    Gm.run_planned(hk_, hp_, p1, 0, 170)
    dk, dp = R.device_state(karts, plans)
    for a, b in ((0, 45), (45, 80), (125, 45)):                        # a search in flight across a block edge (lands 9 steps after 40, 80, ...)
        Gm.run_device(dk, dp, a, b, planner=p2)
    assert R.host_state(dk, karts).tobytes() == hk_.tobytes() and R.host_state(dp, plans).tobytes() == hp_.tobytes()
    for x, y in zip(p1.state(), p2.state()):
        assert np.array_equal(x, y)
    # Duos
    K = 4
    GN = R.RacesN(S.COMPLEX, R.race_params(S.COMPLEX), K)
    karts, plans, beliefs, u = R.start_grid_n(S.COMPLEX, 200, K, seed=43)
    GN.plan_fixed(karts, plans)
    hk_, hp_, hb_, hu_ = karts.copy(), plans.copy(), beliefs.copy(), u.copy()
    GN.run_n(hk_, hp_, hb_, hu_, 0, 150)
    dk, dp, db = R.device_state(karts, plans, beliefs)
    du8 = torch.zeros((200 * K, 8), dtype=torch.float64, device="cuda:0")
    for a, b in ((0, 50), (50, 3), (53, 97)):                          # block edges off the every-4th-step solve: the held controls carry over
        GN.run_n_device(dk, dp, db, du8, a, b)
    assert R.host_state(dk, karts).tobytes() == hk_.tobytes() and R.host_state(dp, plans).tobytes() == hp_.tobytes()
    assert R.host_state(db, beliefs).tobytes() == hb_.tobytes()
    assert np.array_equal(du8.cpu().numpy()[:, :2].reshape(200, K, 2), hu_)
    # Duos with the MCTS planner (team scoring, beliefs handed off): device-resident blocks against one host-pointer call
    prm4 = R.race_params(S.COMPLEX, high_mode_mcts=True)
    prm4.planEvery = 40
    GM = R.RacesN(S.COMPLEX, prm4, K)
    game4 = M.Game(S.COMPLEX, K, prm4.velocityBucketSize)
    karts, plans, beliefs, u = R.start_grid_n(S.COMPLEX, 24, K, seed=44, teams=[0, 0, 1, 1])
    mk = lambda: GM.planner(game4, 24, 16, 9, mode=0, first_iterations=12, reuse_cycles=3, apply_delay=9, max_tree_nodes=1 + 32 * (12 + 12 * 16))
    q1, q2 = mk(), mk()
    hk_, hp_, hb_, hu_ = karts.copy(), plans.copy(), beliefs.copy(), u.copy()
    GM.run_n(hk_, hp_, hb_, hu_, 0, 130, planner=q1)
    dk, dp, db = R.device_state(karts, plans, beliefs)
    du8 = torch.zeros((24 * K, 8), dtype=torch.float64, device="cuda:0")
    for a, b in ((0, 44), (44, 41), (85, 45)):
        GM.run_n_device(dk, dp, db, du8, a, b, planner=q2)
    assert R.host_state(dk, karts).tobytes() == hk_.tobytes() and R.host_state(dp, plans).tobytes() == hp_.tobytes()
    assert R.host_state(db, beliefs).tobytes() == hb_.tobytes() and (hb_["lane"] != 0).any()
    for x, y in zip(q1.state(), q2.state()):
        assert np.array_equal(x, y)
