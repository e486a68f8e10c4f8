"""GPU parity of the races with up to 4 karts and teams (hk_raceN_*: the Duos scenes' 4 agents in 2 teams) against the CPU oracle
composition oracle/np_race.py (recipe restated from the C# text in oracle/np_recipe.py, the C oracle's providers / LQNG solve with the
REAL player count / plant / bookkeeping, the C oracle's sequential tree search), through the C-ABI.  What is new against the 2-kart
loop: the 8 m nearby filter (N in 1..4 players per problem), nearbyAgents-scaled weights, teammates (multiplier / 2, teammate-target
weights), every player's cost in its private order on the ego's joint order (quirk Q3), the solve every 4th step with held controls,
games of three players solved in the 4-player frame with a decoupled dummy player, games of one or two sent to the 2-kart kernel (a
one-player game with a dummy second player), team scoring in the tree search."""
import numpy as np
import pytest

from conftest import rel_err
from hierarchicalkarting_b200 import abi, race as R, scenarios as S
from test_race_cpu import _oracle_races, _track_tables
from test_race_gpu import INT_FIELDS, DBL_FIELDS

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _scatter(track, n, K, seed, spread=6.0):
    """Race states with the karts of a race spread so that the 8 m filter keeps 1..K of them."""
    rng = np.random.default_rng(seed)
    karts, plans, beliefs, u = R.start_grid_n(track, n, K, seed=seed)
    L = track.n_sections
    lanes_xy, head = track.lane_table(), track.heading_table()
    base = rng.integers(0, 2 * L, size=n)
    for e in range(K):
        sec = np.maximum(base + rng.integers(-1, 2, size=n), 0)
        s0 = sec % L
        ln = rng.integers(1, 5, size=n)
        p0, p1 = lanes_xy[s0, ln - 1], lanes_xy[(s0 + 1) % L, ln - 1]
        fr = rng.random(n)[:, None]
        pos = p0 + (p1 - p0) * fr + rng.normal(0, spread * rng.random(n)[:, None], size=(n, 2))
        karts["x"][:, e], karts["z"][:, e] = pos[:, 0], pos[:, 1]
        karts["v"][:, e] = np.where(rng.random(n) < 0.15, rng.uniform(0, 5, n), rng.uniform(5, 15, n))
        karts["h"][:, e] = np.mod(head[s0] + rng.normal(0, 0.2, n), 2 * np.pi)
        karts["section"][:, e], karts["lane"][:, e] = sec, ln
    karts["active"][::53, K - 1] = 0
    on = rng.random(plans["lane"].shape) < 0.6
    plans["lane"][:] = np.where(on, rng.integers(1, 5, size=on.shape), 0)
    plans["vel"][:] = np.where(on, rng.choice([8, 10, 12, 14, 15], size=on.shape), 0)
    on = rng.random(beliefs["lane"].shape) < 0.6
    beliefs["lane"][:] = np.where(on, rng.integers(1, 5, size=on.shape), 0)
    beliefs["vel"][:] = np.where(on, rng.choice([8, 10, 12, 14, 15], size=on.shape), 0)
    return karts, plans, beliefs, u


@pytest.mark.parametrize("track", [S.OVAL, S.COMPLEX])
@pytest.mark.parametrize("mcts", [True, False])
@pytest.mark.parametrize("K,teams", [(4, [0, 0, 1, 1]), (3, [0, 1, 1])])
def test_recipeN_parity(hk, oracle, track, mcts, K, teams):
    """Every field of every real player's description, the player lists and counts, against oracle/np_recipe.py; N takes every value 1..K."""
    from oracle import np_race
    prm = R.race_params(track, high_mode_mcts=mcts)
    G = R.RacesN(track, prm, K)
    tt = _track_tables(track)
    n = 400
    karts, plans, beliefs, _ = _scatter(track, n, K, seed=7 + K)
    karts["team"][:] = teams
    got = G.recipe_n(karts, plans, beliefs)
    seen = set()
    for r in range(n):
        for e in range(K):
            rec = np_race.recipe_agent(tt, prm, karts[r], plans[r], beliefs[r], e)
            b = r * K + e
            N = len(rec["players"])
            seen.add(N)
            assert int(got["n_players"][b]) == N and list(got["players"][b][:N]) == rec["players"] and np.all(got["players"][b][N:] == -1)
            for k in ("x0", "tw"):
                assert np.array_equal(got[k][b, :N], rec[k]), (k, r, e)
            assert np.array_equal(got["cw"][b, :N], rec["cw"]) and np.all(got["cw"][b, N:] == 1.0)
            assert np.max(np.abs(got["target"][b, :N] - rec["target"])) <= 1e-13
            for k in ("aw", "otgt", "otw"):
                assert np.array_equal(got[k][b, :N, :N - 1], rec[k]), (k, r, e, got[k][b, :N, :N - 1], rec[k])
                assert not got[k][b, :N, N - 1:].any() and not got[k][b, N:].any()       # unused private slots and dummy players: zero
            assert not got["tw"][b, N:].any() and not got["x0"][b, N:].any()
    assert seen == set(range(1, K + 1))


def test_two_kart_races_through_the_n_kart_path_equal_the_two_kart_path(hk):
    """K = 2 (no 8 m filter, each kart its own team): hk_raceN_recipe gives the 2-kart recipe, and the loop — recipe in the 4-player layout,
    repacked for the 2-kart kernel, every 4th-step machinery at lqr_every = 1 — drives the karts the same way."""
    track = S.OVAL
    prm = R.race_params(track)
    G2, GN = R.Races(track, prm), R.RacesN(track, prm, 2, lqr_every=1)
    n = 300
    karts, plans = R.start_grid(track, n, seed=3)
    karts["team"][:, 1] = 1
    G2.plan_fixed(karts, plans)
    G2.run(karts, plans, 0, 60)
    beliefs = np.zeros((n, 2, 2), dtype=abi.RACE_BELIEF_DTYPE)
    for e in range(2):
        beliefs["lane"][:, e, 1 - e], beliefs["vel"][:, e, 1 - e] = plans["oppLane"][:, e], plans["oppVel"][:, e]
    a, b = G2.recipe(karts, plans), GN.recipe_n(karts, plans, beliefs)
    assert np.all(b["n_players"] == 2)
    for k in ("x0", "target", "tw", "cw"):
        assert np.array_equal(a[k], b[k][:, :2]), k
    for k in ("aw", "otgt", "otw"):
        assert np.array_equal(a[k], b[k][:, :2, :1]), k
    k2, p2, kn, pn = karts.copy(), plans.copy(), karts.copy(), plans.copy()
    u = np.zeros((n, 2, 2))
    _, bad2 = G2.run(k2, p2, 60, 40)
    badn = GN.run_n(kn, pn, beliefs, u, 60, 40)
    assert bad2 == badn == 0
    for f in INT_FIELDS:
        assert np.array_equal(k2[f], kn[f]), f
    for f in DBL_FIELDS:
        assert np.max(np.abs(k2[f] - kn[f])) <= 1e-9 * max(1.0, float(np.max(np.abs(k2[f])))), f


def _same_n(gk, gp, gb, gu, ok, op, ob, ou):
    for f in INT_FIELDS + ("team",):
        assert np.array_equal(gk[f], ok[f]), f
    for f in DBL_FIELDS:
        assert np.max(np.abs(gk[f] - ok[f])) <= TOL * max(1.0, float(np.max(np.abs(ok[f])))), f
    for f in ("lane", "vel", "sectionTimes", "lapStep"):
        assert np.array_equal(gp[f], op[f]), f
    for f in ("lane", "vel"):
        assert np.array_equal(gb[f], ob[f]), f
    assert rel_err(gu, ou) <= TOL


@pytest.mark.parametrize("track,K,teams", [(S.COMPLEX, 4, [0, 0, 1, 1]), (S.OVAL, 4, [0, 1, 0, 1]), (S.OVAL, 3, [0, 0, 1])])
def test_fixed_mode_loop_parity(hk, oracle, track, K, teams):
    """BASELINE config 3's game inside a loop (Duos, Fixed high level): start grid lanes {2,3,2,3} at sections {0,0,1,1}, planFixed every
    100 steps, the LQNG problem of every agent — N in 1..4 after the 8 m filter — every 4th step, held controls, plant, bookkeeping; device
    against the oracle composition, re-synchronised every 40 steps."""
    from oracle import np_race
    OR, prm = _oracle_races(oracle, track)
    G = R.RacesN(track, prm, K)
    tt = _track_tables(track)
    n = 6
    karts, plans, beliefs, u = R.start_grid_n(track, n, K, seed=17, teams=teams)
    G.plan_fixed(karts, plans)
    counts = set()
    for blk in range(6):
        gk, gp, gb, gu = karts.copy(), plans.copy(), beliefs.copy(), u.copy()
        bad_g = G.run_n(gk, gp, gb, gu, blk * 40, 40)
        bad_o = np_race.run_n(OR, tt, K, G.lqr_every, karts, plans, beliefs, u, blk * 40, 40)
        assert bad_g == bad_o == 0
        _same_n(gk, gp, gb, gu, karts, plans, beliefs, u)
        counts |= set(int(x) for x in G.recipe_n(karts, plans, beliefs)["n_players"])
    assert karts["section"].min() >= 2 and len(counts) >= 2              # the filter changed the games' sizes along the way


def test_mcts_duos_loop_parity(hk, oracle):
    """The full hierarchical loop for 2v2 races: every agent's sequential tree search over the 4-kart game WITH TEAM SCORING
    (KartDiscreteGame.cs:271-310, teamScoreRewardMultiplier 0.75) on the device, root states of all agents within the section window,
    beliefs about the three other karts handed off, trees kept / dropped by the planner's schedule — against the oracle composition."""
    from hierarchicalkarting_b200 import mcts as M, tracks
    from oracle import np_race
    track, K, teams = S.COMPLEX, 4, [0, 0, 1, 1]
    OR, prm = _oracle_races(oracle, track, high_mode_mcts=True)
    prm.planEvery = 40
    G = R.RacesN(track, prm, K)
    tt = _track_tables(track)
    game = M.Game(track, 4, prm.velocityBucketSize)
    gparams = tracks.game_params(track, bucket=prm.velocityBucketSize)
    OG = oracle.Game(track.sections_array(), track.n_sections, tracks.kart_array(4), 4, gparams)
    n, iters, seed = 4, 20, 99
    karts, plans, beliefs, u = R.start_grid_n(track, n, K, seed=23, teams=teams)
    gpl = G.planner(game, n, iters, seed, mode=0, first_iterations=28, reuse_cycles=3, apply_delay=9, max_tree_nodes=1 + 32 * (28 + 12 * iters))
    opl = np_race.PlannerN(OG, gparams, n, K, iters, seed, first_iterations=28, reuse_cycles=3, apply_delay=9)
    for blk in range(5):
        gk, gp, gb, gu = karts.copy(), plans.copy(), beliefs.copy(), u.copy()
        bad_g = G.run_n(gk, gp, gb, gu, blk * 40, 40, planner=gpl)
        bad_o = np_race.run_n(OR, tt, K, G.lqr_every, karts, plans, beliefs, u, blk * 40, 40, planner=opl)
        assert bad_g == bad_o == 0
        _same_n(gk, gp, gb, gu, karts, plans, beliefs, u)
        rv, cy, ts = gpl.state()
        assert np.array_equal(rv.reshape(-1), opl.root_valid) and np.array_equal(cy.reshape(-1), opl.cycles) and not ts.any()
    assert (plans["lane"] != 0).any() and (beliefs["lane"] != 0).any() and karts["section"].min() >= 2


def test_recipe_golden_fixture_on_device(hk):
    """The frozen recipe outputs of tests/golden/recipe_golden.npz (C oracle, cross-checked against the Python restatement before
    freezing) against hk_raceN_recipe: Duos and 2-kart races, both high-level modes, no oracle loaded."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "recipe_golden.npz"))
    for name, track, K in (("duos", S.COMPLEX, 4), ("pair", S.OVAL, 2)):
        karts, plans, beliefs = (np.ascontiguousarray(g[f"{name}_{k}"]) for k in ("karts", "plans", "beliefs"))
        n = karts.shape[0]
        for mcts in (False, True):
            G = R.RacesN(track, R.race_params(track, high_mode_mcts=mcts), K)
            got = G.recipe_n(karts, plans, beliefs)
            tag = f"{name}_{'mcts' if mcts else 'fixed'}"
            assert np.array_equal(got["n_players"].reshape(n, K), g[f"{tag}_n_players"])
            assert np.array_equal(got["players"].reshape(n, K, 4), g[f"{tag}_players"])
            for k in ("x0", "tw", "cw", "aw", "otgt", "otw"):
                assert np.array_equal(got[k].reshape(g[f"{tag}_{k}"].shape), g[f"{tag}_{k}"]), k
            assert np.max(np.abs(got["target"].reshape(g[f"{tag}_target"].shape) - g[f"{tag}_target"])) <= 1e-13
