"""GPU parity of the discrete race game and the rollout kernels against the CPU oracle, through the C-ABI.
Transitions, legal move sets, policy ordering and terminal scores must be BIT-EXACT (BASELINE.json north_star);
rollout decision statistics must match the reference's own sampling procedure distributionally."""
import ctypes as C

import numpy as np
import pytest

from hierarchicalkarting_b200 import abi, mcts, tracks

pytestmark = pytest.mark.gpu


def _oracle_game(oracle, track, n_karts, bucket, tp=100):
    params = tracks.game_params(track, bucket, tp)
    return oracle.Game(track.sections_array(), track.n_sections, tracks.kart_array(n_karts), n_karts, params)


def _random_root(rng, track, n_karts, bucket, teams):
    sec = int(rng.integers(0, 3 * track.n_sections))
    lanes = [int(x) for x in rng.integers(1, 5, n_karts)]
    if rng.random() < 0.5:
        buckets = None                                                     # reference root quirk: (0, bucket)
    else:
        buckets = []
        for _ in range(n_karts):
            v = int(rng.choice(list(range(6, 15, bucket))))
            buckets.append((v, min(v + bucket, 15)))
    times = [0] + [int(x) for x in rng.integers(0, 151, n_karts - 1)]
    st = tracks.root_state(track, sec, lanes, teams=teams, buckets=buckets, tire_age=int(rng.choice([0, 2500, 6000, 9900, 13500])),
                           lane_changes=int(rng.integers(0, 3)), times=times)
    if buckets is None:
        for i in range(n_karts):
            st.karts[i].max_velocity = bucket
    return st


def _state_tuple_np(rec):
    n = int(rec["n_karts"])
    return (n, int(rec["initialSection"]), int(rec["lastCompletedSection"]), int(rec["finalSection"]),
            tuple(tuple(int(rec["karts"][i][k]) for k in abi.KART_STATE_FIELDS) for i in range(n)))


@pytest.mark.parametrize("track_name,n_karts,bucket,teams", [
    ("Oval", 2, 2, [0, 1]), ("Complex", 2, 2, [0, 1]), ("Complex", 2, 1, [0, 1]), ("Oval", 3, 2, [0, 0, 1]),
    ("Complex", 4, 2, [0, 0, 1, 1]), ("Oval", 1, 2, [0]), ("Complex", 4, 1, [0, 1, 2, 3])])
def test_replay_bit_exact(hk, oracle, track_name, n_karts, bucket, teams):
    """Fixed action sequences (legal moves picked at random, plus a few illegal ones): every state, upNext, isOver flag,
    score list, legal set and policy order equal the oracle's, bit for bit."""
    track = tracks.TRACKS[track_name]
    rng = np.random.default_rng(hash((track_name, n_karts, bucket)) % 2**32)
    G = mcts.Game(track, n_karts, bucket)
    OG = _oracle_game(oracle, track, n_karts, bucket)
    batch, ln = 48, n_karts * 8 + 3
    roots, seqs, expect = [], np.zeros((batch, ln, 3), dtype=np.int32), []
    for b in range(batch):
        st = _random_root(rng, track, n_karts, bucket, teams)
        roots.append(st)
        cur, rows = st, []
        for k in range(ln + 1):
            up = OG.up_next(cur)
            over, scores = OG.is_over(cur) if up >= 0 else (-1, np.zeros(0, np.float32))
            mv, gi, cnt = OG.policy_moves(cur) if up >= 0 else ([], [], -1)
            rows.append((cur.astuple(), up, over, scores.copy(), mv, gi, cnt))
            if k == ln:
                break
            if cnt > 0 and rng.random() < 0.9:
                a = mv[int(rng.integers(0, cnt))]
            else:                                                            # arbitrary (possibly illegal) action
                v = int(rng.choice(list(range(6, 15, bucket))))
                a = (v, min(v + bucket, 15), int(rng.integers(1, 5)))
            seqs[b, k] = a
            if up >= 0:
                cur = OG.make_move(cur, a)
        expect.append(rows)
    out = G.replay(roots, seqs)
    for b in range(batch):
        for k in range(ln + 1):
            st, up, over, scores, mv, gi, cnt = expect[b][k]
            assert _state_tuple_np(out["states"][b, k]) == st, (b, k)
            assert out["upnext"][b, k] == up
            assert out["over"][b, k] == over
            assert out["n_scores"][b, k] == len(scores)
            assert out["scores"][b, k, :len(scores)].tobytes() == scores.tobytes()      # bit-exact incl. NaN payloads
            assert out["n_moves"][b, k] == cnt
            if cnt > 0:
                assert [tuple(int(v) for v in out["moves"][b, k, j]) for j in range(cnt)] == mv
                assert [int(v) for v in out["moves_index"][b, k, :cnt]] == gi


@pytest.mark.parametrize("track_name,n_karts,bucket,teams", [("Complex", 2, 2, [0, 1]), ("Oval", 2, 1, [0, 1]), ("Complex", 4, 2, [0, 0, 1, 1]), ("Oval", 3, 2, [0, 1, 2])])
def test_rollout_trace_replays_in_oracle(hk, oracle, track_name, n_karts, bucket, teams):
    """Each GPU rollout, replayed move by move in the oracle with the same Philox stream, gives the same choices, the
    same actions and the same terminal score list bit-exactly."""
    track = tracks.TRACKS[track_name]
    rng = np.random.default_rng(12)
    G = mcts.Game(track, n_karts, bucket)
    OG = _oracle_game(oracle, track, n_karts, bucket)
    for trial in range(3):
        leaf = _random_root(rng, track, n_karts, bucket, teams)
        seed, off, n = 20260003 + trial, 1000 * trial, 256
        tr = G.rollouts_trace(leaf, n, seed=seed, rollout_offset=off)
        for r in range(n):
            ref = OG.rollout(leaf, mode=0, seed=seed, rollout_id=off + r)
            assert tr["n_plies"][r] == ref["n_plies"]
            npl = ref["n_plies"]
            assert [tuple(int(v) for v in tr["actions"][r, k]) for k in range(npl)] == ref["actions"]
            assert [int(v) for v in tr["choice"][r, :npl]] == ref["choices"]
            assert tr["n_scores"][r] == len(ref["scores"])
            assert tr["scores"][r, :len(ref["scores"])].tobytes() == ref["scores"].tobytes()


def test_rollout_statistics_equal_oracle(hk, oracle):
    """hk_mcts_rollouts == the oracle's reduction over the same Philox counters (exact visits; sums to 1e-12),
    and multi-leaf launches equal single-leaf launches."""
    track = tracks.COMPLEX
    G = mcts.Game(track, 2, 2)
    OG = _oracle_game(oracle, track, 2, 2)
    rng = np.random.default_rng(4)
    leaves = [_random_root(rng, track, 2, 2, [0, 1]) for _ in range(3)]
    n = 20000
    for li, leaf in enumerate(leaves):
        got = G.rollouts(leaf, n, seed=99, rollout_offset=li * n)
        ref = OG.rollouts(leaf, n, mode=0, seed=99, rollout_offset=li * n)
        assert np.array_equal(got["visit"], ref["visit"]) and np.array_equal(got["nan_count"], ref["nan_count"])
        assert got["plies"] == ref["plies"]
        assert np.allclose(got["reward_sum"], ref["reward_sum"], rtol=1e-12, atol=1e-9)
    multi = G.rollouts_multi(leaves, n, seed=99, rollout_offset=0)
    for li, leaf in enumerate(leaves):
        single = G.rollouts(leaf, n, seed=99, rollout_offset=li * n)
        assert np.array_equal(multi["visit"][li], single["visit"])
        assert np.allclose(multi["reward_sum"][li], single["reward_sum"], rtol=1e-12, atol=1e-9)


def test_rollout_distribution_matches_reference_procedure(hk, oracle):
    """Decision statistics vs the reference's own sampler (polar Box-Muller N(0,1), float NextGaussian with <= 10
    redraws, RoundToInt(|x|)): chi-square on first-action visit frequencies and z-test on mean terminal score."""
    track = tracks.COMPLEX
    G = mcts.Game(track, 2, 2)
    OG = _oracle_game(oracle, track, 2, 2)
    leaf = tracks.root_state(track, 12, [2, 3], teams=[0, 1], tire_age=2500, times=[0, 37])
    n = 200000
    got = G.rollouts(leaf, n, seed=20260003)
    ref = OG.rollouts(leaf, n, mode=1, seed=5)
    used = (got["visit"] + ref["visit"]) > 0
    a, b = got["visit"][used].astype(float), ref["visit"][used].astype(float)
    big = (a + b) >= 20
    chi2 = float(np.sum((a[big] - b[big]) ** 2 / (a[big] + b[big])))
    dof = int(big.sum()) - 1
    assert chi2 < dof + 6 * np.sqrt(2 * dof) + 10, (chi2, dof)
    for k in range(2):
        ma, mb = got["reward_sum"][:, k].sum() / n, ref["reward_sum"][:, k].sum() / n
        assert abs(ma - mb) < 6 * np.sqrt(0.25 / n * 2) + 1e-9


def test_policy_cdf_and_full_size_rollouts(hk, oracle):
    """CDF tables exported by the library equal the oracle's; 10^6 rollouts (BASELINE config 4) conserve counts and are
    reproducible; disjoint shards (rollout_offset) sum to the whole."""
    for cnt in range(1, abi.HK_MAX_ACTIONS + 1):
        assert np.array_equal(mcts.policy_cdf(cnt), oracle.policy_cdf(cnt))
    track = tracks.COMPLEX
    G = mcts.Game(track, 2, 2)
    leaf = tracks.root_state(track, 3, [2, 3], teams=[0, 1], tire_age=2500, times=[0, 80])
    n = 1_000_000
    whole = G.rollouts(leaf, n, seed=20260003)
    again = G.rollouts(leaf, n, seed=20260003)
    assert whole["visit"].sum() == n and np.array_equal(whole["visit"], again["visit"])
    assert np.allclose(whole["reward_sum"], again["reward_sum"], rtol=1e-12)
    parts = [G.rollouts(leaf, n // 4, seed=20260003, rollout_offset=k * (n // 4)) for k in range(4)]
    assert np.array_equal(sum(p["visit"] for p in parts), whole["visit"])
    assert np.allclose(sum(p["reward_sum"] for p in parts), whole["reward_sum"], rtol=1e-12)
    assert sum(p["plies"] for p in parts) == whole["plies"]
    # 2-kart terminal reward is the constant vector (1, 0) in kart order unless t0 >= 3 t1 (quirk B.6-5)
    tot = whole["reward_sum"].sum(axis=0)
    assert tot[0] == pytest.approx(n - whole["nan_count"].sum(), rel=1e-9) or tot[0] <= n


def test_tree_search_api(hk, oracle):
    """KartMCTS.constructSearchTree / getBestStatesSequence drop-in, both modes of the reference's `parallel` argument.
    parallel=False (what HierarchicalKartAgent calls): the device-resident sequential search handed back as a KartMCTSNode graph —
    node for node the oracle's tree, states replayed lazily, continued by constructSearchTree(root) — and a best-states walk that
    reaches the terminal depth.  parallel=True: the leaf-parallel host tree."""
    track = tracks.COMPLEX
    G = mcts.Game(track, 2, 2)
    st = tracks.root_state(track, 0, [2, 3], teams=[0, 1], tire_age=2500)
    root = mcts.DiscreteGameState(G, st)
    mcts.KartMCTS.random.seed(7)
    node = mcts.KartMCTS.constructSearchTree(root, T=30.0, seed=1, max_iterations=60, parallel=True)
    assert node.numEpisodes > 0 and len(node.children) == len(root.nextMoves())
    best = mcts.KartMCTS.getBestStatesSequence(node)
    assert len(best) >= 1
    assert all(all(k.section == s.lastCompletedSection for k in s.kartStates) for s in best)
    # sequential mode
    OG = _oracle_game(oracle, track, 2, 2)
    node = mcts.KartMCTS.constructSearchTree(root, seed=4242, max_iterations=200)          # sized for the continued call below
    ot = oracle.Tree(OG, st, key=4242)
    assert ot.search(200) == 0
    d = ot.dump(states=True)
    flat = []
    stack = [node]
    while stack:
        n = stack.pop()
        flat.append(n)
        stack.extend(reversed(list(n.children.values())))
    assert len(flat) == ot.size and node.numEpisodes == 200 and node.childrenAsRoot == ot.size - 1
    by_path = {}
    for i in range(ot.size):                                                                  # oracle nodes keyed by their action path
        path, j = [], i
        while d["parent"][j] >= 0:
            path.append(int(d["gen"][j])); j = int(d["parent"][j])
        by_path[tuple(reversed(path))] = i
    for n in flat[::37] + flat[-3:]:
        path, m = [], n
        while m.parent is not None:
            path.append(G.gen_index(next(k for k, v in m.parent.children.items() if v is m))); m = m.parent
        i = by_path[tuple(reversed(path))]
        assert n.numEpisodes == int(d["numEpisodes"][i]) and np.float32(n.totalValue) == d["totalValue"][i]
        assert bytes(n.state.state) == d["states"][i].tobytes()                               # lazily replayed state
    best = mcts.KartMCTS.getBestStatesSequence(node)
    assert len(best) == 8 and len(node._device_best) == 8                                     # the chain simulate() grew, to the terminal depth
    assert all(all(k.section == s.lastCompletedSection for k in s.kartStates) for s in best)
    with pytest.raises(ValueError):
        mcts.KartMCTS.constructSearchTree(node, max_iterations=50)                            # node budget of the first call exhausted
    node = mcts.KartMCTS.constructSearchTree(root, seed=4242, max_iterations=120, reserve_iterations=200)
    node2 = mcts.KartMCTS.constructSearchTree(node, max_iterations=80)                        # constructSearchTree(root): continues on the device
    assert node2.numEpisodes == 200 and node2.childrenAsRoot >= node.childrenAsRoot


@pytest.mark.parametrize("track_name,n_karts,bucket,teams", [("Complex", 2, 2, [0, 1]), ("Oval", 2, 1, [0, 1]), ("Complex", 4, 2, [0, 0, 1, 1]),
                                                             ("Oval", 1, 2, [0]), ("Complex", 3, 2, [0, 1, 2])])
def test_device_tree_search_equals_host_mirror(hk, track_name, n_karts, bucket, teams):
    """hk_mcts_search_batch (one thread block per tree) builds the tree the host mirror of KartMCTS.constructSearchTree builds when
    it is driven by the same Philox streams: same children statistics at the root, same tree size, same getBestStatesSequence."""
    import ctypes as C
    from hierarchicalkarting_b200 import abi, mcts as M, tracks
    track = tracks.COMPLEX if track_name == "Complex" else tracks.OVAL
    G = M.Game(track, n_karts, bucket)
    lanes = [2, 3, 1, 4][:n_karts]
    roots = [tracks.root_state(track, s0, lanes, teams=teams, tire_age=2500, times=[0, 40, 20, 60][:n_karts]) for s0 in (0, 3, 7, 11, 17)]
    K, R, seed = 6, 64, 20260007
    dev = G.search_batch(roots, K, R, seed)
    old_random, old_R = M.KartMCTS.random, M.KartMCTS.rollouts_per_leaf
    try:
        M.KartMCTS.rollouts_per_leaf = R
        for r, root in enumerate(roots):
            M.KartMCTS.random = M.PhiloxPicks(seed + r)
            tree = M.KartMCTS.constructSearchTree(M.DiscreteGameState(G, root), T=1e9, seed=seed + r, max_iterations=K, parallel=True)
            seq = M.KartMCTS.getBestStatesSequence(tree)
            kids = [tree.children[mv] for mv in tree.state.nextMoves()]
            assert int(dev["n_nodes"][r]) == 1 + tree.childrenAsRoot
            assert [int(x) for x in dev["root_episodes"][r][:len(kids)]] == [k.numEpisodes for k in kids]
            assert np.allclose(dev["root_values"][r][:len(kids)], [k.totalValue for k in kids], rtol=1e-9, atol=1e-9)
            assert int(dev["n_best"][r]) == len(seq)
            for a, b in zip(dev["best"][r], seq):
                assert bytes(a) == bytes(b.state)
    finally:
        M.KartMCTS.random, M.KartMCTS.rollouts_per_leaf = old_random, old_R


def test_device_tree_search_chunks(hk):
    """hk_mcts_search_batch works through large batches in chunks of roots (bounded tree slabs); root r keeps Philox key seed + r, so
    the tail of a two-chunk batch equals the same roots searched on their own with the seed shifted."""
    from hierarchicalkarting_b200 import mcts as M, tracks
    track = tracks.OVAL
    G = M.Game(track, 2, 2)
    base = [tracks.root_state(track, s0 % 24, [2, 3], teams=[0, 1], tire_age=2500, times=[0, 10 * (s0 % 7)]) for s0 in range(40)]
    n = 3000
    roots = np.zeros(n, dtype=abi.GAME_STATE_DTYPE)
    for r in range(n):
        roots[r] = np.frombuffer(bytes(base[r % 40]), dtype=abi.GAME_STATE_DTYPE)[0]
    K, R, seed = 80, 4, 99                                  # 2,881 nodes per tree: 2,992 roots per chunk
    big = G.search_batch_array(roots, K, R, seed)
    tail = G.search_batch_array(roots[2990:], K, R, seed + 2990)
    for k in ("n_best", "root_episodes", "n_nodes"):
        assert np.array_equal(big[k][2990:], tail[k]), k
    assert np.allclose(big["root_values"][2990:], tail["root_values"], rtol=1e-12, atol=1e-12)   # double sums through shared-memory atomics
    assert big["best"][2990:].tobytes() == tail["best"].tobytes()


def test_device_tree_search_full_size_properties(hk):
    """BASELINE config 5's planning event at full size (32,768 agents' trees, 24 iterations x 16 rollouts per leaf) through
    size-independent properties: tree sizes and episode counts within their bounds, every returned state has all its karts at
    lastCompletedSection and lies ahead of its root, and the search is reproducible (same seed -> same plans)."""
    from hierarchicalkarting_b200 import mcts as M, race as R, scenarios as S
    track = S.OVAL
    prm = R.race_params(track, high_mode_mcts=True)
    G = R.Races(track, prm)
    game = M.Game(track, 2, prm.velocityBucketSize)
    karts, plans = R.start_grid(track, 16384, seed=20260004)
    G.run(karts, plans, 0, 100)
    roots, _ = R.mcts_root_states_batch(track, prm, karts, plans)
    flat = np.ascontiguousarray(roots.reshape(-1))
    K, RPL = 24, 16
    a = game.search_batch_array(flat, K, RPL, 5)
    b = game.search_batch_array(flat, K, RPL, 5)
    n = flat.shape[0]
    assert n == 32768
    assert a["n_nodes"].min() > 1 and a["n_nodes"].max() <= 1 + K * abi.HK_MAX_ACTIONS
    eps = a["root_episodes"].sum(axis=1)
    assert eps.min() > 0 and eps.max() <= K * abi.HK_MAX_ACTIONS * RPL
    assert a["n_best"].max() <= abi.HK_MCTS_MAX_SEQ and (a["n_best"] >= 1).mean() > 0.9
    for k in range(int(a["n_best"].max())):
        live = a["n_best"] > k
        st = a["best"][live, k]
        for i in range(2):
            on = st["n_karts"] > i
            assert np.all(st["karts"][on, i]["section"] == st["lastCompletedSection"][on])
        assert np.all(st["lastCompletedSection"] > flat["lastCompletedSection"][live])
        assert np.all(st["finalSection"] == flat["finalSection"][live])
    for key in ("n_best", "root_episodes", "n_nodes"):
        assert np.array_equal(a[key], b[key]), key
    assert a["best"].tobytes() == b["best"].tobytes()


@pytest.mark.parametrize("track_name,n_karts,bucket,teams", [("Complex", 2, 2, [0, 1]), ("Oval", 2, 1, [0, 1]), ("Complex", 4, 2, [0, 0, 1, 1])])
def test_device_tree_search_equals_cpu_oracle_tree_search(hk, oracle, track_name, n_karts, bucket, teams):
    """hk_mcts_search_batch against a tree search in which NOTHING comes from the CUDA library: oracle/np_mcts.py restates the tree policy
    of KartMCTS.cs over the C oracle's game primitives and playouts, driven by the Philox streams the ABI documents."""
    from oracle import np_mcts
    track = tracks.COMPLEX if track_name == "Complex" else tracks.OVAL
    G = mcts.Game(track, n_karts, bucket)
    OG = oracle.Game(track.sections_array(), track.n_sections, tracks.kart_array(n_karts), n_karts, tracks.game_params(track, bucket=bucket))
    lanes = [2, 3, 1, 4][:n_karts]
    roots = [tracks.root_state(track, s0, lanes, teams=teams, tire_age=2500, times=[0, 40, 20, 60][:n_karts]) for s0 in (1, 6, 13)]
    K, R, seed = 5, 32, 20260008
    dev = G.search_batch(roots, K, R, seed)
    for r, root in enumerate(roots):
        tree, best, n_nodes = np_mcts.TreeSearch(OG, seed + r).search(root, K, R)
        assert int(dev["n_nodes"][r]) == n_nodes
        assert [int(x) for x in dev["root_episodes"][r][:len(tree.children)]] == [c.numEpisodes for c in tree.children]
        assert np.allclose(dev["root_values"][r][:len(tree.children)], [c.totalValue for c in tree.children], rtol=1e-9, atol=1e-9)
        assert int(dev["n_best"][r]) == len(best)
        for a, b in zip(dev["best"][r], best):
            assert bytes(a) == bytes(b)
