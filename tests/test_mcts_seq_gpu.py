"""GPU parity of the FAITHFUL tree search — KartMCTS.constructSearchTree as the reference's callers run it (parallel == false:
one playout per iteration, every playout state a tree node, KartMCTS.cs:50-106, 194-201, 238-289) — against the CPU oracle
(oracle/hk_oracle_mcts.c), through the C-ABI.  Trees must be BIT-EQUAL node for node (links, numEpisodes, float32 totalValue) under
the same Philox streams; decision statistics of the two GPU search modes are compared with the oracle run with the reference's own
random procedures (System.Random-style uniform picks + the truncated Gaussian of NextGaussian) by chi-square."""
import zlib

import numpy as np
import pytest

from hierarchicalkarting_b200 import abi, mcts, tracks

pytestmark = pytest.mark.gpu


def _oracle_game(oracle, track, n_karts, bucket):
    return oracle.Game(track.sections_array(), track.n_sections, tracks.kart_array(n_karts), n_karts, tracks.game_params(track, bucket))


def _roots(rng, track, n_karts, bucket, teams, n, action_buckets=False):
    out = []
    for _ in range(n):
        sec = int(rng.integers(0, 3 * track.n_sections))
        lanes = [int(x) for x in rng.integers(1, 5, n_karts)]
        buckets = None
        if action_buckets:
            buckets = []
            for _k in range(n_karts):
                v = int(rng.choice(list(range(6, 15, bucket))))
                buckets.append((v, min(v + bucket, 15)))
        times = [0] + [int(x) for x in rng.integers(0, 151, n_karts - 1)]
        st = tracks.root_state(track, sec, lanes, teams=teams, buckets=buckets, tire_age=int(rng.choice([0, 2500, 6000, 13500])),
                               lane_changes=int(rng.integers(0, 3)), times=times)
        if buckets is None:
            for i in range(n_karts):
                st.karts[i].max_velocity = bucket                  # reference root quirk B.6-1: (0, bucket)
        out.append(st)
    return out


def _same_bits(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return np.array_equal(a.view(np.uint32) | np.where(np.isnan(a), 0xFFFFFFFF, 0).astype(np.uint32),
                          b.view(np.uint32) | np.where(np.isnan(b), 0xFFFFFFFF, 0).astype(np.uint32))


def _compare_tree(dev_nodes, otree):
    d = otree.dump()
    assert len(dev_nodes) == otree.size
    assert np.array_equal(dev_nodes["numEpisodes"], d["numEpisodes"])
    assert _same_bits(dev_nodes["totalValue"], d["totalValue"])
    assert np.array_equal(dev_nodes["first_child"], d["first_child"])
    assert np.array_equal(dev_nodes["next_sibling"], d["next_sibling"])
    gen = dev_nodes["gen"].astype(np.int32)
    gen[0] = -1
    assert np.array_equal(gen, d["gen"])
    nch = np.array([bin(int(m)).count("1") for m in dev_nodes["child_mask"]], np.int32)
    assert np.array_equal(nch, d["n_children"])


@pytest.mark.parametrize("track_name,n_karts,bucket,teams,action_buckets", [
    ("Complex", 2, 2, [0, 1], False), ("Oval", 2, 2, [0, 1], False), ("Oval", 2, 1, [0, 1], False), ("Complex", 2, 2, [0, 1], True),
    ("Oval", 3, 2, [0, 0, 1], False), ("Complex", 4, 2, [0, 0, 1, 1], False), ("Complex", 4, 1, [0, 1, 2, 3], True), ("Oval", 1, 2, [0], False)])
def test_faithful_search_trees_bit_equal_oracle(hk, oracle, track_name, n_karts, bucket, teams, action_buckets):
    """Every node of every tree — child lists in insertion order, numEpisodes, totalValue bits — and getBestStatesSequence equal the
    oracle's, fresh and after a continued call on the same trees (constructSearchTree(root), root reuse HierarchicalKartAgent.cs:265-283)."""
    track = tracks.TRACKS[track_name]
    rng = np.random.default_rng(zlib.crc32(repr((track_name, n_karts, bucket, action_buckets)).encode()))
    G = mcts.Game(track, n_karts, bucket)
    OG = _oracle_game(oracle, track, n_karts, bucket)
    n, it1, it2, seed = 24, 96, 40, 20260101
    roots = _roots(rng, track, n_karts, bucket, teams, n, action_buckets)
    F = mcts.Forest(G, n, 1 + (it1 + it2) * n_karts * 8)
    a = F.search(roots, it1, seed)
    assert np.all(a["status"] == 0)
    otrees = []
    for r, root in enumerate(roots):
        ot = oracle.Tree(OG, root, key=seed + r)
        assert ot.search(it1) == 0
        _compare_tree(F.nodes(r), ot)
        best = ot.best_states()
        assert int(a["n_best"][r]) == len(best) and int(a["n_nodes"][r]) == ot.size
        for k, b in enumerate(best):
            assert a["best"][r, k].tobytes() == bytes(b)
        otrees.append(ot)
    # continue half of the trees, restart the others from new roots
    fresh = np.array([r % 2 for r in range(n)], np.int32)
    roots2 = _roots(rng, track, n_karts, bucket, teams, n, action_buckets)
    b = F.search(roots2, it2, seed + 1000, fresh=fresh)
    assert np.all(b["status"] == 0)
    for r in range(n):
        if fresh[r]:
            ot = oracle.Tree(OG, roots2[r], key=seed + 1000 + r)
        else:
            ot = otrees[r]
        assert ot.search(it2) == 0
        _compare_tree(F.nodes(r), ot)
        best = ot.best_states()
        assert int(b["n_best"][r]) == len(best)
        for k, s in enumerate(best):
            assert b["best"][r, k].tobytes() == bytes(s)


def test_fast_path_equals_general_kernel_and_hands_over_when_the_root_fills_up(hk, oracle):
    """The fast path (playouts of a chunk of iterations in parallel, then inserted in order) against the general sequential kernel
    (HK_SEQ_FAST=0) and the oracle, on roots chosen so that BOTH regimes occur: ordinary roots (20 legal moves: never fully expanded) and
    roots with the lane-change budget spent on a straight (3 legal moves: fully expanded after a few iterations, after which findLeaf
    descends by upperConfidenceStrategy and the general kernel must take over mid-call)."""
    import os
    track = tracks.OVAL
    G = mcts.Game(track, 2, 2)
    OG = _oracle_game(oracle, track, 2, 2)
    rng = np.random.default_rng(77)
    roots = _roots(rng, track, 2, 2, [0, 1], 40)
    for r in range(0, 40, 2):                                          # every other root: on a straight with laneChanges == MaxLaneChanges
        roots[r] = tracks.root_state(track, int(rng.choice([0, 1, 10, 11, 12, 22])), [int(x) for x in rng.integers(1, 5, 2)], teams=[0, 1],
                                     tire_age=2500, lane_changes=3, times=[0, int(rng.integers(0, 100))])
        for i in range(2):
            roots[r].karts[i].max_velocity = 2
    its, seed = 150, 4040
    out = {}
    for mode in ("1", "0"):
        os.environ["HK_SEQ_FAST"] = mode
        try:
            F = mcts.Forest(G, 40, 1 + 2 * its * 16)
            a = F.search(roots, its, seed)
            b = F.search(None, its // 2, 0, fresh=np.zeros(40, np.int32))        # continued: the fast path starts from a grown tree
            out[mode] = (a, b, [F.nodes(r).tobytes() for r in range(40)])
        finally:
            os.environ.pop("HK_SEQ_FAST", None)
    for k in ("best", "n_best", "n_nodes", "status"):
        assert out["1"][0][k].tobytes() == out["0"][0][k].tobytes() and out["1"][1][k].tobytes() == out["0"][1][k].tobytes(), k
    assert out["1"][2] == out["0"][2]
    full = 0
    for r in range(40):
        ot = oracle.Tree(OG, roots[r], key=seed + r)
        assert ot.search(its) == 0
        best1 = ot.best_states()
        assert int(out["1"][0]["n_best"][r]) == len(best1)
        assert ot.search(its // 2) == 0
        nodes = np.frombuffer(out["1"][2][r], dtype=abi.MCTS_NODE_DTYPE)
        _compare_tree(nodes, ot)
        full += int(bin(int(nodes["child_mask"][0])).count("1") == int(nodes["n_legal"][0]))
    assert 5 <= full < 40                                              # both regimes were exercised


def test_continued_search_equals_one_long_search(hk):
    """Streams are counted over the life of a tree, so constructSearchTree(root) for k more iterations leaves the tree that one call
    with the total would have built (without the best-states walk in between consuming picks: compared through a fresh forest that
    also walks once in between)."""
    track = tracks.COMPLEX
    G = mcts.Game(track, 2, 2)
    rng = np.random.default_rng(5)
    roots = _roots(rng, track, 2, 2, [0, 1], 64)
    F1, F2 = mcts.Forest(G, 64, 1 + 200 * 16), mcts.Forest(G, 64, 1 + 200 * 16)
    F1.search(roots, 120, 77)
    F1.search(None, 80, 0, fresh=np.zeros(64, np.int32))
    F2.search(roots, 120, 77)
    F2.search(None, 80, 12345, fresh=np.zeros(64, np.int32))           # the seed of a continuing call is ignored
    for r in range(64):
        assert F1.nodes(r).tobytes() == F2.nodes(r).tobytes()
    n = F1.nodes(0)
    assert int(n["numEpisodes"][0]) == 200


def test_one_shot_entry_and_invariants_at_scale(hk):
    """hk_mcts_search_seq_batch over 8,192 roots: root numEpisodes == iterations, children sum to it, the best-states walk reaches the
    terminal depth (8 states for treeSearchDepth 8 — the reference's walk follows the chain its simulate grew), deterministic."""
    track = tracks.COMPLEX
    G = mcts.Game(track, 2, 2)
    rng = np.random.default_rng(11)
    roots = mcts._states_array(_roots(rng, track, 2, 2, [0, 1], 512))
    roots = np.tile(roots, 16)
    it = 128
    a = G.search_seq_batch(roots, it, 99)
    b = G.search_seq_batch(roots, it, 99)
    for k in ("n_best", "root_gen", "root_episodes", "n_nodes"):
        assert np.array_equal(a[k], b[k]), k
    assert a["best"].tobytes() == b["best"].tobytes()
    assert np.all(a["root_episodes"].sum(axis=1) == it)
    assert a["n_nodes"].max() <= 1 + it * 16 and a["n_nodes"].min() > it // 2
    assert (a["n_best"] == 8).mean() > 0.95
    live = a["n_best"] > 0
    st = a["best"][live, 0]
    assert np.all(st["lastCompletedSection"] == roots["lastCompletedSection"][live] + 1)


def _decision_stats(best, n_best, bucket):
    """(first waypoint lane, first waypoint velocity level) of kart 0 and len(bestStates) per tree."""
    k0 = best[:, 0]["karts"][:, 0]
    lane = np.where(n_best > 0, k0["lane"], 0)
    lvl = np.where(n_best > 0, (k0["min_velocity"] - 6) // bucket, -1)
    return lane, lvl, n_best


def _chi2_two_sample(x, y, n_bins):
    """Pearson chi-square statistic and dof of two samples of small non-negative ints (bins with < 10 expected pooled away)."""
    a, b = np.bincount(x, minlength=n_bins).astype(float), np.bincount(y, minlength=n_bins).astype(float)
    keep = (a + b) >= 20
    a2, b2 = np.append(a[keep], a[~keep].sum()), np.append(b[keep], b[~keep].sum())
    ok = (a2 + b2) > 0
    a2, b2 = a2[ok], b2[ok]
    k1, k2 = np.sqrt(b2.sum() / a2.sum()), np.sqrt(a2.sum() / b2.sum())
    stat = float((((k1 * a2 - k2 * b2) ** 2) / (a2 + b2)).sum())
    return stat, max(len(a2) - 1, 1)


@pytest.mark.parametrize("root_section", [3, 11, 30])
def test_decision_statistics_against_reference_procedures(hk, oracle, root_section):
    """BASELINE north_star: 'MCTS decision statistics must match distributionally'.  2,048 seeds per root: the faithful GPU search
    (Philox + closed-form index distribution) against the oracle run with the reference's OWN random procedures (mode 1); compared:
    the first waypoint's lane and velocity level, len(bestStates), and the root's most-visited first action.  chi-square at
    p > 1e-4 per statistic (dof <= 20 -> the bound below).  The leaf-parallel mode is measured with the same statistics and is
    EXPECTED to differ (it expands every child and returns 1-2 states); the test records that it does, so that a change of either mode
    is noticed."""
    from scipy.stats import chi2
    track = tracks.COMPLEX
    bucket, iters, n = 2, 160, 2048
    G = mcts.Game(track, 2, bucket)
    OG = _oracle_game(oracle, track, 2, bucket)
    root = tracks.root_state(track, root_section, [2, 3], teams=[0, 1], tire_age=2500, times=[0, 35])
    for i in range(2):
        root.karts[i].max_velocity = bucket
    roots = mcts._states_array([root] * n)
    dev = G.search_seq_batch(roots, iters, 31337 + root_section)
    d_lane, d_lvl, d_len = _decision_stats(dev["best"], dev["n_best"], bucket)
    d_top = np.array([dev["root_gen"][r][np.argmax(dev["root_episodes"][r])] for r in range(n)])
    rng_states = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) | np.uint64(1)
    ora = oracle.tree_search_batch(OG, roots, iters, seed=0, mode=1, rng_states=rng_states)
    o_lane, o_lvl, o_len = _decision_stats(ora["best"], ora["n_best"], bucket)
    o_top = np.array([ora["root_gen"][r][np.argmax(ora["root_episodes"][r])] for r in range(n)])
    for name, x, y, bins in (("lane", d_lane, o_lane, 5), ("level", d_lvl + 1, o_lvl + 1, 7), ("len", d_len, o_len, 17), ("top", d_top, o_top, 36)):
        stat, dof = _chi2_two_sample(np.asarray(x, np.int64), np.asarray(y, np.int64), bins)
        assert stat < chi2.isf(1e-4, dof), (name, stat, dof, np.bincount(x, minlength=bins), np.bincount(y, minlength=bins))
    # the leaf-parallel mode: same statistics, recorded divergence (INTEGRATION.md quotes these)
    lp = G.search_batch_array(roots[:512], 24, 16, 31337 + root_section)
    l_lane, l_lvl, l_len = _decision_stats(lp["best"], lp["n_best"], bucket)
    stat_len, dof_len = _chi2_two_sample(np.asarray(l_len, np.int64), np.asarray(o_len[:512], np.int64), 17)
    assert stat_len > chi2.isf(1e-4, dof_len)                          # 1-2 states against the reference's chain to the terminal depth
    assert l_len.max() <= 2 and o_len.min() >= 7


def test_game_golden_fixture_on_device(hk):
    """The committed fixture tests/golden/game_golden.npz — kart transitions and sequential tree searches frozen from the CPU oracle after
    two independent restatements agreed with it — reproduced by the CUDA library through the C-ABI, bit for bit."""
    G = np.load("tests/golden/game_golden.npz")
    for name, track, bucket in (("oval2", tracks.OVAL, 2), ("complex1", tracks.COMPLEX, 1)):
        game = mcts.Game(track, 2, bucket)
        ks, acts, new = G[f"{name}_kart_states"], G[f"{name}_actions"], G[f"{name}_new_states"]
        n = len(ks)
        roots = np.zeros(n, dtype=abi.GAME_STATE_DTYPE)
        roots["n_karts"] = 1
        roots["initialSection"] = roots["lastCompletedSection"] = ks["section"]
        roots["finalSection"] = ks["section"] + 8
        for f in abi.KART_STATE_FIELDS:
            roots["karts"][:, 0][f] = ks[f]
        out = game.replay(roots, acts.reshape(n, 1, 3))
        got = out["states"][:, 1]["karts"][:, 0]
        for f in abi.KART_STATE_FIELDS:
            assert np.array_equal(got[f], new[f]), f
        res = game.search_seq_batch(G[f"{name}_roots"], 64, 20260401)
        for k in ("n_best", "root_gen", "root_episodes", "n_nodes"):
            assert np.array_equal(res[k], G[f"{name}_search_{k}"]), k
        assert res["best"].tobytes() == G[f"{name}_search_best"].tobytes()
        assert np.array_equal(res["root_values"].view(np.uint32), G[f"{name}_search_root_values"].view(np.uint32))


def test_edge_cases(hk, oracle):
    """The corners the reference's callers can reach (and the ABI's argument checks): zero iterations (best states of a bare root), roots one
    section before the end of the search window (one-ply games for every kart), a root that is already terminal (a tree of one node that
    collects every episode), a node slab that is too small (status 3, the tree
    stays consistent and no neighbour is touched), invalid arguments."""
    track = tracks.OVAL
    G = mcts.Game(track, 2, 2)
    OG = _oracle_game(oracle, track, 2, 2)
    rng = np.random.default_rng(2)
    roots = _roots(rng, track, 2, 2, [0, 1], 8)
    # zero iterations
    F = mcts.Forest(G, 8, 64)
    out = F.search(roots, 0, 5)
    for r in range(8):
        ot = oracle.Tree(OG, roots[r], key=5 + r)
        assert ot.search(0) == 0
        ob = ot.best_states()
        assert int(out["n_best"][r]) == len(ob) and int(out["n_nodes"][r]) == ot.size == 1 and int(out["status"][r]) == 0
    # one section left for every kart: playouts of n_karts plies
    near = []
    for st in roots:
        s2 = abi.hk_game_state.from_buffer_copy(bytes(st))
        s2.finalSection = s2.lastCompletedSection + 1
        near.append(s2)
    out = F.search(near, 40, 6)
    for r in range(8):
        ot = oracle.Tree(OG, near[r], key=6 + r)
        assert ot.search(40) == 0
        _compare_tree(F.nodes(r), ot)
        ob = ot.best_states()
        assert int(out["n_best"][r]) == len(ob)
        for j, s_ in enumerate(ob):
            assert bytes(s_) == out["best"][r, j].tobytes()
    # a terminal root
    done = []
    for st in roots:
        s2 = abi.hk_game_state.from_buffer_copy(bytes(st))
        s2.finalSection = s2.lastCompletedSection
        for k in range(2):
            s2.karts[k].section = s2.lastCompletedSection
        done.append(s2)
    out = F.search(done, 3, 1)                                       # isOver() at once: simulate returns the root, which collects the episodes
    for r in range(8):
        ot = oracle.Tree(OG, done[r], key=1 + r)
        assert ot.search(3) == 0 and ot.size == 1
        _compare_tree(F.nodes(r), ot)
        assert int(out["n_best"][r]) == len(ot.best_states()) == 0 and int(out["status"][r]) == 0
    # slab too small: 8 trees x 20 nodes, 50 iterations
    F2 = mcts.Forest(G, 8, 20)
    fresh = np.ones(8, np.int32)
    out = F2.search(roots, 50, 7, fresh=fresh)
    assert np.all(out["status"] == 3) and np.all(out["n_nodes"] <= 20)
    for r in range(8):
        nd = F2.nodes(r)
        ot = oracle.Tree(OG, roots[r], key=7 + r)
        done_it = int(nd["numEpisodes"][0])                          # the iterations that fitted are exactly the oracle's first ones;
        assert 0 < done_it < 50 and ot.search(done_it) == 0          # the one that did not left its first nodes behind, without episodes
        d = ot.dump()
        assert ot.size <= len(nd) <= 20 and np.array_equal(nd["numEpisodes"][:ot.size], d["numEpisodes"]) and not nd["numEpisodes"][ot.size:].any()
        assert _same_bits(nd["totalValue"][:ot.size], d["totalValue"])
    # arguments
    lib = abi.load_library()
    assert lib.hk_mcts_forest_search(None, None, None, 1, 0, None, None, None, None) == abi.HK_ERR_INVALID_ARGUMENT
    with pytest.raises(abi.HKError):
        mcts.Forest(G, 0, 10)
    with pytest.raises(abi.HKError):
        F.search(roots, -1, 0)
    bad = abi.hk_game_state.from_buffer_copy(bytes(roots[0]))
    bad.n_karts = 3                                                  # a state of another game
    with pytest.raises(abi.HKError):
        F.search([bad] + roots[1:], 1, 0)


@pytest.mark.parametrize("levels", ["0", "1", "2", "3"])
@pytest.mark.parametrize("track_name,bucket", [("Oval", 2), ("Complex", 1)])
def test_prefix_tables_at_every_depth_leave_the_same_trees(hk, oracle, levels, track_name, bucket):
    """The insertion kernel's auxiliary prefix tables (first 1, 2 or 3 plies: whatever fits the forest's budget; none with HK_SEQ_AUX=0) are not
    part of the tree: with every depth — and across a continued search, a tree that is kept while its neighbours start afresh, and a slab
    that fills up — the node records equal the oracle's."""
    import os
    track = tracks.OVAL if track_name == "Oval" else tracks.COMPLEX
    G = mcts.Game(track, 2, bucket)
    OG = _oracle_game(oracle, track, 2, bucket)
    rng = np.random.default_rng(5 + bucket)
    n = 24
    roots = _roots(rng, track, 2, bucket, [0, 1], n)
    roots2 = _roots(rng, track, 2, bucket, [0, 1], n)
    os.environ["HK_SEQ_AUX_LEVELS"] = levels
    if levels == "0":
        os.environ["HK_SEQ_AUX"] = "0"
    try:
        F = mcts.Forest(G, n, 3000)
        F.search(roots, 120, 900)
        fresh = np.array([1 if r % 3 == 0 else 0 for r in range(n)], np.int32)            # every third tree starts over from another root
        F.search(roots2, 100, 901, fresh=fresh)
        F.search(None, 60, 0, fresh=np.zeros(n, np.int32))                                # the last 20 iterations no longer fit the slab
        got = [F.nodes(r) for r in range(n)]
    finally:
        os.environ.pop("HK_SEQ_AUX_LEVELS", None)
        os.environ.pop("HK_SEQ_AUX", None)
    full = 0
    for r in range(n):
        if fresh[r]:
            ot = oracle.Tree(OG, roots2[r], key=901 + r)
            ot.search(100 + 60)
        else:
            ot = oracle.Tree(OG, roots[r], key=900 + r)
            ot.search(120 + 100 + 60)
        if ot.size <= 3000:
            _compare_tree(got[r], ot)
        else:                                                          # the device tree stopped when its slab was full: the iterations that
            full += 1                                                  # fitted are the oracle's first ones, the one that did not left nodes without episodes
            done = int(got[r]["numEpisodes"][0])
            assert 0 < done < 280 and len(got[r]) <= 3000
            o2 = oracle.Tree(OG, roots[r], key=900 + r)
            o2.search(done)
            d = o2.dump()
            assert np.array_equal(got[r]["numEpisodes"][:o2.size], d["numEpisodes"]) and not got[r]["numEpisodes"][o2.size:].any()
            assert _same_bits(got[r]["totalValue"][:o2.size], d["totalValue"])
    assert full > 0
