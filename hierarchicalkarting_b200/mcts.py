"""Host-side mirror of the reference's `KartGame.AI.MCTS` API over the C-ABI (Python flavour for tests/bench; the C++
flavour is host/KartMCTS.hpp, the C# shim csharp/KartMCTS.cs).

  DiscreteGameState.{upNext,isOver,nextMoves,makeMove}   Assets/Karting/Scripts/AI/MCTS/KartDiscreteGame.cs:188,251,322,420
  KartMCTSNode, KartMCTS.constructSearchTree / getBestStatesSequence / upperConfidenceStrategy / NextGaussian
                                                         Assets/Karting/Scripts/AI/MCTS/KartMCTS.cs:18-38,50-122,167-236
Every game evaluation (transitions, legal moves, terminal scores, playouts) runs in the CUDA library.  constructSearchTree has the
reference's two modes: `parallel=False` (what HierarchicalKartAgent calls, :250,271) is the sequential search — findLeaf, ONE playout
whose every state becomes a node, backpropagate from its terminal node — run entirely on the device (hk_mcts_forest_search, one GPU
thread per tree) and handed back as a KartMCTSNode graph; `parallel=True` is the reference's leaf-parallel `processLeaf` (:124-159)
with R playouts per child: the tree stays on the host and consumes GPU rollout statistics.
"""
from __future__ import annotations

import ctypes as C
import math
import random as _random
import time

import numpy as np

from . import abi
from .tracks import Track, game_params, kart_array


class Game:
    """Immutable device-resident game description (track + kart constants + parameters): hk_game handle."""

    def __init__(self, track: Track, n_karts: int = 2, bucket: int = 2, time_precision: int = 100, depth: int = 8,
                 kart_consts=None, params: abi.hk_game_params | None = None):
        self.track = track
        self.n_karts = n_karts
        self.params = params if params is not None else game_params(track, bucket, time_precision, depth)
        self._sections = track.sections_array()
        self._karts = kart_array(n_karts) if kart_consts is None else kart_array(n_karts, kart_consts)
        self._h = C.c_void_p()
        lib = abi.load_library()
        abi.check(lib.hk_game_create(self._sections, track.n_sections, self._karts, n_karts, None, 0, C.byref(self.params),
                                     C.byref(self._h)))
        self.n_cand = 4 * len(range(6, 15, self.params.velocityBucketSize))

    def close(self):
        if getattr(self, "_h", None):
            abi.load_library().hk_game_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- exact-parity entry ----------------------------------------------------------------------------------------
    def replay(self, roots, actions):
        """roots: list of hk_game_state (or GAME_STATE_DTYPE array); actions: int array [batch][len][3]."""
        if not isinstance(roots, np.ndarray):
            r = np.zeros(len(roots), dtype=abi.GAME_STATE_DTYPE)
            for i, s in enumerate(roots):
                C.memmove(r[i:i + 1].ctypes.data, C.byref(s), C.sizeof(abi.hk_game_state))
            roots = r
        batch = roots.shape[0]
        acts = np.ascontiguousarray(actions, dtype=np.int32).reshape(batch, -1, 3)
        ln = acts.shape[1]
        np1 = (batch, ln + 1)
        out = dict(states=np.zeros(np1, dtype=abi.GAME_STATE_DTYPE), upnext=np.zeros(np1, dtype=np.int32),
                   over=np.zeros(np1, dtype=np.int32), n_scores=np.zeros(np1, dtype=np.int32),
                   scores=np.zeros(np1 + (2 * abi.HK_MAX_KARTS,), dtype=np.float32), n_moves=np.zeros(np1, dtype=np.int32),
                   moves=np.zeros(np1 + (abi.HK_MAX_ACTIONS, 3), dtype=np.int32),
                   moves_index=np.zeros(np1 + (abi.HK_MAX_ACTIONS,), dtype=np.int32))
        lib = abi.load_library()
        abi.check(lib.hk_game_replay_batch(self._h, batch, ln, abi.vptr(roots), abi.vptr(acts), abi.vptr(out["states"]),
                                           abi.vptr(out["upnext"]), abi.vptr(out["over"]), abi.vptr(out["n_scores"]),
                                           abi.vptr(out["scores"]), abi.vptr(out["n_moves"]), abi.vptr(out["moves"]),
                                           abi.vptr(out["moves_index"])))
        return out

    # ---- leaf-parallel rollouts ---------------------------------------------------------------------------------------
    def rollouts(self, leaf: abi.hk_game_state, n_rollouts: int, seed: int = 0, rollout_offset: int = 0):
        visit = np.zeros(abi.HK_MAX_ACTIONS, dtype=np.int64)
        rsum = np.zeros((abi.HK_MAX_ACTIONS, abi.HK_MAX_KARTS))
        nanc = np.zeros(abi.HK_MAX_ACTIONS, dtype=np.int64)
        plies = C.c_int64(0)
        lib = abi.load_library()
        abi.check(lib.hk_mcts_rollouts(self._h, C.byref(leaf), n_rollouts, seed, rollout_offset,
                                       visit.ctypes.data_as(C.POINTER(C.c_int64)), abi.dptr(rsum),
                                       nanc.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(plies)))
        return dict(visit=visit, reward_sum=rsum, nan_count=nanc, plies=plies.value)

    def rollouts_multi(self, leaves, rollouts_per_leaf: int, seed: int = 0, rollout_offset: int = 0):
        n = len(leaves)
        arr = (abi.hk_game_state * n)(*leaves)
        visit = np.zeros((n, abi.HK_MAX_ACTIONS), dtype=np.int64)
        rsum = np.zeros((n, abi.HK_MAX_ACTIONS, abi.HK_MAX_KARTS))
        nanc = np.zeros((n, abi.HK_MAX_ACTIONS), dtype=np.int64)
        plies = np.zeros(n, dtype=np.int64)
        lib = abi.load_library()
        abi.check(lib.hk_mcts_rollouts_multi(self._h, C.cast(arr, C.c_void_p), n, rollouts_per_leaf, seed, rollout_offset,
                                             abi.vptr(visit), abi.vptr(rsum), abi.vptr(nanc), abi.vptr(plies)))
        return dict(visit=visit, reward_sum=rsum, nan_count=nanc, plies=plies)

    def search_batch(self, roots, iterations: int, rollouts_per_leaf: int, seed: int = 0):
        """hk_mcts_search_batch: constructSearchTree + getBestStatesSequence for every root in one launch (one thread block per
        tree).  Returns best (list of lists of hk_game_state), root_episodes / root_values of the root's children in nextMoves()
        order, n_nodes."""
        n = len(roots)
        arr = (abi.hk_game_state * n)(*roots)
        best = (abi.hk_game_state * (n * abi.HK_MCTS_MAX_SEQ))()
        n_best = np.zeros(n, dtype=np.int32)
        eps = np.zeros((n, abi.HK_MAX_ACTIONS), dtype=np.int32)
        vals = np.zeros((n, abi.HK_MAX_ACTIONS))
        nodes = np.zeros(n, dtype=np.int32)
        lib = abi.load_library()
        abi.check(lib.hk_mcts_search_batch(self._h, C.cast(arr, C.c_void_p), n, iterations, rollouts_per_leaf, seed,
                                           C.cast(best, C.c_void_p), abi.vptr(n_best), abi.vptr(eps), abi.vptr(vals), abi.vptr(nodes)))
        seqs = [[best[r * abi.HK_MCTS_MAX_SEQ + k] for k in range(int(n_best[r]))] for r in range(n)]
        return dict(best=seqs, n_best=n_best, root_episodes=eps, root_values=vals, n_nodes=nodes)

    def search_batch_array(self, roots: np.ndarray, iterations: int, rollouts_per_leaf: int, seed: int = 0):
        """search_batch over a numpy array of abi.GAME_STATE_DTYPE; `best` comes back as an array [n][HK_MCTS_MAX_SEQ] of that dtype."""
        roots = np.ascontiguousarray(roots, dtype=abi.GAME_STATE_DTYPE)
        n = roots.shape[0]
        best = np.zeros((n, abi.HK_MCTS_MAX_SEQ), dtype=abi.GAME_STATE_DTYPE)
        n_best = np.zeros(n, dtype=np.int32)
        eps = np.zeros((n, abi.HK_MAX_ACTIONS), dtype=np.int32)
        vals = np.zeros((n, abi.HK_MAX_ACTIONS))
        nodes = np.zeros(n, dtype=np.int32)
        abi.check(abi.load_library().hk_mcts_search_batch(self._h, abi.vptr(roots), n, iterations, rollouts_per_leaf, seed, abi.vptr(best),
                                                          abi.vptr(n_best), abi.vptr(eps), abi.vptr(vals), abi.vptr(nodes)))
        return dict(best=best, n_best=n_best, root_episodes=eps, root_values=vals, n_nodes=nodes)

    def search_seq_batch(self, roots, iterations: int, seed: int = 0):
        """hk_mcts_search_seq_batch: the reference's sequential search (constructSearchTree with parallel == false: one playout per
        iteration, every playout state a tree node) + getBestStatesSequence for every root, one GPU thread per tree.  `roots`: list
        of hk_game_state or array of abi.GAME_STATE_DTYPE.  Returns best [n][HK_MCTS_MAX_SEQ] (GAME_STATE_DTYPE), n_best, and the
        root's children in insertion order: root_gen (-1 past the end), root_episodes, root_values; n_nodes."""
        roots = _states_array(roots)
        n = roots.shape[0]
        best = np.zeros((n, abi.HK_MCTS_MAX_SEQ), dtype=abi.GAME_STATE_DTYPE)
        n_best = np.zeros(n, dtype=np.int32)
        gen = np.zeros((n, abi.HK_MAX_ACTIONS), dtype=np.int32)
        eps = np.zeros((n, abi.HK_MAX_ACTIONS), dtype=np.int32)
        vals = np.zeros((n, abi.HK_MAX_ACTIONS), dtype=np.float32)
        nodes = np.zeros(n, dtype=np.int32)
        abi.check(abi.load_library().hk_mcts_search_seq_batch(self._h, abi.vptr(roots), n, iterations, seed, abi.vptr(best), abi.vptr(n_best),
                                                              abi.vptr(gen), abi.vptr(eps), abi.vptr(vals), abi.vptr(nodes)))
        return dict(best=best, n_best=n_best, root_gen=gen, root_episodes=eps, root_values=vals, n_nodes=nodes)

    def rollouts_trace(self, leaf: abi.hk_game_state, n_rollouts: int, seed: int = 0, rollout_offset: int = 0):
        n = n_rollouts
        out = dict(n_plies=np.zeros(n, dtype=np.int32), actions=np.zeros((n, abi.HK_MAX_PLIES, 3), dtype=np.int32),
                   choice=np.zeros((n, abi.HK_MAX_PLIES), dtype=np.int32), n_scores=np.zeros(n, dtype=np.int32),
                   scores=np.zeros((n, 2 * abi.HK_MAX_KARTS), dtype=np.float32))
        lib = abi.load_library()
        abi.check(lib.hk_mcts_rollouts_trace(self._h, C.byref(leaf), n, seed, rollout_offset, abi.vptr(out["n_plies"]),
                                             abi.vptr(out["actions"]), abi.vptr(out["choice"]), abi.vptr(out["n_scores"]),
                                             abi.vptr(out["scores"])))
        return out

    def gen_index(self, action) -> int:
        mn, _, lane = action
        return ((mn - 6) // self.params.velocityBucketSize) * 4 + lane - 1

    def action_of(self, gi: int):
        b = self.params.velocityBucketSize
        v = 6 + (gi // 4) * b
        return (v, min(v + b, 15), gi % 4 + 1)


def _states_array(roots) -> np.ndarray:
    if isinstance(roots, np.ndarray):
        return np.ascontiguousarray(roots, dtype=abi.GAME_STATE_DTYPE)
    r = np.zeros(len(roots), dtype=abi.GAME_STATE_DTYPE)
    for i, s in enumerate(roots):
        C.memmove(r[i:i + 1].ctypes.data, C.byref(s), C.sizeof(abi.hk_game_state))
    return r


class Forest:
    """hk_mcts_forest: device-resident trees of the reference's sequential search that survive between calls, like
    HierarchicalKartAgent.currentRoot (HierarchicalKartAgent.cs:265-283)."""

    def __init__(self, game: Game, n_trees: int, max_nodes_per_tree: int):
        self.game, self.n_trees, self.max_nodes = game, n_trees, max_nodes_per_tree
        self._h = C.c_void_p()
        abi.check(abi.load_library().hk_mcts_forest_create(game._h, n_trees, max_nodes_per_tree, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            abi.load_library().hk_mcts_forest_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def search(self, roots, iterations: int, seed: int = 0, fresh=None):
        """constructSearchTree(state) for trees with fresh[r] != 0 (None: all), constructSearchTree(root) for the others."""
        n = self.n_trees
        roots = _states_array(roots) if roots is not None else None
        fr = None if fresh is None else np.ascontiguousarray(fresh, dtype=np.int32)
        best = np.zeros((n, abi.HK_MCTS_MAX_SEQ), dtype=abi.GAME_STATE_DTYPE)
        n_best, nodes, status = (np.zeros(n, dtype=np.int32) for _ in range(3))
        abi.check(abi.load_library().hk_mcts_forest_search(self._h, abi.vptr(roots), abi.vptr(fr), iterations, seed, abi.vptr(best),
                                                           abi.vptr(n_best), abi.vptr(nodes), abi.vptr(status)))
        return dict(best=best, n_best=n_best, n_nodes=nodes, status=status)

    def nodes(self, tree: int) -> np.ndarray:
        """Records of one tree in creation order (abi.MCTS_NODE_DTYPE)."""
        n = C.c_int32(0)
        lib = abi.load_library()
        abi.check(lib.hk_mcts_forest_nodes(self._h, tree, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=abi.MCTS_NODE_DTYPE)
        abi.check(lib.hk_mcts_forest_nodes(self._h, tree, abi.vptr(out), n.value, C.byref(n)))
        return out


def policy_cdf(cnt: int) -> np.ndarray:
    out = np.zeros(cnt, dtype=np.uint32)
    abi.check(abi.load_library().hk_policy_cdf(cnt, out.ctypes.data_as(C.POINTER(C.c_uint32))))
    return out


# ---- reference-shaped object API ------------------------------------------------------------------------------------
class DiscreteGameState:
    """KartDiscreteGame.cs:174-447 — value state + a Game handle; every query is one GPU replay call."""

    def __init__(self, game: Game, state: abi.hk_game_state):
        self.game = game
        self.state = state
        self._info = None

    def _query(self):
        if self._info is None:
            out = self.game.replay([self.state], np.zeros((1, 0, 3), dtype=np.int32))
            nm = int(out["n_moves"][0, 0])
            ns = int(out["n_scores"][0, 0])
            self._info = dict(upnext=int(out["upnext"][0, 0]), over=int(out["over"][0, 0]),
                              scores=[float(x) for x in out["scores"][0, 0, :ns]],
                              moves=[tuple(int(v) for v in out["moves"][0, 0, k]) for k in range(max(nm, 0))],
                              moves_index=[int(v) for v in out["moves_index"][0, 0, :max(nm, 0)]])
        return self._info

    def upNext(self) -> int:
        return self._query()["upnext"]

    def isOver(self):
        i = self._query()
        if i["upnext"] < 0:
            raise IndexError("upNext() == -1 (ArgumentOutOfRangeException at KartDiscreteGame.cs:326)")
        return bool(i["over"]), list(i["scores"])

    def nextMoves(self):
        """Legal moves in GENERATION order (KartDiscreteGame.cs:329-340), as the reference returns them."""
        i = self._query()
        if i["upnext"] < 0:
            raise IndexError("upNext() == -1 (ArgumentOutOfRangeException at KartDiscreteGame.cs:326)")
        return [m for _, m in sorted(zip(i["moves_index"], i["moves"]))]

    def policyMoves(self):
        """Legal moves in the rollout policy's order (KartMCTS.cs:256)."""
        return list(self._query()["moves"])

    def makeMove(self, action) -> "DiscreteGameState":
        out = self.game.replay([self.state], np.array([[list(action)]], dtype=np.int32))
        st = abi.hk_game_state()
        C.memmove(C.byref(st), out["states"][0, 1:2].ctypes.data, C.sizeof(abi.hk_game_state))
        return DiscreteGameState(self.game, st)

    @property
    def kartStates(self):
        return [self.state.karts[i] for i in range(self.state.n_karts)]

    @property
    def lastCompletedSection(self):
        return self.state.lastCompletedSection


def philox_first(key: int, counter: int, ply: int) -> int:
    """First word of Philox4x32-10 with counter (counter lo, counter hi, ply, 0) and key `key` — the generator of the rollouts
    (csrc/hk_game.cu) and of the tie-breaking picks of hk_mcts_search_batch."""
    M = 0xFFFFFFFF
    c0, c1, c2, c3 = counter & M, (counter >> 32) & M, ply & M, 0
    k0, k1 = key & M, (key >> 32) & M
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M, p1 & M, ((p0 >> 32) ^ c3 ^ k1) & M, p0 & M
        k0, k1 = (k0 + 0x9E3779B9) & M, (k1 + 0xBB67AE85) & M
    return c0


class PhiloxPicks:
    """Stand-in for KartMCTS.random that replays the pick stream of hk_mcts_search_batch for root seed `rseed` (= seed + root
    index): upperConfidenceStrategy's random initial child (it only matters for exact ties)."""

    def __init__(self, rseed: int):
        self.key, self.ctr = (rseed ^ 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF, 0

    def randrange(self, n: int) -> int:
        v = philox_first(self.key, self.ctr, 0) % n
        self.ctr += 1
        return v

    def getrandbits(self, k: int) -> int:
        return 0


class KartMCTSNode:                                     # KartMCTS.cs:18-38
    def __init__(self, state: DiscreteGameState, parent=None, createdBy=""):
        self.state = state
        self.parent = parent
        self.children = {}
        self.totalValue = 0.0
        self.numEpisodes = 0
        self.childrenAsRoot = 0
        self.createdBy = createdBy


class _DeviceNode(KartMCTSNode):
    """A KartMCTSNode of a tree that lives on the device (hk_mcts_forest): links and statistics are copied from the node records,
    the state — which the device does not store — is replayed from the root along the node's actions on first access."""

    def __init__(self, game: Game, parent, action, rec):
        self._game, self._action, self._state = game, action, None
        self.parent = parent
        self.children = {}
        self.totalValue = float(rec["totalValue"])
        self.numEpisodes = int(rec["numEpisodes"])
        self.childrenAsRoot = 0
        self.createdBy = ""

    @property
    def state(self) -> DiscreteGameState:
        if self._state is None:
            chain = []
            n = self
            while n._state is None:
                chain.append(n)
                n = n.parent
            chain.reverse()
            acts = np.array([[list(c._action) for c in chain]], dtype=np.int32)
            out = self._game.replay([n._state.state], acts)
            for k, c in enumerate(chain):
                st = abi.hk_game_state()
                C.memmove(C.byref(st), out["states"][0, k + 1:k + 2].ctypes.data, C.sizeof(abi.hk_game_state))
                c._state = DiscreteGameState(self._game, st)
        return self._state

    @state.setter
    def state(self, v):
        self._state = v


class KartMCTS:
    random = _random.Random()
    rollouts_per_leaf = 4096                            # GPU batch per expanded leaf (the reference plays 1 per iteration)

    @staticmethod
    def _uct_weight(node) -> float:                     # KartMCTS.cs:162-165 (integer division, no sqrt term: quirk B.6-7)
        if node.numEpisodes == 0:
            raise ZeroDivisionError
        ratio = node.parent.numEpisodes // node.numEpisodes
        lg = math.log(ratio) if ratio > 0 else float("-inf")
        return np.float32(np.float32(node.totalValue) / np.float32(node.numEpisodes)) + np.float32(lg)

    @staticmethod
    def upperConfidenceStrategy(node):                  # KartMCTS.cs:167-192
        keys = list(node.children.keys())
        best = keys[KartMCTS.random.randrange(len(keys))]
        best_uct = KartMCTS._uct_weight(node.children[best])
        for k, child in node.children.items():
            w = KartMCTS._uct_weight(child)
            if w > best_uct:
                best_uct, best = w, k
        return best

    @staticmethod
    def _find_leaf(root):                               # KartMCTS.cs:194-201
        while len(root.children) > 0 and len(root.children) == len(root.state.nextMoves()):
            root = root.children[KartMCTS.upperConfidenceStrategy(root)]
        return root

    @staticmethod
    def _process_leaf(node, seed, offset):
        """GPU form of processLeaf (:124-159): expand every legal child, R rollouts through each (one launch over the
        leaf; rollouts are attributed to children by their first action), then backpropagate the summed results."""
        over, _ = node.state.isOver()
        if over:
            # terminal leaf: the reference's simulate() returns immediately and backpropagates the terminal scores
            _, scores = node.state.isOver()
            KartMCTS._backpropagate(node, scores, 1)
            return 0
        game = node.state.game
        moves = node.state.nextMoves()
        created = 0
        new = [mv for mv in moves if mv not in node.children]   # initials[j] = new KartMCTSNode(node.state.makeMove(action), node) (:142)
        if new:                                         # one batched replay call makes every missing child
            out = game.replay([node.state.state] * len(new), np.array([[list(mv)] for mv in new], dtype=np.int32))
            for j, mv in enumerate(new):
                st = abi.hk_game_state()
                C.memmove(C.byref(st), out["states"][j, 1:2].ctypes.data, C.sizeof(abi.hk_game_state))
                child = DiscreteGameState(game, st)
                nm, ns = int(out["n_moves"][j, 1]), int(out["n_scores"][j, 1])   # the same call already evaluated the child
                child._info = dict(upnext=int(out["upnext"][j, 1]), over=int(out["over"][j, 1]),
                                   scores=[float(x) for x in out["scores"][j, 1, :ns]],
                                   moves=[tuple(int(v) for v in out["moves"][j, 1, k]) for k in range(max(nm, 0))],
                                   moves_index=[int(v) for v in out["moves_index"][j, 1, :max(nm, 0)]])
                node.children[mv] = KartMCTSNode(child, node)
                created += 1
        kids = [node.children[mv] for mv in moves]
        R = KartMCTS.rollouts_per_leaf
        stats = game.rollouts_multi([k.state.state for k in kids], R, seed=seed, rollout_offset=offset)
        for j, child in enumerate(kids):
            visits = int(stats["visit"][j].sum())
            if visits == 0:                             # child already terminal: simulate() returns its scores at once (:246-249)
                _, scores = child.state.isOver()
                KartMCTS._backpropagate(child, scores, R)
                continue
            cnt = visits - int(stats["nan_count"][j].sum())
            rsum = stats["reward_sum"][j].sum(axis=0)   # per-kart sum of terminal scores over this child's rollouts
            n = child                                   # backpropagate (:280-289): every ancestor adds result[its own upNext()]
            while n is not None:
                up = n.state.upNext()
                n.totalValue += float(rsum[up]) if up >= 0 else 0.0
                n.numEpisodes += cnt
                n = n.parent
        return created

    @staticmethod
    def _backpropagate(node, result, count):            # KartMCTS.cs:280-289
        while node is not None:
            node.totalValue += result[node.state.upNext()] * count
            node.numEpisodes += count
            node = node.parent

    iterations_per_second = 600.0                       # sequential mode: how a wall-clock budget T maps to an iteration count
                                                        # (the reference's loop is budgeted in seconds of ITS C# simulate(), KartMCTS.cs:55;
                                                        # ~1 ms per iteration is what LINQ-heavy simulate() manages — a calibration knob)

    @staticmethod
    def _graph_from_device(game: Game, forest: Forest, root_state: DiscreteGameState):
        rec = forest.nodes(0)
        nodes = [None] * len(rec)
        root = _DeviceNode(game, None, None, rec[0])
        root._state = root_state
        nodes[0] = root
        order = [0]
        for i in order:                                 # parents are created before their children (creation order = index order)
            c = int(rec[i]["first_child"])
            while c >= 0:
                act = game.action_of(int(rec[c]["gen"]))
                nodes[c] = _DeviceNode(game, nodes[i], act, rec[c])
                nodes[i].children[act] = nodes[c]       # insertion order = the Dictionary's enumeration order
                order.append(c)
                c = int(rec[c]["next_sibling"])
        root.childrenAsRoot = len(rec) - 1
        root._forest, root._forest_iterations = forest, int(rec[0]["numEpisodes"])
        return root

    @staticmethod
    def _construct_sequential(state_or_root, T, seed, max_iterations, reserve_iterations=None):
        """KartMCTS.cs:50-78 / :80-106 with parallel == false, on the device; the result is a KartMCTSNode graph."""
        iterations = max_iterations if max_iterations is not None else max(1, int(T * KartMCTS.iterations_per_second))
        if isinstance(state_or_root, KartMCTSNode):
            root = state_or_root
            forest = getattr(root, "_forest", None)
            if forest is None and len(root.children) > 0:
                raise ValueError("constructSearchTree(root): the root was not built by this library's sequential search")
            state = root.state
        else:
            root, forest, state = None, None, state_or_root
        game = state.game
        seed = KartMCTS.random.getrandbits(63) if seed is None else seed
        plies = max(1, sum(max(0, state.state.finalSection - state.state.karts[i].section) for i in range(state.state.n_karts)))
        if forest is None:                              # constructSearchTree(state): a new tree
            forest = Forest(game, 1, 1 + max(iterations, reserve_iterations or 0) * plies)
            out = forest.search([state.state], iterations, seed)
        else:                                           # constructSearchTree(root): continue the device-resident tree (root reuse, HKA:265-283)
            done = root._forest_iterations
            if 1 + (done + iterations) * plies > forest.max_nodes:
                raise ValueError("constructSearchTree(root): the tree's node budget is exhausted; size the first call with reserve_iterations")
            out = forest.search(None, iterations, 0, fresh=np.zeros(1, np.int32))
        if int(out["status"][0]) == 2:
            raise ZeroDivisionError("UCTWeight divided by zero inside findLeaf (KartMCTS.cs:164)")
        new_root = KartMCTS._graph_from_device(game, forest, state)
        new_root._device_best = [out["best"][0, k] for k in range(int(out["n_best"][0]))]
        return new_root

    @staticmethod
    def constructSearchTree(state_or_root, T: float = 0.09, parallel: bool = False, seed: int | None = None,
                            max_iterations: int | None = None, reserve_iterations: int | None = None):
        """KartMCTS.cs:50-106.  parallel=False (the reference's callers): the sequential search on the device, `max_iterations`
        iterations (default T * iterations_per_second); `reserve_iterations` sizes the device tree for later constructSearchTree(root)
        calls on the result (root reuse, HierarchicalKartAgent.cs:265-283).  parallel=True: wall-clock budgeted, each iteration expands one leaf with a GPU
        rollout batch (processLeaf with R playouts per child)."""
        if not parallel:
            return KartMCTS._construct_sequential(state_or_root, T, seed, max_iterations, reserve_iterations)
        root = state_or_root if isinstance(state_or_root, KartMCTSNode) else KartMCTSNode(state_or_root)
        seed = KartMCTS.random.getrandbits(63) if seed is None else seed
        total, it = 0.0, 0
        while total < T and (max_iterations is None or it < max_iterations):
            t0 = time.perf_counter()
            leaf = KartMCTS._find_leaf(root)
            root.childrenAsRoot += KartMCTS._process_leaf(leaf, seed, it * KartMCTS.rollouts_per_leaf * abi.HK_MAX_ACTIONS)
            total += time.perf_counter() - t0
            it += 1
        return root

    @staticmethod
    def getBestStatesSequence(node):                    # KartMCTS.cs:108-122
        best = []
        try:
            while len(node.children) > 0:
                node = node.children[KartMCTS.upperConfidenceStrategy(node)]
                s = node.state.state
                if all(s.karts[i].section == s.lastCompletedSection for i in range(s.n_karts)):
                    best.append(node.state)
        except ZeroDivisionError:
            pass
        return best

    @staticmethod
    def NextGaussian(mean: float | None = None, standard_deviation: float | None = None, mn: float | None = None,
                     mx: float | None = None) -> float:
        """KartMCTS.cs:204-236 (host-side helpers used by HierarchicalKartAgent.cs:116,126)."""
        r = KartMCTS.random
        if mean is None:
            while True:
                v1 = 2.0 * r.random() - 1.0
                v2 = 2.0 * r.random() - 1.0
                s = v1 * v1 + v2 * v2
                if not (s >= 1.0 or s == 0.0):
                    break
            return float(np.float32(v1 * math.sqrt((-2.0 * math.log(s)) / s)))
        if mn is None:
            return float(np.float32(mean) + np.float32(r.gauss(0.0, 1.0)) * np.float32(standard_deviation))
        attempts = 0
        while True:
            x = KartMCTS.NextGaussian(mean, standard_deviation)
            attempts += 1
            if not ((x < mn or x > mx) and attempts < 10):
                break
        if attempts == 10 and (x < mn or x > mx):
            return mean
        return x
