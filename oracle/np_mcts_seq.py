"""Second, independently written restatement of the reference's SEQUENTIAL tree search (TEST INFRASTRUCTURE: only tests/ may import
this) — KartMCTS.constructSearchTree with parallel == false, read from the C# text (Assets/Karting/Scripts/AI/MCTS/KartMCTS.cs:18-38,
50-122, 162-201, 238-289), with Python objects, parent references and insertion-ordered dictionaries keyed by the action exactly as
the reference's KartMCTSNode / Dictionary<DiscreteKartAction, KartMCTSNode>.  It exists to check the tree bookkeeping of
oracle/hk_oracle_mcts.c (arrays and indices) against a literal transcription; game primitives come from a `game` object with the
interface of oracle.oracle.Game (up_next / next_moves / policy_moves / make_move / is_over) — the C oracle, or oracle/np_game.py's
numpy float32 game.  Random draws: the Philox streams documented in hk_oracle_mcts.c (mode 0)."""
import math

import numpy as np

from . import oracle as O
from . import structs as S


class KartMCTSNode:                                          # :18-38
    def __init__(self, state, parent=None, order=0):
        self.state = state
        self.parent = parent
        self.children = {}                                   # Dictionary<DiscreteKartAction, KartMCTSNode>: insertion ordered
        self.totalValue = np.float32(0.0)
        self.numEpisodes = 0
        self.childrenAsRoot = 0
        self.order = order                                   # creation index (not in the reference; lets tests line nodes up)


class SequentialSearch:
    def __init__(self, game, key: int):
        self.g = game
        self.key = key & 0xFFFFFFFFFFFFFFFF
        self.picks = 0
        self.iters = 0
        self.created = 0
        self.cdf = {c: O.policy_cdf(c) for c in range(1, S.HK_MAX_ACTIONS + 1)}

    # ---- random sources (mode 0 of hk_oracle_mcts.c) ---------------------------------------------------------------------------
    def _random_next(self, n):                               # random.Next(n) :169
        v = int(O.philox(self.key ^ 0x9E3779B97F4A7C15, self.picks & 0xFFFFFFFF, self.picks >> 32, 0, 0)[0]) % n
        self.picks += 1
        return v

    def _policy_index(self, cnt, ply):                       # :266-269 through the closed-form distribution
        u = int(O.philox(self.key, self.iters & 0xFFFFFFFF, self.iters >> 32, ply, 0)[0])
        return int(np.sum(self.cdf[cnt][:cnt - 1] <= np.uint32(u)))

    # ---- KartMCTS ------------------------------------------------------------------------------------------------------------------
    @staticmethod
    def UCTWeight(node):                                     # :162-165
        ratio = node.parent.numEpisodes // node.numEpisodes  # int / int: ZeroDivisionError = DivideByZeroException
        lg = np.float32(math.log(float(np.float32(ratio)))) if ratio > 0 else np.float32(-np.inf)
        return np.float32(node.totalValue / np.float32(node.numEpisodes)) + np.float32(1.0) * lg

    def upperConfidenceStrategy(self, node):                 # :167-192
        index = self._random_next(len(node.children))
        best = list(node.children.keys())[index]
        best_uct = self.UCTWeight(node.children[best])
        for key, child in node.children.items():
            node_uct = self.UCTWeight(child)
            if node_uct > best_uct:
                best_uct, best = node_uct, key
        return best

    def findLeaf(self, root):                                # :194-201
        while len(root.children) > 0 and len(root.children) == self.g.next_moves(root.state)[2]:
            root = root.children[self.upperConfidenceStrategy(root)]
        return root

    def simulate(self, leaf):                                # :238-278
        new_states, ply = 0, 0
        while True:
            over, scores = self.g.is_over(leaf.state)
            if over:
                return leaf, scores, new_states
            state = leaf.state
            nextActions, _, cnt = self.g.policy_moves(state)  # nextMoves().OrderBy(...).ThenBy...(...) :256
            index = self._policy_index(cnt, ply)
            move = nextActions[index]
            if move not in leaf.children:
                self.created += 1
                leaf.children[move] = KartMCTSNode(self.g.make_move(state, move), leaf, self.created)
                new_states += 1
            leaf = leaf.children[move]
            ply += 1

    def backpropagate(self, node, result):                   # :280-289
        while node is not None:
            up = self.g.up_next(node.state)
            if 0 <= up < len(result):
                node.totalValue = np.float32(node.totalValue + np.float32(result[up]))
            node.numEpisodes += 1
            node = node.parent

    def constructSearchTree(self, root_or_state, iterations):   # :50-78 / :80-106, parallel == false, iteration budget
        root = root_or_state if isinstance(root_or_state, KartMCTSNode) else KartMCTSNode(S.game_state(root_or_state))
        for _ in range(iterations):
            leaf = self.findLeaf(root)
            end, result, new_states = self.simulate(leaf)
            root.childrenAsRoot += new_states
            self.backpropagate(end, result)
            self.iters += 1
        return root

    def getBestStatesSequence(self, node):                   # :108-122
        bestStates = []
        try:
            while len(node.children) > 0:
                node = node.children[self.upperConfidenceStrategy(node)]
                s = node.state
                if all(s.karts[i].section == s.lastCompletedSection for i in range(s.n_karts)):
                    bestStates.append(s)
        except ZeroDivisionError:
            pass
        return bestStates


def flatten(root):
    """Nodes in creation order with the fields hk_oracle_tree_dump reports."""
    nodes = []
    stack = [root]
    while stack:
        n = stack.pop()
        nodes.append(n)
        stack.extend(n.children.values())
    nodes.sort(key=lambda n: n.order)
    index = {id(n): i for i, n in enumerate(nodes)}
    return dict(parent=np.array([index[id(n.parent)] if n.parent is not None else -1 for n in nodes], np.int32),
                totalValue=np.array([n.totalValue for n in nodes], np.float32),
                numEpisodes=np.array([n.numEpisodes for n in nodes], np.int32),
                n_children=np.array([len(n.children) for n in nodes], np.int32),
                first_child=np.array([index[id(next(iter(n.children.values())))] if n.children else -1 for n in nodes], np.int32),
                actions=[None if n.parent is None else next(k for k, v in n.parent.children.items() if v is n) for n in nodes])
