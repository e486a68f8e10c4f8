"""CPU restatement of the reference's tree search for the oracle side of the parity tests (TEST INFRASTRUCTURE: only tests/ may import
this).  KartMCTS.constructSearchTree / findLeaf / upperConfidenceStrategy / UCTWeight / backpropagate / getBestStatesSequence
(Assets/Karting/Scripts/AI/MCTS/KartMCTS.cs:50-122, 162-201, 280-289) in the reference's own leaf-parallel form (processLeaf, :124-159:
every legal child of the selected leaf is created and played), with `rollouts_per_leaf` playouts of simulate (:238-278) through each
child — every game primitive and every playout comes from the C oracle (oracle/hk_oracle_game.c), nothing from the CUDA library.

Streams (what hk_mcts_search_batch documents in include/hk_abi.h): root r of a batch plays rollout k of child j in iteration `it`
as Philox id it * R * HK_MAX_ACTIONS + j * R + k under key seed + r; the random initial pick of upperConfidenceStrategy (:169, it only
decides exact ties) is word 0 of Philox4x32-10 with key (seed + r) ^ 0x9E3779B97F4A7C15 and counter = number of picks so far."""
import math

import numpy as np

from . import oracle as O
from . import structs as S

HK_MAX_ACTIONS = S.HK_MAX_ACTIONS


class Node:                                              # KartMCTSNode, KartMCTS.cs:18-38
    def __init__(self, state, parent=None):
        self.state, self.parent, self.children = state, parent, []
        self.totalValue, self.numEpisodes = 0.0, 0


def _copy(st):
    return S.game_state(st)


class TreeSearch:
    def __init__(self, game: "O.Game", rseed: int):
        self.g, self.rseed, self.picks = game, rseed, 0
        self.key = (rseed ^ 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF

    def _uct(self, node):                                # UCTWeight :162-165: integer division, float32 arithmetic, no sqrt
        if node.numEpisodes == 0:
            raise ZeroDivisionError
        ratio = node.parent.numEpisodes // node.numEpisodes
        lg = np.float32(math.log(ratio)) if ratio > 0 else np.float32(-np.inf)
        return np.float32(np.float32(node.totalValue) / np.float32(node.numEpisodes)) + lg

    def _pick(self, node):                               # upperConfidenceStrategy :167-192
        n = len(node.children)
        best = int(O.philox(self.key, self.picks & 0xFFFFFFFF, self.picks >> 32, 0, 0)[0]) % n
        self.picks += 1
        best_w = self._uct(node.children[best])
        for j, c in enumerate(node.children):
            w = self._uct(c)
            if w > best_w:
                best_w, best = w, j
        return best

    def search(self, root_state, iterations: int, R: int):
        root = Node(_copy(root_state))
        n_nodes = 1
        for it in range(iterations):
            node = root                                  # findLeaf :194-201
            while node.children:
                node = node.children[self._pick(node)]
            up = self.g.up_next(node.state)
            moves, _, n_moves = self.g.next_moves(node.state)
            over, scores = self.g.is_over(node.state)
            if over:                                     # simulate() returns at once (:246-249)
                n = node
                while n is not None:
                    u = self.g.up_next(n.state)
                    if 0 <= u < len(scores):
                        n.totalValue += float(scores[u])
                    n.numEpisodes += 1
                    n = n.parent
                continue
            for mv in moves:                             # processLeaf :142: every legal child, generation order
                node.children.append(Node(self.g.make_move(node.state, mv), node))
            n_nodes += len(moves)
            offset = it * R * HK_MAX_ACTIONS
            for j, c in enumerate(node.children):
                st = self.g.rollouts(c.state, R, mode=0, seed=self.rseed, rollout_offset=offset + j * R)
                visits = int(st["visit"].sum())
                if visits == 0:                          # terminal child: its own scores, R times
                    _, sc = self.g.is_over(c.state)
                    contrib, count = [float(x) * R for x in sc], R
                else:
                    contrib, count = list(st["reward_sum"].sum(axis=0)), visits - int(st["nan_count"].sum())
                n = c
                while n is not None:                     # backpropagate :280-289
                    u = self.g.up_next(n.state)
                    if 0 <= u < min(len(contrib), S.HK_MAX_KARTS):
                        n.totalValue += float(contrib[u])
                    n.numEpisodes += count
                    n = n.parent
        best, node = [], root                            # getBestStatesSequence :108-122
        try:
            while node.children:
                node = node.children[self._pick(node)]
                s = node.state
                if all(s.karts[i].section == s.lastCompletedSection for i in range(s.n_karts)):
                    best.append(s)
        except ZeroDivisionError:
            pass
        return root, best, n_nodes
