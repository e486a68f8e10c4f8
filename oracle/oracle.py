"""ctypes binding of oracle/libhk_oracle.so — the CPU ORACLE (test infrastructure, NOT the product).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import structs as S

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhk_oracle.so")
_lib = None
_dp, _ip, _lp, _fp, _up = (C.POINTER(t) for t in (C.c_double, C.c_int32, C.c_int64, C.c_float, C.c_uint32))
_vp = C.c_void_p


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("hk_oracle_lqng.c", "hk_oracle_game.c", "hk_oracle_mcts.c", "hk_oracle_race.c", "hk_oracle.h", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.hk_oracle_bicycle_A.argtypes = [C.c_double, _dp, _dp]
        L.hk_oracle_bicycle_B.argtypes = [C.c_double, _dp]
        L.hk_oracle_cost.argtypes = [C.c_int, _dp, _dp, C.c_double, _dp, _dp, _dp, _dp, _dp, _dp]
        L.hk_oracle_lqng_solve.argtypes = [C.c_int, C.c_int, C.c_int] + [_dp] * 10
        L.hk_oracle_lqng_solve_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int] + [_dp] * 10 + [_ip, C.c_int]
        L.hk_oracle_game_create.argtypes = [_vp, C.c_int, _vp, C.c_int,
                                            _vp, C.c_int, _vp, C.POINTER(C.c_void_p)]
        L.hk_oracle_game_destroy.argtypes = [C.c_void_p]
        L.hk_oracle_max_speed_for_radius_and_wear.argtypes = [_vp, C.c_float, C.c_float]
        L.hk_oracle_max_speed_for_radius_and_wear.restype = C.c_float
        L.hk_oracle_compute_toc.argtypes = [_vp] + [C.c_float] * 5
        L.hk_oracle_compute_toc.restype = C.c_float
        L.hk_oracle_apply_action.argtypes = [C.c_void_p, _vp, S.hk_action]
        L.hk_oracle_apply_action.restype = S.hk_kart_state
        L.hk_oracle_up_next.argtypes = [C.c_void_p, _vp]
        L.hk_oracle_next_moves.argtypes = [C.c_void_p, _vp, _vp, _ip]
        L.hk_oracle_make_move.argtypes = [C.c_void_p, _vp, S.hk_action, _ip]
        L.hk_oracle_make_move.restype = S.hk_game_state
        L.hk_oracle_is_over.argtypes = [C.c_void_p, _vp, _fp, _ip]
        L.hk_oracle_policy_moves.argtypes = [C.c_void_p, _vp, _vp, _ip]
        L.hk_oracle_philox4x32_10.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _up]
        L.hk_oracle_policy_cdf.argtypes = [C.c_int, _up]
        L.hk_oracle_policy_index.argtypes = [C.c_int, _up, C.c_uint32]
        L.hk_oracle_reference_policy_index.argtypes = [C.c_int, C.POINTER(C.c_uint64)]
        L.hk_oracle_rollout.argtypes = [C.c_void_p, _vp, C.c_int, C.c_uint64, C.c_uint64,
                                        C.POINTER(C.c_uint64), _vp, _ip, _fp, _ip, _vp]
        L.hk_oracle_rollouts.argtypes = [C.c_void_p, _vp, C.c_int64, C.c_int, C.c_uint64, C.c_uint64,
                                         _lp, _dp, _lp, _lp]
        for f in ("hk_oracle_distance_to_travel", "hk_oracle_radius_of_lane"):
            getattr(L, f).argtypes = [_vp, C.c_int, C.c_int]
            getattr(L, f).restype = C.c_float
        L.hk_oracle_tire_load.argtypes = [_vp, C.c_float, C.c_int, C.c_int]
        L.hk_oracle_tire_load.restype = C.c_float
        L.hk_oracle_tree_create.argtypes = [_vp, _vp]
        L.hk_oracle_tree_create.restype = C.c_void_p
        L.hk_oracle_tree_destroy.argtypes = [_vp]
        L.hk_oracle_tree_search.argtypes = [_vp, C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_uint64)]
        L.hk_oracle_tree_best_states.argtypes = [_vp, C.c_int, C.c_uint64, C.POINTER(C.c_uint64), _vp, C.c_int]
        L.hk_oracle_tree_size.argtypes = [_vp]
        L.hk_oracle_tree_children_as_root.argtypes = [_vp]
        L.hk_oracle_tree_children_as_root.restype = C.c_longlong
        L.hk_oracle_tree_dump.argtypes = [_vp] + [_vp] * 8
        L.hk_oracle_tree_search_batch.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_uint64, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


# ---- LQNG ------------------------------------------------------------------------------------------------------
def bicycle_A(dt: float, x0) -> np.ndarray:
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    A = np.empty((4, 4))
    lib().hk_oracle_bicycle_A(dt, _p(x0), _p(A))
    return A


def bicycle_B(dt: float) -> np.ndarray:
    B = np.empty((4, 2))
    lib().hk_oracle_bicycle_B(dt, _p(B))
    return B


def cost(target, tw, cw, aw, otgt, otw):
    aw = np.ascontiguousarray(aw, dtype=np.float64).reshape(-1, 2)
    n_other = aw.shape[0]
    n = 4 * (1 + n_other)
    target, tw = (np.ascontiguousarray(v, dtype=np.float64) for v in (target, tw))
    otgt = np.ascontiguousarray(otgt, dtype=np.float64).reshape(n_other, 4)
    otw = np.ascontiguousarray(otw, dtype=np.float64).reshape(n_other, 3)
    Q, q, R = np.empty((n, n)), np.empty(n), np.empty((2, 2))
    lib().hk_oracle_cost(n_other, _p(target), _p(tw), float(cw), _p(aw), _p(otgt), _p(otw), _p(Q), _p(q), _p(R))
    return Q, q, R


def lqng_solve_batch(A, B, Q, q, R, x0, horizon: int, time_varying: bool = False, threads: int = 1, full: bool = True):
    """Arrays in the hk_abi.h layout with a leading batch dimension. Returns dict(u0, P, alpha, traj, status)."""
    A, B, Q, q, R, x0 = (np.ascontiguousarray(v, dtype=np.float64) for v in (A, B, Q, q, R, x0))
    batch, n = x0.shape
    N, m, T = n // 4, n // 2, horizon + 1
    u0 = np.empty((batch, m))
    P = np.empty((batch, T, m, n)) if full else None
    alpha = np.empty((batch, T, m)) if full else None
    traj = np.empty((batch, T + 1, n)) if full else None
    status = np.zeros(batch, dtype=np.int32)
    lib().hk_oracle_lqng_solve_batch(batch, N, horizon, int(time_varying), _p(A), _p(B), _p(Q), _p(q), _p(R), _p(x0),
                                     _p(u0), _p(P), _p(alpha), _p(traj), status.ctypes.data_as(_ip), threads)
    return dict(u0=u0, P=P, alpha=alpha, traj=traj, status=status)


# ---- discrete game ---------------------------------------------------------------------------------------------
class Game:
    def __init__(self, sections, n_sections, karts, n_karts, params, env_karts=None, n_env_karts=0):
        self._h = C.c_void_p()
        rc = lib().hk_oracle_game_create(sections, n_sections, karts, n_karts, env_karts, n_env_karts, C.byref(params),
                                         C.byref(self._h))
        if rc != 0:
            raise ValueError("hk_oracle_game_create failed")
        self.params = params

    def __del__(self):
        if getattr(self, "_h", None):
            lib().hk_oracle_game_destroy(self._h)
            self._h = None

    def up_next(self, st) -> int:
        return lib().hk_oracle_up_next(self._h, C.byref(st))

    def apply_action(self, ks, a):
        return lib().hk_oracle_apply_action(self._h, C.byref(ks), S.action(a))

    def next_moves(self, st):
        mv = (S.hk_action * S.HK_MAX_ACTIONS)()
        gi = (C.c_int32 * S.HK_MAX_ACTIONS)()
        n = lib().hk_oracle_next_moves(self._h, C.byref(st), mv, gi)
        return [mv[i].astuple() for i in range(max(n, 0))], [gi[i] for i in range(max(n, 0))], n

    def policy_moves(self, st):
        mv = (S.hk_action * S.HK_MAX_ACTIONS)()
        gi = (C.c_int32 * S.HK_MAX_ACTIONS)()
        n = lib().hk_oracle_policy_moves(self._h, C.byref(st), mv, gi)
        return [mv[i].astuple() for i in range(max(n, 0))], [gi[i] for i in range(max(n, 0))], n

    def make_move(self, st, a):
        return lib().hk_oracle_make_move(self._h, C.byref(st), S.action(a), None)

    def is_over(self, st):
        sc = (C.c_float * (2 * S.HK_MAX_KARTS))()
        ns = C.c_int32(0)
        over = lib().hk_oracle_is_over(self._h, C.byref(st), sc, C.byref(ns))
        return over, np.array([sc[i] for i in range(ns.value)], dtype=np.float32)

    def rollout(self, leaf, mode=0, seed=0, rollout_id=0, rng_state=None):
        acts = (S.hk_action * S.HK_MAX_PLIES)()
        ch = (C.c_int32 * S.HK_MAX_PLIES)()
        sc = (C.c_float * (2 * S.HK_MAX_KARTS))()
        ns = C.c_int32(0)
        term = S.hk_game_state()
        rs = C.c_uint64(rng_state if rng_state else 88172645463325252)
        n = lib().hk_oracle_rollout(self._h, C.byref(leaf), mode, seed, rollout_id, C.byref(rs), acts, ch, sc, C.byref(ns),
                                    C.byref(term))
        return dict(n_plies=n, actions=[acts[i].astuple() for i in range(max(n, 0))], choices=[ch[i] for i in range(max(n, 0))],
                    scores=np.array([sc[i] for i in range(ns.value)], dtype=np.float32), terminal=term)

    def rollouts(self, leaf, n_rollouts, mode=0, seed=0, rollout_offset=0):
        visit = np.zeros(S.HK_MAX_ACTIONS, dtype=np.int64)
        rsum = np.zeros((S.HK_MAX_ACTIONS, S.HK_MAX_KARTS))
        nanc = np.zeros(S.HK_MAX_ACTIONS, dtype=np.int64)
        plies = C.c_int64(0)
        rc = lib().hk_oracle_rollouts(self._h, C.byref(leaf), n_rollouts, mode, seed, rollout_offset,
                                      visit.ctypes.data_as(_lp), _p(rsum), nanc.ctypes.data_as(_lp), C.byref(plies))
        if rc != 0:
            raise RuntimeError("oracle rollouts failed")
        return dict(visit=visit, reward_sum=rsum, nan_count=nanc, plies=plies.value)


class Tree:
    """hk_oracle_mcts.c: the reference's sequential search (constructSearchTree with parallel == false, KartMCTS.cs:50-106) on
    one KartMCTSNode graph that survives between calls.  mode 0 = Philox streams shared with the CUDA library (key = seed + tree
    index), mode 1 = the reference's own random procedures from an xorshift state."""

    def __init__(self, game: Game, root, key: int = 0, mode: int = 0, rng_state: int = 88172645463325252):
        self.game, self.key, self.mode = game, key & 0xFFFFFFFFFFFFFFFF, mode
        self._rng = C.c_uint64(rng_state if rng_state else 1)
        self._root = S.game_state(root)
        self._h = C.c_void_p(lib().hk_oracle_tree_create(game._h, C.byref(self._root)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().hk_oracle_tree_destroy(self._h)
            self._h = None

    def search(self, iterations: int) -> int:
        return lib().hk_oracle_tree_search(self._h, iterations, self.mode, self.key, C.byref(self._rng))

    def best_states(self, max_out: int = S.HK_MCTS_MAX_SEQ):
        out = (S.hk_game_state * max_out)()
        n = lib().hk_oracle_tree_best_states(self._h, self.mode, self.key, C.byref(self._rng), out, max_out)
        return [S.hk_game_state.from_buffer_copy(bytes(out[i])) for i in range(n)]

    @property
    def size(self) -> int:
        return lib().hk_oracle_tree_size(self._h)

    @property
    def children_as_root(self) -> int:
        return lib().hk_oracle_tree_children_as_root(self._h)

    def dump(self, states: bool = False):
        n = self.size
        out = dict(parent=np.zeros(n, np.int32), gen=np.zeros(n, np.int32), totalValue=np.zeros(n, np.float32),
                   numEpisodes=np.zeros(n, np.int32), n_children=np.zeros(n, np.int32), first_child=np.zeros(n, np.int32),
                   next_sibling=np.zeros(n, np.int32))
        st = np.zeros(n, dtype=S.GAME_STATE_DTYPE) if states else None
        lib().hk_oracle_tree_dump(self._h, *(S.ref(out[k]) for k in ("parent", "gen", "totalValue", "numEpisodes", "n_children",
                                                                      "first_child", "next_sibling")), S.ref(st))
        if states:
            out["states"] = st
        return out


def tree_search_batch(game: Game, roots, iterations: int, seed: int = 0, mode: int = 0, rng_states=None, threads: int = 0):
    """hk_oracle_tree_search_batch: n independent sequential searches + getBestStatesSequence (tree r: key seed + r), OpenMP over
    trees.  roots: array of S.GAME_STATE_DTYPE-compatible records.  Same outputs as the product's search_seq_batch."""
    roots = np.ascontiguousarray(roots)
    n = roots.shape[0]
    assert roots.dtype.itemsize == C.sizeof(S.hk_game_state)
    best = np.zeros((n, S.HK_MCTS_MAX_SEQ), dtype=S.GAME_STATE_DTYPE)
    n_best, nodes = np.zeros(n, np.int32), np.zeros(n, np.int32)
    gen, eps = np.zeros((n, S.HK_MAX_ACTIONS), np.int32), np.zeros((n, S.HK_MAX_ACTIONS), np.int32)
    vals = np.zeros((n, S.HK_MAX_ACTIONS), np.float32)
    rs = None if rng_states is None else np.ascontiguousarray(rng_states, dtype=np.uint64)
    rc = lib().hk_oracle_tree_search_batch(game._h, S.ref(roots), n, iterations, mode, seed, S.ref(rs), S.ref(best), S.ref(n_best),
                                           S.HK_MCTS_MAX_SEQ, S.ref(gen), S.ref(eps), S.ref(vals), S.ref(nodes),
                                           threads if threads > 0 else (os.cpu_count() or 1))
    if rc != 0:
        raise RuntimeError(f"oracle tree search failed ({rc})")
    return dict(best=best, n_best=n_best, root_gen=gen, root_episodes=eps, root_values=vals, n_nodes=nodes)


def policy_cdf(cnt: int) -> np.ndarray:
    out = np.zeros(cnt, dtype=np.uint32)
    lib().hk_oracle_policy_cdf(cnt, out.ctypes.data_as(_up))
    return out


def philox(seed, c0, c1, c2, c3) -> np.ndarray:
    out = np.zeros(4, dtype=np.uint32)
    lib().hk_oracle_philox4x32_10(seed, c0, c1, c2, c3, out.ctypes.data_as(_up))
    return out


# ---- closed loop without PhysX (hk_oracle_race.c) ---------------------------------------------------------------
class Races:
    """Oracle twin of hierarchicalkarting_b200.race.Races (same array layouts)."""

    def __init__(self, sections, trig, fwd, lane, n_sections, params):
        self.sections, self.n = sections, n_sections
        self.trig, self.fwd, self.lane = (np.ascontiguousarray(a, dtype=np.float64) for a in (trig, fwd, lane))
        self.params = params
        L = lib()
        vp = C.c_void_p
        geo = [_vp, _dp, _dp, _dp, C.c_int, _vp]
        L.hk_oracle_race_recipe.argtypes = geo + [C.c_int, vp, vp] + [_dp] * 7
        L.hk_oracle_race_plan_fixed.argtypes = [_vp, C.c_int, _vp, C.c_int, vp, vp]
        L.hk_oracle_race_step.argtypes = geo + [C.c_int, C.c_int, _dp, vp, vp]
        L.hk_oracle_race_run.argtypes = geo + [C.c_int, C.c_int, C.c_int, vp, vp, _dp]
        L.hk_oracle_race_run.restype = C.c_longlong

    def _geo(self):
        return [self.sections, _p(self.trig), _p(self.fwd), _p(self.lane), self.n, C.byref(self.params)]

    def recipe(self, karts, plans):
        nb = 2 * karts.shape[0]
        out = dict(x0=np.zeros((nb, 2, 4)), target=np.zeros((nb, 2, 4)), tw=np.zeros((nb, 2, 4)), cw=np.zeros((nb, 2)),
                   aw=np.zeros((nb, 2, 1, 2)), otgt=np.zeros((nb, 2, 1, 4)), otw=np.zeros((nb, 2, 1, 3)))
        lib().hk_oracle_race_recipe(*self._geo(), karts.shape[0], S.ref(karts), S.ref(plans),
                                    *(_p(out[k]) for k in ("x0", "target", "tw", "cw", "aw", "otgt", "otw")))
        out["dt"] = self.params.dt
        return out

    def recipe_n_one(self, K, karts_r, plans_r, beliefs_e, e):
        """hk_oracle_raceN_recipe_one: the ego e's problem in one race of K karts (karts_r [K], plans_r [K], beliefs_e [K]: the ego's
        beliefs).  Arrays trimmed to the real players: x0/target/tw [n][4], cw [n], aw [n][n-1][2], otgt [n][n-1][4], otw [n][n-1][3]."""
        L = lib()
        L.hk_oracle_raceN_recipe_one.argtypes = [_vp, _dp, _dp, _dp, C.c_int, _vp, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                                 C.POINTER(C.c_int), C.POINTER(C.c_int)] + [_dp] * 7
        n = C.c_int(0)
        players = (C.c_int * 4)()
        o = dict(x0=np.zeros((4, 4)), target=np.zeros((4, 4)), tw=np.zeros((4, 4)), cw=np.zeros(4), aw=np.zeros((4, 3, 2)),
                 otgt=np.zeros((4, 3, 4)), otw=np.zeros((4, 3, 3)))
        kr, pr, br = (np.ascontiguousarray(a) for a in (karts_r, plans_r, beliefs_e))
        L.hk_oracle_raceN_recipe_one(*self._geo(), K, S.ref(kr), S.ref(pr), S.ref(br), e, C.byref(n), players,
                                     *(_p(o[k]) for k in ("x0", "target", "tw", "cw", "aw", "otgt", "otw")))
        N = n.value
        out = {k: o[k][:N] for k in ("x0", "target", "tw", "cw")}
        for k in ("aw", "otgt", "otw"):
            out[k] = o[k][:N, :max(N - 1, 0)]
        out["players"] = [players[i] for i in range(N)]
        return out

    def plan_fixed(self, karts, plans):
        lib().hk_oracle_race_plan_fixed(self.sections, self.n, C.byref(self.params), karts.size, S.ref(karts), S.ref(plans))

    def step(self, karts, plans, u, episode_step):
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(karts.size, 2)
        lib().hk_oracle_race_step(*self._geo(), karts.size, episode_step, _p(u), S.ref(karts), S.ref(plans))

    def mcts_root(self, game_params, karts_race, plans_race, ego):
        """planWithMCTS's root state for agent `ego` of one race (hk_oracle_race_mcts_root). Returns (hk_game_state, nearby list)."""
        L = lib()
        L.hk_oracle_race_mcts_root.argtypes = [_vp, _vp, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp, _ip]
        st = S.hk_game_state()
        nearby = (C.c_int32 * S.HK_MAX_KARTS)()
        n = L.hk_oracle_race_mcts_root(C.byref(self.params), C.byref(game_params), self.n, S.ref(karts_race), S.ref(plans_race),
                                       karts_race.shape[0], ego, C.byref(st), nearby)
        return st, [nearby[i] for i in range(n)]

    def apply_best(self, karts_race, plans_race, ego, nearby, best_states):
        """The waypoint hand-off (hk_oracle_race_apply_best); best_states: list of objects with hk_game_state's bytes."""
        L = lib()
        L.hk_oracle_race_apply_best.argtypes = [C.c_int, _vp, _vp, C.c_int, _ip, _vp, C.c_int]
        nb = len(best_states)
        arr = (S.hk_game_state * max(nb, 1))(*[S.game_state(b) for b in best_states])
        near = (C.c_int32 * S.HK_MAX_KARTS)(*(list(nearby) + [-1] * (S.HK_MAX_KARTS - len(nearby))))
        L.hk_oracle_race_apply_best(self.n, S.ref(karts_race), S.ref(plans_race), ego, near, arr, nb)

    def planner(self, game: "Game", game_params, n_races: int, iterations: int, seed: int = 0, first_iterations: int = 0,
                reuse_cycles: int = 3, apply_delay: int = 0):
        return Planner(self, game, game_params, n_races, S.hk_race_mcts_params(mode=0, iterations=iterations, first_iterations=first_iterations,
                                                                               rollouts_per_leaf=0, reuse_cycles=reuse_cycles,
                                                                               apply_delay=apply_delay, seed=seed))

    def run(self, karts, plans, first_step, n_steps):
        u = np.zeros((karts.shape[0], 2, 2))
        bad = lib().hk_oracle_race_run(*self._geo(), karts.shape[0], first_step, n_steps, S.ref(karts), S.ref(plans), _p(u))
        return u, int(bad)


class Planner:
    """hk_oracle_planner: the MCTS high level of the oracle's race loop (sequential search; schedule of HierarchicalKartAgent.cs:85-93,
    172-283, 331-353, 660-661)."""

    def __init__(self, races: Races, game: Game, game_params, n_races: int, mp):
        L = lib()
        L.hk_oracle_planner_create.argtypes = [_vp, _vp, _vp, C.c_int]
        L.hk_oracle_planner_create.restype = C.c_void_p
        L.hk_oracle_planner_destroy.argtypes = [_vp]
        L.hk_oracle_planner_state.argtypes = [_vp, _vp, _vp]
        L.hk_oracle_race_run_planned.argtypes = [_vp, _dp, _dp, _dp, C.c_int, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _dp]
        L.hk_oracle_race_run_planned.restype = C.c_longlong
        self.races, self.game, self.n_races, self.mp = races, game, n_races, mp
        self._h = C.c_void_p(L.hk_oracle_planner_create(game._h, C.byref(game_params), C.byref(mp), n_races))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().hk_oracle_planner_destroy(self._h)
            self._h = None

    def state(self):
        rv, cy = np.zeros((self.n_races, 2), np.int32), np.zeros((self.n_races, 2), np.int32)
        lib().hk_oracle_planner_state(self._h, S.ref(rv), S.ref(cy))
        return rv, cy

    def run(self, karts, plans, first_step, n_steps):
        u = np.zeros((karts.shape[0], 2, 2))
        R = self.races
        bad = lib().hk_oracle_race_run_planned(R.sections, _p(R.trig), _p(R.fwd), _p(R.lane), R.n, C.byref(R.params), self._h,
                                               karts.shape[0], first_step, n_steps, S.ref(karts), S.ref(plans), _p(u))
        if bad < 0:
            raise RuntimeError("oracle planned race loop: a tree search failed")
        return u, int(bad)
