"""Host-side mirror of the closed loop without PhysX over the C-ABI.

What the reference does per FixedUpdate and agent — `planFixed` (Assets/Karting/Scripts/AI/HierarchicalKartAgent.cs:145-166),
`SolveLQR` (:699-1224: problem recipe, `KartLQR.solveFeedbackLQR`, actuator map) and `OnTriggerEnter` (:611-662) — with
Unity's PhysX kart replaced by the kinematic model the planners assume (MPC/KartMPCDynamics.cs:55-70).  Everything runs in
libhk_b200.so (hk_race.cu); this module only builds the inputs and wraps the calls.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from .scenarios import DT
from .tracks import KART_COMPETE, Track

COASTING_DRAG = 5.0          # KartClassic_Player.prefab:204-245
MAX_STEER, MIN_STEER = 4.0, 1.0


def steer_for_wear(wear: float) -> float:
    """m_FinalStats.Steer for a tyre-wear proportion (ArcadeKart.cs:300-306): wear 0.25 in Experiment mode."""
    return MAX_STEER - wear * (MAX_STEER - MIN_STEER)


def race_params(track: Track, high_mode_mcts: bool = False, laps: int | None = None, bucket: int = 2, depth: int = 8,
                horizon: int = 3) -> abi.hk_race_params:
    accel, braking, top = KART_COMPETE[0], KART_COMPETE[1], KART_COMPETE[2]
    return abi.hk_race_params(dt=DT, accel=accel, braking=braking, coastingDrag=COASTING_DRAG, topSpeed=top,
                              gateHalfWidth=track.rows[0][3], maxLaneChanges=track.max_lane_changes,
                              goalSection=(track.laps if laps is None else laps) * track.n_sections,
                              highModeMcts=int(high_mode_mcts), velocityBucketSize=bucket, treeSearchDepth=depth,
                              planEvery=100, horizon=horizon)


def geometry(track: Track):
    """(sections, trigger_xz [n][2], forward_xz [n][2], lane_xz [n][4][2]) — SURVEY.md Appendix C."""
    head = track.heading_table()
    fwd = np.ascontiguousarray(np.stack([np.cos(head), np.sin(head)], axis=-1))
    return track.sections_array(), np.ascontiguousarray(track.trigger_table()), fwd, np.ascontiguousarray(track.lane_table())


def start_grid(track: Track, n_races: int, seed: int, wear: float = 0.25, jitter: float = 0.3):
    """Race/Experiment-mode start (RacingEnvController.cs:526-527: lanes {2,3} at section 0, tyre wear 0.25 :501-502) with a
    seeded position / heading / speed jitter so that the races differ.  Returns (karts [n_races][2], plans [n_races][2])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    karts = np.zeros((n_races, 2), dtype=abi.RACE_KART_DTYPE)
    plans = np.zeros((n_races, 2), dtype=abi.RACE_PLAN_DTYPE)
    lanes_xy, head = track.lane_table(), track.heading_table()
    for e, lane in enumerate((2, 3)):
        p0 = lanes_xy[0, lane - 1]
        fwd = np.array([np.cos(head[0]), np.sin(head[0])])
        along = rng.uniform(0.5, 2.5, size=n_races)
        karts["x"][:, e] = p0[0] + fwd[0] * along + rng.uniform(-jitter, jitter, size=n_races)
        karts["z"][:, e] = p0[1] + fwd[1] * along + rng.uniform(-jitter, jitter, size=n_races)
        karts["v"][:, e] = rng.uniform(0.0, 3.0, size=n_races)
        h = head[0] + rng.normal(0.0, 0.05, size=n_races)
        karts["h"][:, e] = np.mod(h, 2 * np.pi)
        karts["lane"][:, e] = lane
    karts["steer"] = steer_for_wear(wear)
    karts["section"] = 0
    karts["active"] = 1
    return karts, plans


class Races:
    """A batch of independent 2-kart races on one track."""

    def __init__(self, track: Track, params: abi.hk_race_params | None = None):
        self.track = track
        self.params = params if params is not None else race_params(track)
        self._sections, self._trig, self._fwd, self._lane = geometry(track)
        self._h = C.c_void_p()
        lib = abi.load_library()
        abi.check(lib.hk_track_create(self._sections, abi.dptr(self._trig), abi.dptr(self._fwd), abi.dptr(self._lane), track.n_sections,
                                      C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            abi.load_library().hk_track_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def recipe(self, karts: np.ndarray, plans: np.ndarray) -> dict:
        """Compact LQNG problems [2 n_races] (problem = 2 race + ego) as hk_lqng_assemble_solve_batch takes them."""
        n_races = karts.shape[0]
        nb = 2 * n_races
        out = dict(x0=np.zeros((nb, 2, 4)), target=np.zeros((nb, 2, 4)), tw=np.zeros((nb, 2, 4)), cw=np.zeros((nb, 2)),
                   aw=np.zeros((nb, 2, 1, 2)), otgt=np.zeros((nb, 2, 1, 4)), otw=np.zeros((nb, 2, 1, 3)))
        abi.check(abi.load_library().hk_race_recipe(self._h, C.byref(self.params), n_races, abi.vptr(karts), abi.vptr(plans),
                                                    *(abi.dptr(out[k]) for k in ("x0", "target", "tw", "cw", "aw", "otgt", "otw"))))
        out["dt"] = self.params.dt
        return out

    def plan_fixed(self, karts: np.ndarray, plans: np.ndarray) -> None:
        abi.check(abi.load_library().hk_race_plan_fixed(self._h, C.byref(self.params), karts.size, abi.vptr(karts), abi.vptr(plans)))

    def step(self, karts: np.ndarray, plans: np.ndarray, u: np.ndarray, episode_step: int) -> None:
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(karts.size, 2)
        abi.check(abi.load_library().hk_race_step(self._h, C.byref(self.params), karts.size, episode_step, abi.dptr(u), abi.vptr(karts),
                                                  abi.vptr(plans)))

    def run(self, karts: np.ndarray, plans: np.ndarray, first_step: int, n_steps: int):
        """n_steps of the full loop on the GPU, in place. Returns (u_last [n_races][2][2], n_lqng_status_nonzero)."""
        n_races = karts.shape[0]
        u = np.zeros((n_races, 2, 2))
        bad = C.c_int64(0)
        abi.check(abi.load_library().hk_race_run(self._h, C.byref(self.params), n_races, first_step, n_steps, abi.vptr(karts),
                                                 abi.vptr(plans), abi.dptr(u), C.byref(bad)))
        return u, bad.value


    def run_planned(self, karts: np.ndarray, plans: np.ndarray, planner: "Planner", first_step: int, n_steps: int):
        """hk_race_run_planned: the loop with the MCTS high level on the GPU and the planner's state carried between calls."""
        n_races = karts.shape[0]
        u = np.zeros((n_races, 2, 2))
        bad = C.c_int64(0)
        abi.check(abi.load_library().hk_race_run_planned(self._h, C.byref(self.params), planner._h, n_races, first_step, n_steps,
                                                         abi.vptr(karts), abi.vptr(plans), abi.dptr(u), C.byref(bad)))
        return u, bad.value

    def run_device(self, d_karts, d_plans, first_step: int, n_steps: int, planner: "Planner | None" = None, d_u=None, stream: int = 0) -> int:
        """hk_race_run_device: the loop on race states that stay in device memory.  d_karts / d_plans: objects with data_ptr() (torch uint8
        tensors holding the hk_race_kart / hk_race_plan records, e.g. device_state()) or raw device addresses; d_u: [n_agents][4] doubles
        or None; stream: a cudaStream_t handle (0 = the calling thread's own).  Returns the number of solves with a zero pivot."""
        ptr = lambda x: None if x is None else C.c_void_p(x.data_ptr() if hasattr(x, "data_ptr") else int(x))
        n_races = self._n_races_of(d_karts, C.sizeof(abi.hk_race_kart) * 2)
        bad = C.c_int64(0)
        abi.check(abi.load_library().hk_race_run_device(self._h, C.byref(self.params), planner._h if planner is not None else None, n_races,
                                                        first_step, n_steps, ptr(d_karts), ptr(d_plans), ptr(d_u), C.byref(bad),
                                                        C.c_void_p(stream) if stream else None))
        return bad.value

    @staticmethod
    def _n_races_of(d_karts, bytes_per_race: int) -> int:
        n = d_karts.numel() * d_karts.element_size()
        assert n % bytes_per_race == 0
        return n // bytes_per_race

    def run_mcts(self, karts: np.ndarray, plans: np.ndarray, game, iterations: int, rollouts_per_leaf: int, seed: int, first_step: int,
                 n_steps: int):
        """hk_race_run_mcts: the loop with the MCTS high level entirely on the GPU (root states, tree search, waypoint hand-off between
        two steps whenever episodeSteps % planEvery == 0).  `game` is a hierarchicalkarting_b200.mcts.Game for this track."""
        n_races = karts.shape[0]
        u = np.zeros((n_races, 2, 2))
        bad = C.c_int64(0)
        abi.check(abi.load_library().hk_race_run_mcts(self._h, C.byref(self.params), game._h, iterations, rollouts_per_leaf, seed, n_races,
                                                      first_step, n_steps, abi.vptr(karts), abi.vptr(plans), abi.dptr(u), C.byref(bad)))
        return u, bad.value


def start_grid_n(track: Track, n_races: int, karts_per_race: int, seed: int, teams=None, wear: float = 0.25, jitter: float = 0.3):
    """Race / Experiment-mode start of the Duos scenes (RacingEnvController.cs:526-527: lanes {2,3,2,3} at sections {0,0,1,1}, tyre wear
    0.25) for races of `karts_per_race` karts; teams default to [0, 0, 1, 1][:K].  Returns (karts [n_races][K], plans [n_races][K],
    beliefs [n_races][K][K], u_hold [n_races][K][2])."""
    K = karts_per_race
    teams = [0, 0, 1, 1][:K] if teams is None else list(teams)
    rng = np.random.Generator(np.random.PCG64(seed))
    karts = np.zeros((n_races, K), dtype=abi.RACE_KART_DTYPE)
    plans = np.zeros((n_races, K), dtype=abi.RACE_PLAN_DTYPE)
    beliefs = np.zeros((n_races, K, K), dtype=abi.RACE_BELIEF_DTYPE)
    lanes_xy, head = track.lane_table(), track.heading_table()
    for e in range(K):
        lane, sec = (2, 3, 2, 3)[e], (0, 0, 1, 1)[e]
        p0 = lanes_xy[sec, lane - 1]
        fwd = np.array([np.cos(head[sec]), np.sin(head[sec])])
        along = rng.uniform(0.5, 2.5, size=n_races)
        karts["x"][:, e] = p0[0] + fwd[0] * along + rng.uniform(-jitter, jitter, size=n_races)
        karts["z"][:, e] = p0[1] + fwd[1] * along + rng.uniform(-jitter, jitter, size=n_races)
        karts["v"][:, e] = rng.uniform(0.0, 3.0, size=n_races)
        karts["h"][:, e] = np.mod(head[sec] + rng.normal(0.0, 0.05, size=n_races), 2 * np.pi)
        karts["lane"][:, e] = lane
        karts["section"][:, e] = sec
        karts["team"][:, e] = teams[e]
    karts["steer"] = steer_for_wear(wear)
    karts["active"] = 1
    return karts, plans, beliefs, np.zeros((n_races, K, 2))


class RacesN(Races):
    """Races of K = 2..4 karts with teams (hk_raceN_*): SolveLQR's more-than-two-agents branches (8 m nearby filter, N in 1..4 players per
    problem, teammates, private / joint ordering), the solve every `lqr_every`-th step, MCTS with the game's team scoring."""

    def __init__(self, track: Track, params: abi.hk_race_params | None = None, karts_per_race: int = 4, lqr_every: int | None = None):
        super().__init__(track, params)
        self.K = karts_per_race
        self.lqr_every = lqr_every if lqr_every is not None else (4 if karts_per_race > 2 else 1)     # HierarchicalKartAgent.cs:317

    def recipe_n(self, karts: np.ndarray, plans: np.ndarray, beliefs: np.ndarray) -> dict:
        n = karts.size
        out = dict(n_players=np.zeros(n, np.int32), players=np.zeros((n, 4), np.int32), x0=np.zeros((n, 4, 4)), target=np.zeros((n, 4, 4)),
                   tw=np.zeros((n, 4, 4)), cw=np.zeros((n, 4)), aw=np.zeros((n, 4, 3, 2)), otgt=np.zeros((n, 4, 3, 4)), otw=np.zeros((n, 4, 3, 3)))
        abi.check(abi.load_library().hk_raceN_recipe(self._h, C.byref(self.params), self.K, karts.shape[0], abi.vptr(karts), abi.vptr(plans),
                                                     abi.vptr(beliefs), abi.vptr(out["n_players"]), abi.vptr(out["players"]),
                                                     *(abi.vptr(out[k]) for k in ("x0", "target", "tw", "cw", "aw", "otgt", "otw"))))
        out["dt"] = self.params.dt
        return out

    def planner(self, game, n_races: int, iterations: int, seed: int = 0, **kw) -> "Planner":
        return Planner(game, n_races, iterations, seed, karts_per_race=self.K, n_api=True, **kw)

    def run_n(self, karts, plans, beliefs, u_hold, first_step: int, n_steps: int, planner: "Planner | None" = None):
        bad = C.c_int64(0)
        abi.check(abi.load_library().hk_raceN_run(self._h, C.byref(self.params), planner._h if planner is not None else None, self.K, self.lqr_every,
                                                  karts.shape[0], first_step, n_steps, abi.vptr(karts), abi.vptr(plans), abi.vptr(beliefs),
                                                  abi.vptr(u_hold), C.byref(bad)))
        return bad.value

    def run_n_device(self, d_karts, d_plans, d_beliefs, d_u_hold, first_step: int, n_steps: int, planner: "Planner | None" = None, stream: int = 0) -> int:
        """hk_raceN_run_device: the loop on states that stay in device memory (torch tensors from device_state(); d_u_hold: float64 tensor
        [n_agents][8], zeros at the start of a race)."""
        n_races = d_karts.numel() * d_karts.element_size() // (C.sizeof(abi.hk_race_kart) * self.K)
        bad = C.c_int64(0)
        abi.check(abi.load_library().hk_raceN_run_device(self._h, C.byref(self.params), planner._h if planner is not None else None, self.K,
                                                         self.lqr_every, n_races, first_step, n_steps, C.c_void_p(d_karts.data_ptr()),
                                                         C.c_void_p(d_plans.data_ptr()), C.c_void_p(d_beliefs.data_ptr()),
                                                         C.c_void_p(d_u_hold.data_ptr()), C.byref(bad), C.c_void_p(stream) if stream else None))
        return bad.value


def device_state(*arrays, device="cuda:0"):
    """Structured numpy arrays (karts, plans, beliefs, ...) as torch uint8 tensors on the device, for the *_device entries; back with
    host_state()."""
    import torch
    return [torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(device) for a in arrays]


def host_state(tensor, like: np.ndarray) -> np.ndarray:
    """A device tensor made by device_state() back as a numpy array with the dtype and shape of `like`."""
    return tensor.cpu().numpy().view(like.dtype).reshape(like.shape)


class Planner:
    """hk_race_planner: the MCTS high level of a batch of races — per agent the device-resident tree (currentRoot), CyclesRootProcessed
    and the pending result of a search (HierarchicalKartAgent.cs:172-283, 331-353, 660-661)."""

    def __init__(self, game, n_races: int, iterations: int, seed: int = 0, mode: int = 0, first_iterations: int = 0,
                 rollouts_per_leaf: int = 0, reuse_cycles: int = 3, apply_delay: int = 0, max_tree_nodes: int = 0, karts_per_race: int = 2,
                 n_api: bool = False):
        self.params = abi.hk_race_mcts_params(mode=mode, iterations=iterations, first_iterations=first_iterations,
                                              rollouts_per_leaf=rollouts_per_leaf, reuse_cycles=reuse_cycles, apply_delay=apply_delay, seed=seed,
                                              max_tree_nodes=max_tree_nodes)
        self.game, self.n_races, self.K = game, n_races, karts_per_race
        self._h = C.c_void_p()
        if not n_api:
            assert karts_per_race == 2
            abi.check(abi.load_library().hk_race_planner_create(game._h, C.byref(self.params), n_races, C.byref(self._h)))
        else:
            abi.check(abi.load_library().hk_raceN_planner_create(game._h, C.byref(self.params), karts_per_race, n_races, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            abi.load_library().hk_race_planner_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def state(self):
        """(root_valid, cycles, tree_status), each [n_races][karts_per_race]"""
        rv, cy, ts = (np.zeros((self.n_races, self.K), np.int32) for _ in range(3))
        abi.check(abi.load_library().hk_race_planner_state(self._h, abi.vptr(rv), abi.vptr(cy), abi.vptr(ts)))
        return rv, cy, ts


# ---- MCTS high level: root state and waypoint hand-off ----------------------------------------
def mcts_root_state(track: Track, params: abi.hk_race_params, karts_race: np.ndarray, plans_race: np.ndarray, ego: int,
                    section_window: int = 2, time_precision: int = 100):
    """The root DiscreteGameState `planWithMCTS` builds for agent `ego` of one race (HierarchicalKartAgent.cs:180-245).
    Returns (hk_game_state, nearby) where nearby[i] is the race-local index of game kart i."""
    me = karts_race[ego]
    nearby, initial, furthest = [], int(me["section"]), ego
    for a in range(karts_race.shape[0]):                                    # foreach agent in m_envController.Agents (:182)
        if abs(int(karts_race[a]["section"]) - int(me["section"])) < section_window:
            nearby.append(a)
            initial = max(initial, int(karts_race[a]["section"]))
            if initial == int(karts_race[a]["section"]):
                furthest = a
    st = abi.hk_game_state()
    st.n_karts = len(nearby)
    st.initialSection = st.lastCompletedSection = initial
    st.finalSection = initial + params.treeSearchDepth
    L = track.n_sections
    for i, a in enumerate(nearby):
        k = karts_race[a]
        t_at = 0
        if int(k["section"]) != initial:                                    # :211-214, float32 product then (int)
            d = int(plans_race[a]["sectionTimes"][int(k["section"]) % L]) - int(plans_race[furthest]["sectionTimes"][int(k["section"]) % L])
            t_at = int(np.float32(np.float32(d) * np.float32(0.02)) * np.float32(time_precision))
        wear = (np.float32(MAX_STEER) - np.float32(k["steer"])) / np.float32(MAX_STEER - MIN_STEER)
        st.karts[i] = abi.hk_kart_state(player=0, team=a, section=initial, timeAtSection=t_at, min_velocity=0,     # quirks B.6-1, B.6-2
                                        max_velocity=min(params.velocityBucketSize, int(params.topSpeed)), lane=int(k["lane"]),
                                        tireAge=int(np.float32(wear * np.float32(10000))), laneChanges=int(k["laneChanges"]), infeasible=0)
    return st, nearby


def apply_best_states(track: Track, karts_race: np.ndarray, plans_race: np.ndarray, ego: int, nearby, best_states) -> None:
    """The waypoint hand-off of FixedUpdate (HierarchicalKartAgent.cs:366-402): the ego's own lanes / velocities for sections
    beyond the next checkpoint, and its belief about the other karts' (2-kart races: one opponent table)."""
    L = track.n_sections
    sec = int(karts_race[ego]["section"])
    for gs in best_states:
        s = gs.state
        for i in range(s.n_karts):
            ks = s.karts[i]
            if nearby[i] == ego:
                if ks.section > sec + (0 if sec == 0 else 1):
                    plans_race[ego]["lane"][ks.section % L] = ks.lane
                    plans_race[ego]["vel"][ks.section % L] = ks.max_velocity
            else:
                plans_race[ego]["oppLane"][ks.section % L] = ks.lane
                plans_race[ego]["oppVel"][ks.section % L] = ks.max_velocity


def mcts_root_states_batch(track: Track, params: abi.hk_race_params, karts: np.ndarray, plans: np.ndarray, section_window: int = 2,
                           time_precision: int = 100):
    """`mcts_root_state` for every agent of every 2-kart race at once (vectorised; same arithmetic, float32 where the reference is
    float32).  Returns (roots [n_races][2] of abi.GAME_STATE_DTYPE, nearby [n_races][2][2]: race-local agent of game kart i, -1 = none)."""
    n_races, L = karts.shape[0], track.n_sections
    ar = np.arange(n_races)
    sec = karts["section"].astype(np.int64)
    stimes = plans["sectionTimes"]
    roots = np.zeros((n_races, 2), dtype=abi.GAME_STATE_DTYPE)
    nearby = np.full((n_races, 2, 2), -1, dtype=np.int32)
    wear = (np.float32(MAX_STEER) - karts["steer"].astype(np.float32)) / np.float32(MAX_STEER - MIN_STEER)
    tire = (wear * np.float32(10000)).astype(np.float32).astype(np.int32)
    for e in (0, 1):
        o = 1 - e
        near = np.abs(sec[:, o] - sec[:, e]) < section_window                   # foreach agent in m_envController.Agents (:182)
        initial = np.where(near, np.maximum(sec[:, 0], sec[:, 1]), sec[:, e])
        furthest = np.where(near, np.where(sec[:, 1] >= sec[:, 0], 1, 0), e)
        g = roots[:, e]
        g["n_karts"] = np.where(near, 2, 1)
        g["initialSection"] = initial
        g["lastCompletedSection"] = initial
        g["finalSection"] = initial + params.treeSearchDepth
        for slot in (0, 1):
            a = np.where(near, slot, e)
            valid = near | (slot == 0)
            sa = sec[ar, a]
            d = stimes[ar, a, sa % L].astype(np.int64) - stimes[ar, furthest, sa % L].astype(np.int64)
            t_at = ((d.astype(np.float32) * np.float32(0.02)).astype(np.float32) * np.float32(time_precision)).astype(np.float32).astype(np.int32)
            ks = g["karts"][:, slot]
            ks["team"] = np.where(valid, a, 0)
            ks["section"] = np.where(valid, initial, 0)
            ks["timeAtSection"] = np.where(valid & (sa != initial), t_at, 0)              # :211-214
            ks["max_velocity"] = np.where(valid, min(params.velocityBucketSize, int(params.topSpeed)), 0)   # quirks B.6-1, B.6-2
            ks["lane"] = np.where(valid, karts["lane"][ar, a], 0)
            ks["tireAge"] = np.where(valid, tire[ar, a], 0)
            ks["laneChanges"] = np.where(valid, karts["laneChanges"][ar, a], 0)
            nearby[:, e, slot] = np.where(valid, a, -1)
    return roots, nearby


def apply_best_states_batch(track: Track, karts: np.ndarray, plans: np.ndarray, nearby: np.ndarray, best: np.ndarray, n_best: np.ndarray) -> None:
    """`apply_best_states` for every agent at once: best [n_races][2][HK_MCTS_MAX_SEQ] of abi.GAME_STATE_DTYPE, n_best [n_races][2]."""
    n_races, L = karts.shape[0], track.n_sections
    ar = np.arange(n_races)
    for e in (0, 1):
        sec = karts["section"][:, e].astype(np.int64)
        bound = sec + np.where(sec == 0, 0, 1)
        pl = plans[:, e]
        for k in range(int(n_best[:, e].max()) if n_races else 0):
            gs = best[:, e, k]
            for slot in (0, 1):
                ks = gs["karts"][:, slot]
                live = (k < n_best[:, e]) & (slot < gs["n_karts"])
                who = nearby[:, e, slot]
                key = ks["section"].astype(np.int64) % L
                m = live & (who == e) & (ks["section"] > bound)
                pl["lane"][ar[m], key[m]] = ks["lane"][m]
                pl["vel"][ar[m], key[m]] = ks["max_velocity"][m]
                m = live & (who != e) & (who >= 0)
                pl["oppLane"][ar[m], key[m]] = ks["lane"][m]
                pl["oppVel"][ar[m], key[m]] = ks["max_velocity"][m]


def plan_mcts_batch(track: Track, params: abi.hk_race_params, game, karts: np.ndarray, plans: np.ndarray, iterations: int,
                    rollouts_per_leaf: int, seed: int = 0):
    """planWithMCTS + the waypoint hand-off for EVERY agent of every race in one GPU tree-search launch (hk_mcts_search_batch: one
    thread block per agent's tree).  Root of agent (race r, ego e) is root index 2 r + e.  Returns the search result dict."""
    n_races = karts.shape[0]
    roots, nearby = mcts_root_states_batch(track, params, karts, plans)
    out = game.search_batch_array(np.ascontiguousarray(roots.reshape(-1)), iterations, rollouts_per_leaf, seed)
    apply_best_states_batch(track, karts, plans, nearby, out["best"].reshape(n_races, 2, abi.HK_MCTS_MAX_SEQ), out["n_best"].reshape(n_races, 2))
    return out


def plan_with_mcts(track: Track, params: abi.hk_race_params, game, karts_race: np.ndarray, plans_race: np.ndarray, ego: int,
                   T: float = 0.9, max_iterations: int | None = None, seed: int | None = None, parallel: bool = False):
    """planWithMCTS + hand-off for one agent: root state, KartMCTS.constructSearchTree (GPU leaf-parallel rollouts),
    getBestStatesSequence, apply.  `game` is a hierarchicalkarting_b200.mcts.Game for this track.  Returns the search root."""
    from . import mcts as M
    st, nearby = mcts_root_state(track, params, karts_race, plans_race, ego)
    root = M.KartMCTS.constructSearchTree(M.DiscreteGameState(game, st), T=T, seed=seed, max_iterations=max_iterations, parallel=parallel)
    best = M.KartMCTS.getBestStatesSequence(root)
    apply_best_states(track, karts_race, plans_race, ego, nearby, best)
    return root, best
