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
