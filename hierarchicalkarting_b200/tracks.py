"""Track / kart / game-parameter tables the discrete game and the synthetic LQNG scenarios read.

Reference sources: section fields `DiscretePositionTracker.cs:35-40`; kart constants = effective `m_FinalStats` of the
Compete scenes (`KartClassic_Player.prefab:204-245`, `CompeteAgents-Oval.unity:4900-4926`); `gameParams` of the MCTS-LQR
agent (`CompeteAgents-Oval.unity:5160-5186`); environment fields (`CompeteAgents-Oval.unity:7405-7411`,
`CompeteAgents-Complex.unity:4442-4448`, `RacingEnvController.cs:89`).  See SURVEY.md Appendix C.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

from . import abi
from ._track_tables import COMPLEX_SECTIONS, OVAL_SECTIONS

# lane local x-offsets in the waypoint frame (Waypoint.prefab:30,471,376,186,281)
LANE_OFFSETS = (-3.5, -1.25, 1.25, 3.5)


@dataclasses.dataclass(frozen=True)
class Track:
    name: str
    rows: tuple
    max_lane_changes: int
    laps: int
    max_episode_steps: int = 6000

    @property
    def n_sections(self) -> int:
        return len(self.rows)

    def sections_array(self):
        arr = (abi.hk_section * self.n_sections)()
        for i, r in enumerate(self.rows):
            arr[i] = abi.hk_section(r[1], r[2], r[3], r[5], r[4], r[6])
        return arr

    def is_straight(self, s: int) -> bool:
        return self.rows[s % self.n_sections][1] == 0.0

    def trigger(self, s: int):
        return self.rows[s % self.n_sections][7]

    def heading(self, s: int) -> float:
        return math.radians(self.rows[s % self.n_sections][8])

    def lane_xy(self, s: int, lane: int):
        return self.rows[s % self.n_sections][9][lane - 1]

    def lane_table(self) -> np.ndarray:
        """[n_sections][4][2] lane collider positions (x, z).  Unity's transform.position is a float32 Vector3 and SolveLQR subtracts
        such positions in float32 (HierarchicalKartAgent.cs:819-831) before widening, so the table holds float32-representable values."""
        return np.array([[list(p) for p in r[9]] for r in self.rows], dtype=np.float32).astype(np.float64)

    def trigger_table(self) -> np.ndarray:
        return np.array([list(r[7]) for r in self.rows], dtype=np.float32).astype(np.float64)      # float32-representable, see lane_table

    def heading_table(self) -> np.ndarray:
        return np.radians(np.array([r[8] for r in self.rows], dtype=np.float64))

    def straight_table(self) -> np.ndarray:
        return np.array([r[1] == 0.0 for r in self.rows])


OVAL = Track("Oval", OVAL_SECTIONS, max_lane_changes=3, laps=4)
COMPLEX = Track("Complex", COMPLEX_SECTIONS, max_lane_changes=4, laps=3)
TRACKS = {"Oval": OVAL, "Complex": COMPLEX}

# accel, braking, topSpeed, reverseSpeed, maxGs, minGs, tireWearFactor
KART_COMPETE = (7.0, 16.0, 15.0, 10.0, 2.0, 0.5, 0.001)


def kart_array(n: int, consts=KART_COMPETE):
    arr = (abi.hk_kart * n)()
    for i in range(n):
        arr[i] = abi.hk_kart(*consts)
    return arr


def game_params(track: Track, bucket: int = 2, time_precision: int = 100, depth: int = 8) -> abi.hk_game_params:
    return abi.hk_game_params(velocityBucketSize=bucket, timePrecision=time_precision, sectionWindow=2,
                              treeSearchDepth=depth, maxLaneChanges=track.max_lane_changes, collisionWindow=0.1,
                              teamScoreRewardMultiplier=0.75, maxEpisodeSteps=track.max_episode_steps)


def root_state(track: Track, initial_section: int, lanes, teams=None, buckets=None, tire_age=2500, lane_changes=0,
               times=None, depth: int = 8) -> abi.hk_game_state:
    """Root state as `HierarchicalKartAgent.planWithMCTS` builds it (HKA:194-245): every kart at `initialSection`,
    `player` always 0 (quirk B.6-2), velocity bucket (0, bucket) unless `buckets` is given (quirk B.6-1)."""
    n = len(lanes)
    st = abi.hk_game_state()
    st.n_karts = n
    st.initialSection = initial_section
    st.lastCompletedSection = initial_section
    st.finalSection = initial_section + depth
    for i in range(n):
        mn, mx = buckets[i] if buckets is not None else (0, 2)
        st.karts[i] = abi.hk_kart_state(player=0, team=(teams[i] if teams is not None else i), section=initial_section,
                                        timeAtSection=(times[i] if times is not None else 0), min_velocity=mn,
                                        max_velocity=mx, lane=lanes[i], tireAge=tire_age, laneChanges=lane_changes,
                                        infeasible=0)
    return st
