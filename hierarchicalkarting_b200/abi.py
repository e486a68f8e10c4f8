"""ctypes mirror of include/hk_abi.h and loader of libhk_b200.so (the CUDA product library).

There is no CPU fallback: if the library was not built, loading raises, and if no CUDA device is present every
compute entry returns HK_ERR_NO_DEVICE, which `check()` turns into an exception.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HK_ABI_VERSION = 1
HK_XDIM, HK_UDIM = 4, 2
HK_MAX_PLAYERS = 4
HK_MAX_HORIZON = 31
HK_MAX_KARTS = 4
HK_MAX_ACTIONS = 36
HK_MCTS_MAX_SEQ = 16
HK_MAX_PLIES = 64
HK_MAX_SECTIONS = 64
HK_MAX_LAPS = 8

HK_OK = 0
HK_ERR_INVALID_ARGUMENT = -1
HK_ERR_NO_DEVICE = -2
HK_ERR_CUDA = -3
HK_ERR_OUT_OF_MEMORY = -4
HK_ERR_NO_UPNEXT = -5


class hk_section(C.Structure):
    _fields_ = [("insideR", C.c_float), ("length", C.c_float), ("width", C.c_float), ("turnDeg", C.c_float),
                ("leftTurn", C.c_int32), ("optimalLane", C.c_int32)]


class hk_kart(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("accel", "braking", "topSpeed", "reverseSpeed", "maxGs", "minGs", "tireWearFactor")]


class hk_game_params(C.Structure):
    _fields_ = [("velocityBucketSize", C.c_int32), ("timePrecision", C.c_int32), ("sectionWindow", C.c_int32),
                ("treeSearchDepth", C.c_int32), ("maxLaneChanges", C.c_int32), ("collisionWindow", C.c_float),
                ("teamScoreRewardMultiplier", C.c_float), ("maxEpisodeSteps", C.c_int32)]


KART_STATE_FIELDS = ("player", "team", "section", "timeAtSection", "min_velocity", "max_velocity", "lane", "tireAge",
                     "laneChanges", "infeasible")


class hk_kart_state(C.Structure):
    _fields_ = [(k, C.c_int32) for k in KART_STATE_FIELDS]

    def astuple(self):
        return tuple(getattr(self, k) for k in KART_STATE_FIELDS)


class hk_action(C.Structure):
    _fields_ = [("min_velocity", C.c_int32), ("max_velocity", C.c_int32), ("lane", C.c_int32)]

    def astuple(self):
        return (self.min_velocity, self.max_velocity, self.lane)


class hk_game_state(C.Structure):
    _fields_ = [("n_karts", C.c_int32), ("initialSection", C.c_int32), ("lastCompletedSection", C.c_int32),
                ("finalSection", C.c_int32), ("karts", hk_kart_state * HK_MAX_KARTS)]

    def astuple(self):
        return (self.n_karts, self.initialSection, self.lastCompletedSection, self.finalSection,
                tuple(self.karts[i].astuple() for i in range(self.n_karts)))


class hk_race_kart(C.Structure):
    _fields_ = [("x", C.c_double), ("z", C.c_double), ("v", C.c_double), ("h", C.c_double), ("steer", C.c_float),
                ("section", C.c_int32), ("lane", C.c_int32), ("laneChanges", C.c_int32), ("illegalLaneChanges", C.c_int32),
                ("sectionStep", C.c_int32), ("active", C.c_int32), ("team", C.c_int32)]


class hk_race_plan(C.Structure):
    _fields_ = [("lane", C.c_int8 * HK_MAX_SECTIONS), ("vel", C.c_float * HK_MAX_SECTIONS),
                ("oppLane", C.c_int8 * HK_MAX_SECTIONS), ("oppVel", C.c_float * HK_MAX_SECTIONS),
                ("sectionTimes", C.c_int32 * HK_MAX_SECTIONS), ("lapStep", C.c_int32 * HK_MAX_LAPS),
                ("avgLaneDiff", C.c_float), ("avgVelDiff", C.c_float)]


class hk_race_belief(C.Structure):
    _fields_ = [("lane", C.c_int8 * HK_MAX_SECTIONS), ("vel", C.c_float * HK_MAX_SECTIONS)]


class hk_race_params(C.Structure):
    _fields_ = [("dt", C.c_double), ("accel", C.c_float), ("braking", C.c_float), ("coastingDrag", C.c_float),
                ("topSpeed", C.c_float), ("gateHalfWidth", C.c_float), ("maxLaneChanges", C.c_int32),
                ("goalSection", C.c_int32), ("highModeMcts", C.c_int32), ("velocityBucketSize", C.c_int32),
                ("treeSearchDepth", C.c_int32), ("planEvery", C.c_int32), ("horizon", C.c_int32)]


class hk_mcts_node(C.Structure):
    _fields_ = [("child_mask", C.c_uint64), ("totalValue", C.c_float), ("numEpisodes", C.c_int32), ("first_child", C.c_int32),
                ("last_child", C.c_int32), ("next_sibling", C.c_int32), ("gen", C.c_uint8), ("n_legal", C.c_uint8), ("upnext", C.c_int8),
                ("pad_", C.c_uint8)]


class hk_race_mcts_params(C.Structure):
    _fields_ = [("mode", C.c_int32), ("iterations", C.c_int32), ("first_iterations", C.c_int32), ("rollouts_per_leaf", C.c_int32),
                ("reuse_cycles", C.c_int32), ("apply_delay", C.c_int32), ("seed", C.c_uint64), ("max_tree_nodes", C.c_int32), ("pad_", C.c_int32)]


# numpy views of the same layouts (for batched buffers)
RACE_KART_DTYPE = np.dtype([("x", np.float64), ("z", np.float64), ("v", np.float64), ("h", np.float64), ("steer", np.float32),
                            ("section", np.int32), ("lane", np.int32), ("laneChanges", np.int32),
                            ("illegalLaneChanges", np.int32), ("sectionStep", np.int32), ("active", np.int32), ("team", np.int32)])
RACE_PLAN_DTYPE = np.dtype([("lane", np.int8, (HK_MAX_SECTIONS,)), ("vel", np.float32, (HK_MAX_SECTIONS,)),
                            ("oppLane", np.int8, (HK_MAX_SECTIONS,)), ("oppVel", np.float32, (HK_MAX_SECTIONS,)),
                            ("sectionTimes", np.int32, (HK_MAX_SECTIONS,)), ("lapStep", np.int32, (HK_MAX_LAPS,)),
                            ("avgLaneDiff", np.float32), ("avgVelDiff", np.float32)])
RACE_BELIEF_DTYPE = np.dtype([("lane", np.int8, (HK_MAX_SECTIONS,)), ("vel", np.float32, (HK_MAX_SECTIONS,))])
assert RACE_BELIEF_DTYPE.itemsize == C.sizeof(hk_race_belief) == 320
assert RACE_KART_DTYPE.itemsize == C.sizeof(hk_race_kart) == 64
assert RACE_PLAN_DTYPE.itemsize == C.sizeof(hk_race_plan) == 936
KART_STATE_DTYPE = np.dtype([(k, np.int32) for k in KART_STATE_FIELDS])
ACTION_DTYPE = np.dtype([("min_velocity", np.int32), ("max_velocity", np.int32), ("lane", np.int32)])
GAME_STATE_DTYPE = np.dtype([("n_karts", np.int32), ("initialSection", np.int32), ("lastCompletedSection", np.int32),
                             ("finalSection", np.int32), ("karts", KART_STATE_DTYPE, (HK_MAX_KARTS,))])
assert GAME_STATE_DTYPE.itemsize == C.sizeof(hk_game_state)
MCTS_NODE_DTYPE = np.dtype([("child_mask", np.uint64), ("totalValue", np.float32), ("numEpisodes", np.int32), ("first_child", np.int32),
                            ("last_child", np.int32), ("next_sibling", np.int32), ("gen", np.uint8), ("n_legal", np.uint8),
                            ("upnext", np.int8), ("pad_", np.uint8)])
assert MCTS_NODE_DTYPE.itemsize == 32
assert ACTION_DTYPE.itemsize == C.sizeof(hk_action)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_fp = C.POINTER(C.c_float)
_up = C.POINTER(C.c_uint32)

# name -> (restype, argtypes); this is also the export list the CPU test checks against include/hk_abi.h
PROTOTYPES = {
    "hk_abi_version": (C.c_int, []),
    "hk_init": (C.c_int, [C.c_int]),
    "hk_shutdown": (None, []),
    "hk_last_error": (C.c_char_p, []),
    "hk_device_count": (C.c_int, []),
    "hk_kernel_launch_count": (C.c_longlong, []),
    "hk_probe_fp64_peak": (C.c_int, [C.c_double, _dp, _dp, _dp]),
    "hk_lqng_solve_batch": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip]),
    "hk_lqng_solve_batch_device": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 11 + [C.c_void_p]),
    "hk_lqng_solve_one": (C.c_int, [C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "hk_lqng_assemble_solve_batch": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip]),
    "hk_lqng_assemble_solve_packed": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, _dp, _dp, _ip]),
    "hk_game_create": (C.c_int, [C.POINTER(hk_section), C.c_int, C.POINTER(hk_kart), C.c_int, C.POINTER(hk_kart), C.c_int,
                                 C.POINTER(hk_game_params), C.POINTER(C.c_void_p)]),
    "hk_game_destroy": (None, [C.c_void_p]),
    "hk_game_replay_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 10),
    "hk_mcts_rollouts": (C.c_int, [C.c_void_p, C.POINTER(hk_game_state), C.c_int64, C.c_uint64, C.c_uint64, _lp, _dp, _lp, _lp]),
    "hk_mcts_rollouts_trace": (C.c_int, [C.c_void_p, C.POINTER(hk_game_state), C.c_int64, C.c_uint64, C.c_uint64] + [C.c_void_p] * 5),
    "hk_mcts_rollouts_multi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_uint64] + [C.c_void_p] * 4),
    "hk_mcts_search_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64] + [C.c_void_p] * 5),
    "hk_mcts_forest_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "hk_mcts_forest_destroy": (None, [C.c_void_p]),
    "hk_mcts_forest_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64] + [C.c_void_p] * 4),
    "hk_mcts_forest_nodes": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, _ip]),
    "hk_mcts_search_seq_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64] + [C.c_void_p] * 6),
    "hk_policy_cdf": (C.c_int, [C.c_int, _up]),
    "hk_track_create": (C.c_int, [C.POINTER(hk_section), _dp, _dp, _dp, C.c_int, C.POINTER(C.c_void_p)]),
    "hk_track_destroy": (None, [C.c_void_p]),
    "hk_race_recipe": (C.c_int, [C.c_void_p, C.POINTER(hk_race_params), C.c_int, C.c_void_p, C.c_void_p] + [_dp] * 7),
    "hk_race_step": (C.c_int, [C.c_void_p, C.POINTER(hk_race_params), C.c_int, C.c_int, _dp, C.c_void_p, C.c_void_p]),
    "hk_race_plan_fixed": (C.c_int, [C.c_void_p, C.POINTER(hk_race_params), C.c_int, C.c_void_p, C.c_void_p]),
    "hk_race_run": (C.c_int, [C.c_void_p, C.POINTER(hk_race_params), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, _dp, _lp]),
    "hk_race_planner_create": (C.c_int, [C.c_void_p, C.POINTER(hk_race_mcts_params), C.c_int, C.POINTER(C.c_void_p)]),
    "hk_race_planner_destroy": (None, [C.c_void_p]),
    "hk_race_planner_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hk_race_run_planned": (C.c_int, [C.c_void_p, C.POINTER(hk_race_params), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, _dp, _lp]),
    "hk_raceN_recipe": (C.c_int, [C.c_void_p, C.POINTER(hk_race_params), C.c_int, C.c_int] + [C.c_void_p] * 12),
    "hk_raceN_planner_create": (C.c_int, [C.c_void_p, C.POINTER(hk_race_mcts_params), C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "hk_raceN_run": (C.c_int, [C.c_void_p, C.POINTER(hk_race_params), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, _lp]),
    "hk_race_run_device": (C.c_int, [C.c_void_p, C.POINTER(hk_race_params), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, _lp,
                                     C.c_void_p]),
    "hk_raceN_run_device": (C.c_int, [C.c_void_p, C.POINTER(hk_race_params), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, _lp, C.c_void_p]),
    "hk_race_run_mcts": (C.c_int, [C.c_void_p, C.POINTER(hk_race_params), C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, _dp, _lp]),
}

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libhk_b200.so")
_lib = None


class HKError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"hk status {status}: {message}")
        self.status = status


def load_library(path: str | None = None):
    """Load libhk_b200.so and attach prototypes. Raises if the CUDA library has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError(
            f"{p} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()'). "
            "There is no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    if lib.hk_abi_version() != HK_ABI_VERSION:
        raise RuntimeError("hk ABI version mismatch")
    if path is None:
        _lib = lib
    return lib


def check(status: int):
    if status != HK_OK:
        msg = load_library().hk_last_error()
        raise HKError(status, msg.decode() if msg else "")


def dptr(a: np.ndarray | None):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def iptr(a: np.ndarray | None):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_ip)


def vptr(a: np.ndarray | None):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)
