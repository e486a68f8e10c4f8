"""The oracle's OWN ctypes / numpy definitions of the plain-data structs of include/hk_abi.h (TEST INFRASTRUCTURE).

Written from the header, not imported from the product package: a layout bug in hierarchicalkarting_b200/abi.py must not be
common-mode between the product and its checker.  tests/test_abi_cpu.py parses include/hk_abi.h and checks BOTH sets of
definitions against it (field names, order, C types, sizes)."""
import ctypes as C

import numpy as np

HK_MAX_KARTS = 4
HK_MAX_ACTIONS = 36
HK_MAX_PLIES = 64
HK_MCTS_MAX_SEQ = 16
HK_MAX_SECTIONS = 64
HK_MAX_LAPS = 8
HK_MAX_PLAYERS = 4

i32, f32, f64, i8 = C.c_int32, C.c_float, C.c_double, C.c_int8


class _S(C.Structure):
    def astuple(self):
        out = []
        for name, typ, *_ in self._fields_:
            v = getattr(self, name)
            out.append(tuple(x.astuple() if isinstance(x, _S) else x for x in v) if isinstance(v, C.Array) else v)
        return tuple(out)


class hk_section(_S):
    _fields_ = [("insideR", f32), ("length", f32), ("width", f32), ("turnDeg", f32), ("leftTurn", i32), ("optimalLane", i32)]


class hk_kart(_S):
    _fields_ = [("accel", f32), ("braking", f32), ("topSpeed", f32), ("reverseSpeed", f32), ("maxGs", f32), ("minGs", f32),
                ("tireWearFactor", f32)]


class hk_game_params(_S):
    _fields_ = [("velocityBucketSize", i32), ("timePrecision", i32), ("sectionWindow", i32), ("treeSearchDepth", i32),
                ("maxLaneChanges", i32), ("collisionWindow", f32), ("teamScoreRewardMultiplier", f32), ("maxEpisodeSteps", i32)]


class hk_kart_state(_S):
    _fields_ = [("player", i32), ("team", i32), ("section", i32), ("timeAtSection", i32), ("min_velocity", i32),
                ("max_velocity", i32), ("lane", i32), ("tireAge", i32), ("laneChanges", i32), ("infeasible", i32)]


class hk_action(_S):
    _fields_ = [("min_velocity", i32), ("max_velocity", i32), ("lane", i32)]


class hk_game_state(_S):
    _fields_ = [("n_karts", i32), ("initialSection", i32), ("lastCompletedSection", i32), ("finalSection", i32),
                ("karts", hk_kart_state * HK_MAX_KARTS)]

    def astuple(self):
        return (self.n_karts, self.initialSection, self.lastCompletedSection, self.finalSection,
                tuple(self.karts[i].astuple() for i in range(self.n_karts)))


class hk_race_kart(_S):
    _fields_ = [("x", f64), ("z", f64), ("v", f64), ("h", f64), ("steer", f32), ("section", i32), ("lane", i32),
                ("laneChanges", i32), ("illegalLaneChanges", i32), ("sectionStep", i32), ("active", i32), ("team", i32)]


class hk_race_plan(_S):
    _fields_ = [("lane", i8 * HK_MAX_SECTIONS), ("vel", f32 * HK_MAX_SECTIONS), ("oppLane", i8 * HK_MAX_SECTIONS),
                ("oppVel", f32 * HK_MAX_SECTIONS), ("sectionTimes", i32 * HK_MAX_SECTIONS), ("lapStep", i32 * HK_MAX_LAPS),
                ("avgLaneDiff", f32), ("avgVelDiff", f32)]


class hk_race_belief(_S):
    _fields_ = [("lane", i8 * HK_MAX_SECTIONS), ("vel", f32 * HK_MAX_SECTIONS)]


class hk_race_params(_S):
    _fields_ = [("dt", f64), ("accel", f32), ("braking", f32), ("coastingDrag", f32), ("topSpeed", f32), ("gateHalfWidth", f32),
                ("maxLaneChanges", i32), ("goalSection", i32), ("highModeMcts", i32), ("velocityBucketSize", i32),
                ("treeSearchDepth", i32), ("planEvery", i32), ("horizon", i32)]


class hk_race_mcts_params(_S):
    _fields_ = [("mode", i32), ("iterations", i32), ("first_iterations", i32), ("rollouts_per_leaf", i32), ("reuse_cycles", i32),
                ("apply_delay", i32), ("seed", C.c_uint64), ("max_tree_nodes", i32), ("pad_", i32)]


class hk_mcts_node(_S):
    _fields_ = [("child_mask", C.c_uint64), ("totalValue", f32), ("numEpisodes", i32), ("first_child", i32), ("last_child", i32),
                ("next_sibling", i32), ("gen", C.c_uint8), ("n_legal", C.c_uint8), ("upnext", i8), ("pad_", C.c_uint8)]


STRUCTS = {c.__name__: c for c in (hk_section, hk_kart, hk_game_params, hk_kart_state, hk_action, hk_game_state, hk_race_kart,
                                   hk_race_plan, hk_race_belief, hk_race_params, hk_race_mcts_params, hk_mcts_node)}

_NP = {i32: np.int32, f32: np.float32, f64: np.float64, i8: np.int8, C.c_uint64: np.uint64, C.c_uint8: np.uint8}


def np_dtype(cls) -> np.dtype:
    """numpy record dtype with the layout of a ctypes struct above (nested structs and arrays included)."""
    fields = []
    for name, typ in cls._fields_:
        if issubclass(typ, C.Array):
            base = typ._type_
            fields.append((name, np_dtype(base) if issubclass(base, C.Structure) else _NP[base], (typ._length_,)))
        elif issubclass(typ, C.Structure):
            fields.append((name, np_dtype(typ)))
        else:
            fields.append((name, _NP[typ]))
    dt = np.dtype(fields, align=True)
    assert dt.itemsize == C.sizeof(cls), (cls.__name__, dt.itemsize, C.sizeof(cls))
    return dt


GAME_STATE_DTYPE = np_dtype(hk_game_state)
RACE_KART_DTYPE = np_dtype(hk_race_kart)
RACE_PLAN_DTYPE = np_dtype(hk_race_plan)


def ref(x):
    """void* of a ctypes struct / array or a numpy array — whichever package's class it is an instance of."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data)
    if isinstance(x, (C.Structure, C.Array)):
        return C.cast(C.byref(x), C.c_void_p)
    return x


def game_state(x) -> hk_game_state:
    """Copy of any object with hk_game_state's bytes (product ctypes struct, numpy record, this module's struct)."""
    if isinstance(x, np.void) or isinstance(x, np.ndarray):
        return hk_game_state.from_buffer_copy(x.tobytes())
    return hk_game_state.from_buffer_copy(bytes(x))


def action(a) -> hk_action:
    if isinstance(a, hk_action):
        return a
    if isinstance(a, C.Structure):
        return hk_action(a.min_velocity, a.max_velocity, a.lane)
    return hk_action(*[int(v) for v in a])
