"""Host-side mirror of the reference's `KartGame.AI.LQR` API over the C-ABI (Python flavour, used by tests/bench;
the C++ flavour is host/KartLQR.hpp and the C# shim a maintainer would drop in is csharp/KartLQR.cs).

Same names, argument meaning and error behaviour as
  KartLQR.solveFeedbackLQR          Assets/Karting/Scripts/AI/LQR/KartLQR.cs:17
  KartLQRDynamics/LinearizedBicycle Assets/Karting/Scripts/AI/LQR/KartLQRDynamics.cs:14-73
  KartLQRCosts/LQRCheckpointReachAvoidCost  Assets/Karting/Scripts/AI/LQR/KartLQRCosts.cs:13-141
The provider classes only *describe* a problem (they fill plain arrays); every solve runs in the CUDA library.
"""
from __future__ import annotations

import math

import numpy as np

from . import abi

xIndex, zIndex, vIndex, hIndex = 0, 1, 2, 3          # MPC/KartMPC.cs:15-18


class KartLQRDynamics:                                 # KartLQRDynamics.cs:14-20
    def getA(self) -> np.ndarray: raise NotImplementedError
    def getB(self) -> np.ndarray: raise NotImplementedError
    def getXDim(self) -> int: raise NotImplementedError
    def getUDim(self) -> int: raise NotImplementedError


class LinearizedBicycle(KartLQRDynamics):              # KartLQRDynamics.cs:25-73
    xDim, uDim = 4, 2

    def __init__(self, dt: float, initial):
        self.dt = float(dt)
        self.initial = np.array(initial, dtype=np.float64)
        self._A = self._B = None

    def getA(self):
        if self._A is None:
            x, dt = self.initial, self.dt
            A = np.eye(4)
            A[xIndex, vIndex] = math.cos(x[hIndex]) * dt
            A[zIndex, vIndex] = math.sin(x[hIndex]) * dt
            A[xIndex, hIndex] = -math.sin(x[hIndex]) * dt * x[vIndex]
            A[zIndex, hIndex] = math.cos(x[hIndex]) * dt * x[vIndex]
            self._A = A
        return self._A

    def getB(self):
        if self._B is None:
            B = np.zeros((4, 2))
            B[vIndex, 0] = self.dt
            B[hIndex, 1] = self.dt
            self._B = B
        return self._B

    def getXDim(self): return self.xDim
    def getUDim(self): return self.uDim


class KartLQRCosts:                                    # KartLQRCosts.cs:13-19
    def getQVec(self) -> np.ndarray: raise NotImplementedError
    def getQMatrix(self) -> np.ndarray: raise NotImplementedError
    def getRMatrix(self) -> np.ndarray: raise NotImplementedError


class LQRCheckpointReachAvoidCost(KartLQRCosts):       # KartLQRCosts.cs:25-141
    def __init__(self, targetState, targetWeights, controlWeight, currentDynamics, opponentTargetStates,
                 opponentTargetWeights, avoidWeights, avoidIndices, avoidDynamics):
        self.targetState = np.array(targetState, dtype=np.float64)
        self.targetWeights = dict(targetWeights)
        self.controlWeight = float(controlWeight)
        self.currentDynamics = currentDynamics
        self.opponentTargetStates = [np.array(t, dtype=np.float64) for t in opponentTargetStates]
        self.opponentTargetWeights = [dict(w) for w in opponentTargetWeights]
        self.avoidWeights = {k: list(v) for k, v in avoidWeights.items()}
        self.avoidIndices = {k: list(v) for k, v in avoidIndices.items()}
        self.avoidDynamics = list(avoidDynamics)
        self._Q = self._q = self._R = None

    def _total(self):
        return self.currentDynamics.getXDim() + sum(d.getXDim() for d in self.avoidDynamics)

    def getQMatrix(self):                              # :57-98
        if self._Q is None:
            n = self._total()
            Q = np.zeros((n, n))
            for s in self.avoidWeights:                # dictionary insertion order = x, z (HKA:964-969)
                curr = self.currentDynamics.getXDim()
                total = 0.0
                for i, d in enumerate(self.avoidDynamics):
                    t = curr + self.avoidIndices[s][i]
                    w = self.avoidWeights[s][i]
                    Q[s, t] = w
                    Q[t, s] = w
                    Q[t, t] = -w
                    total -= w
                    curr += d.getXDim()
                Q[s, s] = total
            for s, w in self.targetWeights.items():
                Q[s, s] += w
            curr = self.currentDynamics.getXDim()
            for i, tw in enumerate(self.opponentTargetWeights):
                for o, w in tw.items():
                    Q[curr + o, curr + o] = -w
                curr += self.avoidDynamics[i].getXDim()
            self._Q = Q
        return self._Q

    def getQVec(self):                                 # :103-127
        if self._q is None:
            n = self._total()
            q = np.zeros(n)
            xd = self.currentDynamics.getXDim()
            q[:xd] = -self.targetState
            for s, w in self.targetWeights.items():
                q[s] = q[s] * w
            curr = xd
            for i, t in enumerate(self.opponentTargetStates):
                d = self.avoidDynamics[i].getXDim()
                q[curr:curr + d] = t
                for o, w in self.opponentTargetWeights[i].items():
                    q[curr + o] = q[curr + o] * -w
                curr += d
            self._q = q
        return self._q

    def getRMatrix(self):                              # :132-140
        if self._R is None:
            self._R = np.eye(self.currentDynamics.getUDim()) * self.controlWeight
        return self._R


def flatten(dynamics, costs, initials):
    """What the shim does before crossing the ABI: stack per-player blocks into the hk_abi.h record layout.
    Raises ValueError where MathNet would raise ArgumentException (dimension mismatch)."""
    N = len(dynamics)
    if N < 1 or N > abi.HK_MAX_PLAYERS or len(costs) != N or len(initials) != N:
        raise ValueError("player count mismatch")
    for d in dynamics:
        if d.getXDim() != 4 or d.getUDim() != 2:
            raise ValueError("only 4-state / 2-control players are supported (LinearizedBicycle.xDim/uDim)")
    n = 4 * N
    A = np.stack([np.asarray(d.getA(), dtype=np.float64) for d in dynamics])
    B = np.stack([np.asarray(d.getB(), dtype=np.float64) for d in dynamics])
    Q = np.stack([np.asarray(c.getQMatrix(), dtype=np.float64) for c in costs])
    q = np.stack([np.asarray(c.getQVec(), dtype=np.float64) for c in costs])
    R = np.stack([np.asarray(c.getRMatrix(), dtype=np.float64) for c in costs])
    x0 = np.concatenate([np.asarray(v, dtype=np.float64) for v in initials])
    if A.shape != (N, 4, 4) or B.shape != (N, 4, 2) or Q.shape != (N, n, n) or q.shape != (N, n) or R.shape != (N, 2, 2) \
            or x0.shape != (n,):
        raise ValueError("dimension mismatch")
    return tuple(np.ascontiguousarray(v) for v in (A, B, Q, q, R, x0))


class KartLQR:
    @staticmethod
    def solveFeedbackLQR(dynamics, costs, initials, horizon: int) -> np.ndarray:
        """Drop-in for KartLQR.cs:17 — returns player 0's first control (length 2), computed on the GPU."""
        A, B, Q, q, R, x0 = flatten(dynamics, costs, initials)
        N = len(dynamics)
        u0 = np.empty(2 * N)
        lib = abi.load_library()
        abi.check(lib.hk_lqng_solve_one(N, int(horizon), abi.dptr(A), abi.dptr(B), abi.dptr(Q), abi.dptr(q), abi.dptr(R),
                                        abi.dptr(x0), abi.dptr(u0)))
        return u0[:2].copy()


def solve_batch(A, B, Q, q, R, x0, horizon: int, time_varying: bool = False, full: bool = True):
    """hk_lqng_solve_batch with host arrays in the record layout; returns dict(u0, P, alpha, traj, status)."""
    A, B, Q, q, R, x0 = (np.ascontiguousarray(v, dtype=np.float64) for v in (A, B, Q, q, R, x0))
    if x0.ndim != 2 or x0.shape[1] % 4:
        raise ValueError("x0 must be [batch][4N]")
    batch, n = x0.shape
    N, m, T = n // 4, n // 2, horizon + 1
    tv = (T,) if time_varying else ()
    for name, arr, shp in (("A", A, (N, 4, 4)), ("B", B, (N, 4, 2)), ("Q", Q, (N, n, n)), ("q", q, (N, n)), ("R", R, (N, 2, 2))):
        if arr.shape != (batch,) + tv + shp:
            raise ValueError(f"{name} has shape {arr.shape}, expected {(batch,) + tv + shp}")
    u0 = np.empty((batch, m))
    P = np.empty((batch, T, m, n)) if full else None
    alpha = np.empty((batch, T, m)) if full else None
    traj = np.empty((batch, T + 1, n)) if full else None
    status = np.zeros(batch, dtype=np.int32)
    lib = abi.load_library()
    abi.check(lib.hk_lqng_solve_batch(batch, N, int(horizon), int(time_varying), abi.dptr(A), abi.dptr(B), abi.dptr(Q),
                                      abi.dptr(q), abi.dptr(R), abi.dptr(x0), abi.dptr(u0), abi.dptr(P), abi.dptr(alpha),
                                      abi.dptr(traj), abi.iptr(status)))
    return dict(u0=u0, P=P, alpha=alpha, traj=traj, status=status)


def assemble_solve_batch(prob: dict, horizon: int):
    """hk_lqng_assemble_solve_batch on a compact problem dict (see scenarios.py)."""
    x0 = np.ascontiguousarray(prob["x0"], dtype=np.float64)
    batch, N = x0.shape[0], x0.shape[1]
    arrs = [np.ascontiguousarray(prob[k], dtype=np.float64) for k in ("target", "tw", "cw", "aw", "otgt", "otw")]
    u0 = np.empty((batch, 2 * N))
    status = np.zeros(batch, dtype=np.int32)
    lib = abi.load_library()
    abi.check(lib.hk_lqng_assemble_solve_batch(batch, N, int(horizon), float(prob["dt"]), abi.dptr(x0),
                                               *[abi.dptr(a) for a in arrs], abi.dptr(u0), abi.iptr(status)))
    return dict(u0=u0, status=status)


def pack_records(prob: dict) -> np.ndarray:
    """The packed form of a compact problem dict: [batch][13 N + 9 N (N-1)] doubles, fields in the order x0 | target | tw | cw | aw |
    otgt | otw (hk_lqng_assemble_solve_packed)."""
    x0 = np.asarray(prob["x0"], dtype=np.float64)
    batch = x0.shape[0]
    parts = [np.asarray(prob[k], dtype=np.float64).reshape(batch, -1) for k in ("x0", "target", "tw", "cw", "aw", "otgt", "otw")]
    return np.ascontiguousarray(np.concatenate(parts, axis=1))


def assemble_solve_packed(records: np.ndarray, n_players: int, horizon: int, dt: float, u0: np.ndarray | None = None,
                          status: np.ndarray | None = None):
    """hk_lqng_assemble_solve_packed on packed records (pack_records)."""
    records = np.ascontiguousarray(records, dtype=np.float64)
    batch = records.shape[0]
    u0 = np.empty((batch, 2 * n_players)) if u0 is None else u0
    status = np.zeros(batch, dtype=np.int32) if status is None else status
    abi.check(abi.load_library().hk_lqng_assemble_solve_packed(batch, n_players, int(horizon), float(dt), abi.dptr(records), abi.dptr(u0),
                                                               abi.iptr(status)))
    return dict(u0=u0, status=status)
