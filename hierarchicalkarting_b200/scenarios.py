"""Seeded synthetic LQNG problems along the Oval / Complex tracks and their dense assembly.

`make_problems` follows the problem-construction recipe of `HierarchicalKartAgent.SolveLQR`
(Assets/Karting/Scripts/AI/HierarchicalKartAgent.cs:699-1197) minus its Unity raycast branches: the initial states
(:730-736), targets (:745-817), the raycast-free target-heading rules (:821-823, :891-902, :919-923), the MCTS-mode
weights (:930-962), avoid weights computed in float32 (:985-1023), opponent / teammate target weights (:1071-1180) and
control weights (:1192-1196).  The output is the compact per-player description that `hk_lqng_assemble_solve_batch`
takes; `assemble_dense` expands it to the A,B,Q,q,R records of `hk_lqng_solve_batch` exactly as
`LinearizedBicycle` (KartLQRDynamics.cs:40-62) and `LQRCheckpointReachAvoidCost` (KartLQRCosts.cs:57-140) would.
"""
from __future__ import annotations

import numpy as np

from .tracks import COMPLEX, OVAL, Track

DT = float(np.float32(0.02))            # (double)Time.fixedDeltaTime, HKA:707; TimeManager.asset:5
F32 = np.float32


def _mathf_atan2(y, x):
    """Mathf.Atan2(float, float) = (float)Math.Atan2((double)y, (double)x)."""
    return np.arctan2(np.asarray(y, dtype=np.float32).astype(np.float64), np.asarray(x, dtype=np.float32).astype(np.float64)).astype(np.float32)


def _mathf_pow(x, p):
    """Mathf.Pow(float, float) = (float)Math.Pow((double)x, (double)p)."""
    return np.power(np.asarray(x, dtype=np.float32).astype(np.float64), float(F32(p))).astype(np.float32)


def _wrap2pi(a):
    return np.where(a < 0, a + F32(2) * F32(np.pi), a)


def _angle_difference(a1, a2):          # HKA:1341-1344 (double)
    return np.arctan2(np.sin(a2 - a1), np.cos(a2 - a1))


def private_order(N: int, teams) -> np.ndarray:
    """order[i] = joint indices [self, its otherAgents..., its teamAgents...] of player i (KartLQRCosts.cs:62-94 fed by
    HKA:1004-1190).  Joint order is [this, this.teamAgents..., this.otherAgents...] (HKA:702)."""
    out = np.zeros((N, N), dtype=np.int64)
    for i in range(N):
        others = [j for j in range(N) if teams[j] != teams[i]]
        mates = [j for j in range(N) if teams[j] == teams[i] and j != i]
        out[i] = [i] + others + mates
    return out


def make_problems(track: Track, batch: int, n_players: int, seed: int, slow_fraction: float = 0.05, high_mode_mcts: bool = True):
    """Compact problem batch: dict(x0[b,N,4], target[b,N,4], tw[b,N,4], cw[b,N], aw[b,N,N-1,2], otgt[b,N,N-1,4],
    otw[b,N,N-1,3], dt, teams, order)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    N, L = n_players, track.n_sections
    lanes_xy, trig, head = track.lane_table(), track.trigger_table(), track.heading_table()
    straight = track.straight_table()
    teams = [0, 1] if N == 2 else ([0] if N == 1 else ([0, 0, 1] if N == 3 else [0, 0, 1, 1]))
    order = private_order(N, teams)

    sec = rng.integers(0, L, size=batch)                                   # ego section
    frac = rng.random((batch, N)) if N == 2 else np.repeat(rng.random((batch, 1)), N, axis=1)
    lane = rng.integers(1, 5, size=(batch, N))
    lat = rng.uniform(-0.4, 0.4, size=(batch, N))
    gap = rng.uniform(-6.0, 6.0, size=(batch, N)) if N == 2 else rng.uniform(-3.0, 3.0, size=(batch, N))
    gap[:, 0] = 0.0
    v = rng.uniform(5.5, 15.0, size=(batch, N))
    slow = rng.random((batch, N)) < slow_fraction
    v = np.where(slow, rng.uniform(0.0, 5.0, size=(batch, N)), v)
    hnoise = rng.normal(0.0, 0.08, size=(batch, N))
    tgt_lane = rng.integers(1, 5, size=(batch, N))
    nxt_lane = rng.integers(1, 5, size=(batch, N))
    bucket_max = rng.choice(np.array([8, 10, 12, 14, 15]), size=(batch, N))
    nbucket_max = rng.choice(np.array([8, 10, 12, 14, 15]), size=(batch, N))

    # positions: along the lane line from checkpoint `sec` to `sec+1`, plus along-track gap and lateral noise
    s0 = sec[:, None] % L
    s1 = (sec[:, None] + 1) % L
    s2 = (sec[:, None] + 2) % L
    p0 = lanes_xy[s0, lane - 1]                                            # [b,N,2]
    p1 = lanes_xy[s1, lane - 1]
    seg = p1 - p0
    seglen = np.linalg.norm(seg, axis=-1, keepdims=True)
    tdir = seg / np.maximum(seglen, 1e-9)
    ndir = np.stack([-tdir[..., 1], tdir[..., 0]], axis=-1)
    pos = p0 + seg * frac[..., None] + tdir * gap[..., None] + ndir * lat[..., None]
    pos = pos.astype(np.float32).astype(np.float64)                        # Unity transform.position is float32
    v = v.astype(np.float32).astype(np.float64)
    seg_heading = np.arctan2(tdir[..., 1], tdir[..., 0])
    heading = _wrap2pi((seg_heading + hnoise).astype(np.float32))          # Mathf.Atan2 + wrap, HKA:734-736
    heading = np.where(heading >= F32(2 * np.pi), heading - F32(2 * np.pi), heading).astype(np.float64)
    x0 = np.concatenate([pos, v[..., None], heading[..., None]], axis=-1)  # [b,N,4]

    # targets (HKA:745-817): lane collider of section+1 / section+2, MCTS-mode velocity = min(max, bucket_max + 2*bucket)
    tl = lanes_xy[s1, tgt_lane - 1]
    nl = lanes_xy[s2, nxt_lane - 1]
    vel = np.minimum(15.0, bucket_max + (4.0 if high_mode_mcts else 0.0))
    nvel = np.minimum(15.0, nbucket_max + (4.0 if high_mode_mcts else 0.0))
    stopped = v <= 5.0
    tx, tz = tl[..., 0].copy(), tl[..., 1].copy()
    tv = np.where(stopped, 0.0, vel)
    d_t = np.linalg.norm((tl - pos).astype(np.float32), axis=-1)
    near = d_t <= np.where(straight[s0], F32(10.5), F32(7.5))             # HKA:823
    d_c = np.linalg.norm((trig[s1] - pos).astype(np.float32), axis=-1)     # stand-in for centerLine.ClosestPoint distance
    follow = near & (d_c <= 4.0)                                           # HKA:877-890: target the following checkpoint
    f1 = _mathf_atan2(tl[..., 1] - pos[..., 1], tl[..., 0] - pos[..., 0])
    f2 = _mathf_atan2(nl[..., 1] - tl[..., 1], nl[..., 0] - tl[..., 0])
    f6 = _mathf_atan2(nl[..., 1] - pos[..., 1], nl[..., 0] - pos[..., 0])
    f1w, f2w, f6w = _wrap2pi(f1).astype(np.float64), _wrap2pi(f2).astype(np.float64), _wrap2pi(f6).astype(np.float64)
    h0 = heading
    blend = f1w - _angle_difference(f2w, f1w) * float(F32(0.4))            # HKA:896
    blend = np.where(blend < 0, blend + 2 * float(F32(np.pi)), blend)
    th_far = h0 - _angle_difference(h0, f1w)                               # HKA:921
    th_blend = h0 - _angle_difference(h0, blend)                           # HKA:898
    th_follow = h0 - _angle_difference(h0, f6w)                            # HKA:887
    th = np.where(follow, th_follow, np.where(near, th_blend, th_far))
    tx = np.where(follow, nl[..., 0], tx)
    tz = np.where(follow, nl[..., 1], tz)
    tv = np.where(follow & ~stopped, nvel, tv)
    target = np.stack([tx, tz, tv, th], axis=-1)

    # own target weights (HKA:930-962)
    nb = max(N - 1, 1) if N > 2 else 1                                     # nearbyAgents (all within 8 m by construction)
    vmax1 = np.maximum(1.0, v)
    if N > 2:
        w_h = np.full((batch, N), (3.5 if high_mode_mcts else 2.5) * nb)   # HKA:932
    else:
        w_h = np.full((batch, N), 3.5 if high_mode_mcts else 1.9)
    w_xz = np.where(stopped, nb * 0.3 * 3.1, nb * 0.3 * 3.1 / vmax1)
    w_v = np.where(stopped, float(nb * -2), nb * 5e-4)
    tw = np.stack([w_xz, w_xz, w_v, w_h], axis=-1)
    cw = np.full((batch, N), (0.25 if high_mode_mcts else 0.135) if N > 2 else 0.115)   # HKA:1192-1196

    # avoid / opponent-target / teammate-target weights in each player's private ordering (HKA:964-1190)
    K = max(N - 1, 0)
    aw = np.zeros((batch, N, K, 2))
    otgt = np.zeros((batch, N, K, 4))
    otw = np.zeros((batch, N, K, 3))
    for i in range(N):
        if N > 2:
            mult = F32((1.0 if high_mode_mcts else 0.55) if i == 0 else 1.7) / F32(nb)   # HKA:985-987 (float / int)
        else:
            mult = F32((1.0 if high_mode_mcts else 0.45) if i == 0 else 1.3)             # HKA:999-1002
        n_opp_near = np.zeros(batch, dtype=np.int64)
        for k in range(K):
            o = order[i, 1 + k]
            is_mate = teams[o] == teams[i]
            dist32 = np.linalg.norm((pos[:, o] - pos[:, i]).astype(np.float32), axis=-1).astype(np.float32)
            far = dist32 > 8
            m_eff = (mult / F32(2.0)) if is_mate else mult                 # multiplier2, HKA:1113
            w32 = F32(1.0) / (_mathf_pow(dist32, 1.5) * m_eff)              # 1f/(Mathf.Pow(d,1.5f)*mult), HKA:1019
            w = np.where(far, 0.0, w32.astype(np.float64))
            aw[:, i, k, 0] = w
            aw[:, i, k, 1] = w
            # other's target: lane collider of ITS next section (same section by construction), heading entry 0
            otgt[:, i, k, 0] = tl[:, o, 0]
            otgt[:, i, k, 1] = tl[:, o, 1]
            if is_mate:
                otgt[:, i, k, 2] = 15.0                                    # getMaxSpeedForState() stand-in, HKA:1147-1160
                if N > 2:
                    wxz = -(3e-5 if high_mode_mcts else 0.0) / (vmax1[:, i] * nb)      # HKA:1178-1180
                else:
                    wxz = -(2e-4 if high_mode_mcts else 1e-4) / vmax1[:, i]
                zero = far | (n_opp_near < 1)                              # HKA:1168
                otw[:, i, k, 0] = np.where(zero, 0.0, wxz)
                otw[:, i, k, 1] = np.where(zero, 0.0, wxz)
                otw[:, i, k, 2] = 0.0
            else:
                otgt[:, i, k, 2] = vel[:, o]
                n_opp_near += (~far).astype(np.int64)
                if N > 2:
                    wxz = (0.2 if high_mode_mcts else 0.1) / (vmax1[:, i] * nb)        # HKA:1083-1085
                    wv = 0.08 / nb
                else:
                    wxz = (0.2 if high_mode_mcts else 0.1) / vmax1[:, i]              # HKA:1089-1091
                    wv = 0.08
                otw[:, i, k, 0] = np.where(far, 0.0, wxz)
                otw[:, i, k, 1] = np.where(far, 0.0, wxz)
                otw[:, i, k, 2] = np.where(far, 0.0, wv)
    return dict(x0=x0, target=target, tw=tw, cw=cw, aw=aw, otgt=otgt, otw=otw, dt=DT, teams=teams, order=order,
                track=track.name, seed=seed,
                sampled=dict(section=sec, tgt_lane=tgt_lane, nxt_lane=nxt_lane, bucket_max=bucket_max, nbucket_max=nbucket_max))


def assemble_dense(prob: dict):
    """Compact -> dense records (A[b,N,4,4], B[b,N,4,2], Q[b,N,n,n], q[b,N,n], R[b,N,2,2], x0[b,n]); vectorised numpy
    form of LinearizedBicycle.getA/getB and LQRCheckpointReachAvoidCost.getQMatrix/getQVec/getRMatrix."""
    x0, dt = prob["x0"], prob["dt"]
    b, N = x0.shape[0], x0.shape[1]
    n, K = 4 * N, N - 1
    A = np.zeros((b, N, 4, 4))
    A[..., np.arange(4), np.arange(4)] = 1.0
    h, v = x0[..., 3], x0[..., 2]
    A[..., 0, 2] = np.cos(h) * dt
    A[..., 1, 2] = np.sin(h) * dt
    A[..., 0, 3] = -np.sin(h) * dt * v
    A[..., 1, 3] = np.cos(h) * dt * v
    B = np.zeros((b, N, 4, 2))
    B[..., 2, 0] = dt
    B[..., 3, 1] = dt
    Q = np.zeros((b, N, n, n))
    q = np.zeros((b, N, n))
    aw, otgt, otw, tw, target = prob["aw"], prob["otgt"], prob["otw"], prob["tw"], prob["target"]
    for s in range(2):                                                     # KartLQRCosts.cs:64-80
        total = np.zeros((b, N))
        for k in range(K):
            t = 4 * (1 + k) + s
            w = aw[:, :, k, s]
            Q[:, :, s, t] = w
            Q[:, :, t, s] = w
            Q[:, :, t, t] = -w
            total = total - w
        Q[:, :, s, s] = total
    for s in range(4):                                                     # :81-84
        Q[:, :, s, s] += tw[:, :, s]
    for k in range(K):                                                     # :86-94 (assignment)
        for o in range(3):
            Q[:, :, 4 * (1 + k) + o, 4 * (1 + k) + o] = -otw[:, :, k, o]
    q[:, :, :4] = -target                                                  # :109
    q[:, :, :4] = q[:, :, :4] * tw                                         # :110-113
    for k in range(K):                                                     # :115-124
        q[:, :, 4 * (1 + k):4 * (2 + k)] = otgt[:, :, k]
        for o in range(3):
            q[:, :, 4 * (1 + k) + o] = q[:, :, 4 * (1 + k) + o] * -otw[:, :, k, o]
    R = np.zeros((b, N, 2, 2))
    R[..., 0, 0] = prob["cw"]
    R[..., 1, 1] = prob["cw"]
    return tuple(np.ascontiguousarray(a) for a in (A, B, Q, q, R, x0.reshape(b, n)))


def config1():
    """BASELINE config 1: one fixed 2-kart Oval problem."""
    p = make_problems(OVAL, 1, 2, seed=1)
    p["x0"][0, 0] = [15.87 + 1.25, 5.0, 12.0, np.pi / 2]
    p["x0"][0, 1] = [15.87 - 1.25, 8.0, 11.0, np.pi / 2]
    return _refresh_fixed(OVAL, p)


def _refresh_fixed(track, p):
    """Recompute weights of a hand-placed 2-kart problem (targets: lane 3 / lane 2 of section 1)."""
    x0 = p["x0"]
    lanes_xy = track.lane_table()
    for i, ln in enumerate((3, 2)):
        tl = lanes_xy[1, ln - 1]
        th = float(_wrap2pi(_mathf_atan2(tl[1] - x0[0, i, 1], tl[0] - x0[0, i, 0])))
        h0 = x0[0, i, 3]
        p["target"][0, i] = [tl[0], tl[1], 15.0, h0 - _angle_difference(h0, th)]
        v = x0[0, i, 2]
        p["tw"][0, i] = [0.3 * 3.1 / max(1.0, v), 0.3 * 3.1 / max(1.0, v), 5e-4, 3.5]
    for i in range(2):
        o = 1 - i
        d = F32(np.linalg.norm((x0[0, o, :2] - x0[0, i, :2]).astype(np.float32)))
        mult = F32(1.0 if i == 0 else 1.3)
        w = float(F32(1.0) / (_mathf_pow(d, 1.5) * mult)) if d <= 8 else 0.0
        p["aw"][0, i, 0] = [w, w]
        p["otgt"][0, i, 0] = [p["target"][0, o, 0], p["target"][0, o, 1], 15.0, 0.0]
        p["otw"][0, i, 0] = [0.2 / max(1.0, x0[0, i, 2]), 0.2 / max(1.0, x0[0, i, 2]), 0.08]
    return p


def config2(batch: int = 65536):
    """BASELINE config 2: 65,536 2-kart Oval problems, seed 20260001."""
    return make_problems(OVAL, batch, 2, seed=20260001)


def config3(batch: int = 1 << 20):
    """BASELINE config 3: 4-kart 2v2 Complex problems, seed 20260002."""
    return make_problems(COMPLEX, batch, 4, seed=20260002)
