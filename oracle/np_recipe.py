"""Second, independent restatement of the LQNG problem recipe of HierarchicalKartAgent.SolveLQR (TEST INFRASTRUCTURE: only tests/ may
import this) — written from the C# text (Assets/Karting/Scripts/AI/HierarchicalKartAgent.cs:699-1201), NOT from the C oracle
(oracle/hk_oracle_race.c) and not from the product package, so that the C recipe is pinned by something that is neither its own
transcription nor product code.  Scalar Python / numpy float32 + float64, one problem at a time, for ANY number of karts in the
environment (2-kart head-to-head scenes and the 4-kart Duos scenes):

  players                 :702-725   [this] + teamAgents + otherAgents; with more than 2 agents in the environment only those within
                                     8 m of the ego (float32 magnitude) take part, nearbyAgents = max(#kept - 1, 1)
  initial states          :730-739   (x, z, |v|, atan2 heading wrapped to [0, 2 pi) in float32)
  targets                 :745-817   lane collider / Trigger of section + 1 and section + 2 from the ego's own plan (k == this) or from its
                                     belief about kart k; velocity min(maxSpeed, planned + 2 bucket) in MCTS mode; 0 if |v| <= 5
  target heading          :819-926   raycast-free branches only (Unity's Physics.Raycast is out of scope): "follow" when the kart is
                                     within 4 m of the checkpoint centre (stand-in for centerLine.ClosestPoint), the blended "normal
                                     case" near the target, the plain case far from it; always mirrored about the current heading
  own target weights      :928-962
  avoid / opponent-target / teammate-target weights in player k's PRIVATE order [its otherAgents..., its teamAgents...]  :964-1190
  control weight          :1192-1196
Unity's Mathf.X(float) is (float)Math.X((double)x); Vector3.magnitude is (float)Math.Sqrt((double)(x*x + y*y + z*z)) on float32 components.
Unity subtracts float32 positions; the kinematic stand-in keeps kart positions in double, so a position difference is formed in double
and rounded to float32 once — identical to Unity's float32 subtraction whenever both positions are float32 values (the difference of two
float32 numbers is exact in double).
otherAgents / teamAgents are serialized scene arrays; they are taken in environment order (opponents = other teams, mates = same team)."""
import math

import numpy as np

F = np.float32
PI_F = F(3.14159274)


def _atan2f(y, x):
    return F(math.atan2(float(F(y)), float(F(x))))


def _mag(dx, dz):
    dx, dz = F(dx), F(dz)
    return F(math.sqrt(float(F(dx * dx) + F(dz * dz))))


def _angle_difference(a1, a2):                                          # :1341-1344
    return math.atan2(math.sin(a2 - a1), math.cos(a2 - a1))


def solve_lqr_recipe(track, karts, ego, own_lane, own_vel, belief_lane, belief_vel, high_mode_mcts, bucket, top_speed):
    """track: dict(trig [L][2], lane [L][4][2], straight [L] bool).  karts: list of dicts (x, z, v, h, section, active, team), the
    environment's Agents in order.  own_lane / own_vel [L]: m_UpcomingLanes / m_UpcomingVelocities of the ego (lane 0 = key absent).
    belief_lane / belief_vel [n_agents][L]: opponentUpcomingLanes / Velocities of the ego about every other agent.
    Returns dict(players, x0 [N][4], target [N][4], tw [N][4], cw [N], aw [N][N-1][2], otgt [N][N-1][4], otw [N][N-1][3])."""
    L = len(track["trig"])
    n_env = len(karts)
    me = karts[ego]
    mates_of = lambda k: [a for a in range(n_env) if a != k and karts[a]["team"] == karts[k]["team"]]
    others_of = lambda k: [a for a in range(n_env) if karts[a]["team"] != karts[k]["team"]]
    all_players = [ego] + mates_of(ego) + others_of(ego)                # :702
    fixed = not high_mode_mcts

    def dist(a, b):                                                     # (a.position - b.position).magnitude, y equal
        return _mag(F(karts[a]["x"] - karts[b]["x"]), F(karts[a]["z"] - karts[b]["z"]))
    nearby = -1
    if n_env > 2:                                                       # :709-721
        actual = []
        for k in all_players:
            if dist(k, ego) < 8:
                nearby += 1
                actual.append(k)
    else:
        actual = list(all_players)
    nearby = max(nearby, 1)                                             # :726
    N = len(actual)

    def plan_target(k, idx):
        if k == ego:
            ln, vl = int(own_lane[idx]), float(own_vel[idx])
        else:
            ln, vl = int(belief_lane[k][idx]), float(belief_vel[k][idx])
        if ln != 0:
            xz = track["lane"][idx][ln - 1]
            v = min(float(top_speed), vl + (bucket * 2 if high_mode_mcts else 0))       # :757, :771
        else:
            xz = track["trig"][idx]
            v = float(top_speed)
        return (float(xz[0]), float(xz[1])), v

    out = dict(players=actual, x0=np.zeros((N, 4)), target=np.zeros((N, 4)), tw=np.zeros((N, 4)), cw=np.zeros(N),
               aw=np.zeros((N, max(N - 1, 0), 2)), otgt=np.zeros((N, max(N - 1, 0), 4)), otw=np.zeros((N, max(N - 1, 0), 3)))
    for i, k in enumerate(actual):
        kk = karts[k]
        x, z, v, h = float(kk["x"]), float(kk["z"]), float(kk["v"]), float(kk["h"])
        out["x0"][i] = [x, z, v, h]                                      # :730-736 (the plant keeps h wrapped to [0, 2 pi))
        s = kk["section"] + 1                                            # :745
        idx, idx2 = s % L, (s + 1) % L
        (lx, lz), vel = plan_target(k, idx)
        (nx, nz), nvel = plan_target(k, idx2)
        cx, cz = float(track["trig"][idx][0]), float(track["trig"][idx][1])
        stopped = F(v) <= F(5.0)                                          # :808
        tx, tz, tv = lx, lz, (0.0 if stopped else vel)
        th_tgt = _atan2f(F(lz - z), F(lx - x))                      # :819
        if th_tgt < 0:
            th_tgt = F(th_tgt + F(2) * PI_F)
        near = _mag(F(lx - x), F(lz - z)) <= (F(10.5) if track["straight"][kk["section"] % L] else F(7.5))   # :821
        if near:
            f1 = _atan2f(F(lz - z), F(lx - x))
            f2 = _atan2f(F(nz - lz), F(nx - lx))
            f6 = _atan2f(F(nz - z), F(nx - x))
            if _mag(F(cx - x), F(cz - z)) <= F(4.0):               # :877-890 (ClosestPoint stand-in: the trigger centre)
                tx, tz = nx, nz
                if F(v) > F(5.0):
                    tv = nvel
                if f6 < 0:
                    f6 = F(f6 + F(2) * PI_F)
                final = float(f6)
                final = h - _angle_difference(h, final)
            else:                                                         # :891-902 normal case
                if f1 < 0:
                    f1 = F(f1 + F(2) * PI_F)
                if f2 < 0:
                    f2 = F(f2 + F(2) * PI_F)
                final = float(f1) - _angle_difference(float(f2), float(f1)) * float(F(0.4))
                if final < 0:
                    final += float(F(2) * PI_F)
                final = h - _angle_difference(h, final)
        else:
            final = h - _angle_difference(h, float(th_tgt))               # :919-923
        out["target"][i] = [tx, tz, tv, final]
        # own target weights :928-962
        vmax1 = max(1.0, v)
        if N > 2:
            w_h = (2.5 if fixed else 3.5) * nearby
        else:
            w_h = 1.9 if fixed else 3.5
        if stopped:
            w_xz, w_v = nearby * 0.3 * 3.1, float(nearby * -2)
        else:
            w_xz, w_v = nearby * 0.3 * 3.1 / vmax1, nearby * 5e-4
        out["tw"][i] = [w_xz, w_xz, w_v, w_h]
        # multiplier :977-1003
        if n_env > 2 and N > 2:
            mult = F(F(0.55 if fixed else 1.0) / F(nearby)) if k == ego else F(F(1.7) / F(nearby))
        else:
            mult = F(0.45 if fixed else 1.0) if k == ego else F(1.3)
        slot = 0
        nearby_opponents = 0
        for o in others_of(k):                                            # :1004-1096
            if o not in actual:
                continue
            d = dist(o, k)
            off = d > 8 or not karts[o]["active"]
            if off:
                w = 0.0
            else:
                w = float(F(1.0) / F(F(math.pow(float(d), float(F(1.5)))) * mult))      # 1f / (Mathf.Pow(d, 1.5f) * multiplier) :1019
                nearby_opponents += 1
            out["aw"][i, slot] = [w, w]
            (ox, oz), ovel = plan_target(o, (karts[o]["section"] + 1) % L)
            out["otgt"][i, slot] = [ox, oz, ovel, 0.0]
            if off:
                out["otw"][i, slot] = [0.0, 0.0, 0.0]
            elif N > 2:
                wxz = (0.1 if fixed else 0.2) / (vmax1 * nearby)
                out["otw"][i, slot] = [wxz, wxz, 0.08 / nearby]
            else:
                wxz = (0.1 if fixed else 0.2) / vmax1
                out["otw"][i, slot] = [wxz, wxz, 0.08]
            slot += 1
        for o in mates_of(k):                                             # :1099-1190
            if o not in actual:
                continue
            d = dist(o, k)
            off = d > 8 or not karts[o]["active"]
            if off:
                w = 0.0
            else:
                mult2 = F(mult / F(2.0))                                   # :1113
                w = float(F(1.0) / F(F(math.pow(float(d), float(F(1.5)))) * mult2))
            out["aw"][i, slot] = [w, w]
            (ox, oz), _ = plan_target(o, (karts[o]["section"] + 1) % L)
            out["otgt"][i, slot] = [ox, oz, float(top_speed), 0.0]        # getMaxSpeedForState() stand-in; its weight is 0 (:1180, :1186)
            if off or nearby_opponents < 1:
                out["otw"][i, slot] = [0.0, 0.0, 0.0]
            elif N > 2:
                wxz = -(0.0 if fixed else 3e-5) / (vmax1 * nearby)
                out["otw"][i, slot] = [wxz, wxz, 0.0]
            else:
                wxz = -(1e-4 if fixed else 2e-4) / vmax1
                out["otw"][i, slot] = [wxz, wxz, 0.0]
            slot += 1
        out["cw"][i] = (0.135 if fixed else 0.25) if N > 2 else 0.115     # :1192-1196
    return out


def race_recipe_2kart(track, params, karts_race, plans_race, ego):
    """The 2-kart races of hk_race_* (each kart its own team; the ego's single belief table is about the other kart)."""
    ks = [dict(x=float(karts_race[a]["x"]), z=float(karts_race[a]["z"]), v=float(karts_race[a]["v"]), h=float(karts_race[a]["h"]),
               section=int(karts_race[a]["section"]), active=bool(karts_race[a]["active"]), team=a) for a in range(2)]
    L = len(track["trig"])
    belief_lane = [plans_race[ego]["oppLane"][:L]] * 2
    belief_vel = [plans_race[ego]["oppVel"][:L]] * 2
    return solve_lqr_recipe(track, ks, ego, plans_race[ego]["lane"][:L], plans_race[ego]["vel"][:L], belief_lane, belief_vel,
                            bool(params.highModeMcts), int(params.velocityBucketSize), float(params.topSpeed))
