"""Second, independent restatement of the reference's discrete race game in numpy float32 / int32 (TEST INFRASTRUCTURE: only tests/
may import this).  Written from the C# TEXT — not from oracle/hk_oracle_game.c — so that the C oracle has a second opinion that is
not a transcription of itself:
  DiscreteKartState.getAverageVelocity / computeTOC / applyAction   Assets/Karting/Scripts/AI/MCTS/KartDiscreteGame.cs:58-61, 67-122, 127-171
  DiscreteGameState.upNext / isOver / nextMoves / makeMove          ...:188-243, 251-317, 322-415, 420-446
  DiscretePositionTracker.Start / radiusOfLane / distanceToTravel / tireLoad / isStraight / getOptimalLaneSign
                                                                    Assets/Karting/Scripts/DiscretePositionTracker.cs:72-88, 153-199, 235-245
  RacingEnvController accessors (section % Sections.Length)         Assets/Karting/Scripts/RacingEnvController.cs:758-784
  ArcadeKart.GetMaxSpeed / getMaxLateralGsForWear / getMaxSpeedForRadiusAndWear
                                                                    Assets/Karting/Scripts/KartSystems/ArcadeKart.cs:210, 517-520, 536-547
  KartMCTS.simulate's ordering of the legal moves                   Assets/Karting/Scripts/AI/MCTS/KartMCTS.cs:252-256
The arithmetic core (`apply_actions`) is vectorised over (kart state, action) pairs so that 10^5 pairs can be compared bit for bit with
the C oracle; the game-level functions use it per state.  C# float expressions are evaluated operation by operation in np.float32
(Mathf.Sqrt(x) = (float)Math.Sqrt((double)x) is the correctly rounded float32 square root); (int) of a float follows Mono/x64
cvttss2si (NaN / out of range -> int.MinValue).  `game` objects expose the interface of oracle.oracle.Game, so oracle/np_mcts_seq.py
can run the whole sequential tree search on this module instead of the C game."""
import numpy as np

from . import structs as S

F = np.float32
INT_MIN = np.int32(-2147483648)


def cs_int(x):
    """C# (int)float on Mono / .NET x64: truncation, int.MinValue for NaN and values outside the int range."""
    x = np.asarray(x, dtype=np.float32)
    ok = np.isfinite(x) & (x > F(-2147483904.0)) & (x < F(2147483648.0))
    out = np.full(x.shape, INT_MIN, dtype=np.int32)
    with np.errstate(invalid="ignore"):
        out[ok] = np.trunc(x[ok]).astype(np.int64).astype(np.int32)
    return out


class NpGame:
    def __init__(self, sections, n_sections, karts, n_karts, params, env_karts=None, n_env_karts=0):
        def table(arr, n, fields):
            return {f: np.array([getattr(arr[i], f) for i in range(n)]) for f in fields}
        sec = table(sections, n_sections, ("insideR", "length", "width", "turnDeg", "leftTurn", "optimalLane"))
        self.L = n_sections
        self.insideR, self.length, self.width, self.turnDeg = (sec[k].astype(np.float32) for k in ("insideR", "length", "width", "turnDeg"))
        self.leftTurn, self.optimalLane = sec["leftTurn"].astype(bool), sec["optimalLane"].astype(np.int32)
        # DiscretePositionTracker.Start (:72-88): radiuses[lane - 1]
        q = [F(1.0) / F(4.0), F(2.0) / F(4.0), F(3.0) / F(4.0)]
        rad = np.zeros((n_sections, 4), np.float32)
        for s in range(n_sections):
            ladder = [self.insideR[s]] + [self.insideR[s] + self.width[s] * qq for qq in q]
            rad[s] = ladder if self.leftTurn[s] else ladder[::-1]
        self.radiuses = rad
        kf = ("accel", "braking", "topSpeed", "reverseSpeed", "maxGs", "minGs", "tireWearFactor")
        self.karts = {k: v.astype(np.float32) for k, v in table(karts, n_karts, kf).items()}
        if env_karts is None:
            self.env_karts = self.karts
        else:
            self.env_karts = {k: v.astype(np.float32) for k, v in table(env_karts, n_env_karts, kf).items()}
        self.p = params
        self.n_karts = n_karts

    # ---- track / kart formulas, vectorised ----------------------------------------------------------------------------------------
    def is_straight(self, section):
        return self.insideR[np.asarray(section) % self.L] == F(0.0)                              # :197, RacingEnvController.cs:760

    def radius_of_lane(self, section, a, b):
        s = np.asarray(section) % self.L
        r = (self.radiuses[s, np.asarray(a) - 1] + self.radiuses[s, np.asarray(b) - 1]) / F(2.0)
        return np.where(self.is_straight(section), F(0.0), r).astype(np.float32)                 # :153-158

    def distance_to_travel(self, section, a, b):
        s = np.asarray(section) % self.L
        w = ((np.abs(np.asarray(a) - np.asarray(b)).astype(np.float32) * F(1.0)) / F(3.0)) * self.width[s]
        straight = np.sqrt(w * w + self.length[s] * self.length[s])                              # :167-168
        turn = ((F(3.14159274) / F(180.0)) * self.turnDeg[s]) * self.radius_of_lane(section, a, b)   # :172-173
        return np.where(self.is_straight(section), straight, turn).astype(np.float32)

    def tire_load(self, section, velocity, a, b):
        d = self.distance_to_travel(section, a, b)
        v = np.asarray(velocity, dtype=np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            gs = (v * v) / self.radius_of_lane(section, a, b)
            turn = (gs * d) * F(0.01)                                                            # :188-190
        return np.where(self.is_straight(section), d * F(0.01), turn).astype(np.float32)        # :184

    @staticmethod
    def max_speed_for_radius_and_wear(k, radius, wear):
        """ArcadeKart.cs:536-547 with getMaxLateralGsForWear :517-520; k = dict of per-row kart constants."""
        radius, wear = np.asarray(radius, np.float32), np.asarray(wear, np.float32)
        gs = (F(1) - wear) * (k["maxGs"] - k["minGs"]) + k["minGs"]
        with np.errstate(invalid="ignore"):
            v = np.sqrt((gs * F(9.81)) * np.abs(radius))
        v = np.where(np.isinf(v) | np.isnan(v), k["topSpeed"], v)
        v = np.where(v < F(0.0001), F(0.0001), np.where(v > k["topSpeed"], k["topSpeed"], v))     # Mathf.Clamp
        return np.where(radius == F(0), k["topSpeed"], v).astype(np.float32)

    @staticmethod
    def compute_toc(k, distance, radius, wear, initV, finalV):
        """KartDiscreteGame.cs:67-122, vectorised; every intermediate is float32, `t2 > 0.001` compares in double."""
        acc, brk = k["accel"], k["braking"]
        distance, initV, finalV = (np.asarray(x, np.float32) for x in (distance, initV, finalV))
        with np.errstate(all="ignore"):
            bad1 = (finalV > initV) & (((finalV * finalV - initV * initV) / (F(2) * acc)) > distance)
            bad2 = (initV > finalV) & (((initV * initV - finalV * finalV) / (F(2) * brk)) > distance)
            ms = NpGame.max_speed_for_radius_and_wear(k, radius, wear)
            t1 = np.where(ms >= initV, (ms - initV) / acc, (initV - ms) / brk).astype(np.float32)
            t3 = np.where(ms >= finalV, (ms - finalV) / brk, (finalV - ms) / acc).astype(np.float32)
            x1 = (F(0.5) * (initV + ms)) * t1
            x3 = (F(0.5) * (finalV + ms)) * t3
            x2 = (distance - x1) - x3
            t2 = x2 / ms
            long_enough = t2.astype(np.float64) > 0.001
            cruise = (t1 + t2) + t3
            num = (((F(2) * distance) * -brk) * acc + ((-brk * initV) * initV)) - ((acc * finalV) * finalV)
            peak = np.sqrt(num / (-acc - brk))
            short = ((peak - initV) / acc) + ((peak - finalV) / brk)
        out = np.where(long_enough, cruise, np.where(initV <= ms, short, F(-1.0)))
        return np.where(bad1 | bad2, F(-1.0), out).astype(np.float32)

    def apply_actions(self, ks, a_min, a_max, a_lane):
        """applyAction (:127-171) for arrays of kart states (structured array with hk_kart_state's fields) and actions."""
        k = {f: v[ks["player"]] for f, v in self.env_karts.items()}                              # environment.Agents[player].m_Kart :129
        out = np.zeros(ks.shape, dtype=ks.dtype)
        out["team"], out["player"] = ks["team"], ks["player"]
        out["section"] = ks["section"] + 1
        out["min_velocity"], out["max_velocity"], out["lane"] = a_min, a_max, a_lane
        flip = self.is_straight(ks["section"]) != self.is_straight(ks["section"] + 1)
        out["laneChanges"] = np.where(flip, 0, np.where(a_lane != ks["lane"], ks["laneChanges"] + np.abs(a_lane - ks["lane"]), ks["laneChanges"]))
        dist = self.distance_to_travel(ks["section"], ks["lane"], a_lane)
        rad = self.radius_of_lane(ks["section"], ks["lane"], a_lane)
        init_v = (F(1.0) * (ks["min_velocity"] + ks["max_velocity"]).astype(np.float32)) / F(2.0)
        final_v = (F(1.0) * (np.asarray(a_min) + np.asarray(a_max)).astype(np.float32)) / F(2.0)
        wear0 = np.zeros(ks.shape, np.float32) / F(10000)                                        # newState.tireAge is still 0 here (:154)
        toc = self.compute_toc(k, dist, rad, wear0, init_v, final_v)
        time_update = cs_int(toc * F(self.p.timePrecision))
        out["infeasible"] = (time_update < 0).astype(np.int32)
        out["timeAtSection"] = (ks["timeAtSection"].astype(np.uint32) + time_update.astype(np.uint32)).view(np.int32)   # unchecked int add
        load = self.tire_load(ks["section"], np.asarray(a_max).astype(np.float32), ks["lane"], a_lane)
        out["tireAge"] = cs_int((ks["tireAge"].astype(np.float32) / F(10000) + load * k["tireWearFactor"]) * F(10000))
        return out

    # ---- game level (oracle.oracle.Game interface; `st` = anything with hk_game_state's bytes) ------------------------------------------
    @staticmethod
    def _rec(st):
        return np.frombuffer(bytes(st) if not isinstance(st, np.void) else st.tobytes(), dtype=S.GAME_STATE_DTYPE)[0].copy()

    @staticmethod
    def _avg(k):
        return (F(1.0) * F(int(k["min_velocity"]) + int(k["max_velocity"]))) / F(2.0)

    def up_next(self, st) -> int:
        """:188-243.  List.Sort(Comparison) is an introspective sort: sizes 2 and 3 use exchange networks (size 3 is NOT stable), larger
        lists an insertion sort for <= 16 elements; the first sorted kart whose section != lastCompletedSection + 1 is looked up by value
        (the name field makes karts distinct)."""
        g = self._rec(st)
        n = int(g["n_karts"])
        ks = [g["karts"][i] for i in range(n)]

        def cmp(a, b):
            if a["section"] < b["section"]:
                return -1
            if a["section"] > b["section"]:
                return 1
            if a["timeAtSection"] < b["timeAtSection"]:
                return -1
            if a["timeAtSection"] == b["timeAtSection"]:
                va, vb = self._avg(a), self._avg(b)
                return -1 if va > vb else (0 if va == vb else 1)
            return 1
        order = list(range(n))

        def swap_if_greater(i, j):
            if cmp(ks[order[i]], ks[order[j]]) > 0:
                order[i], order[j] = order[j], order[i]
        if n == 2:
            swap_if_greater(0, 1)
        elif n == 3:
            swap_if_greater(0, 1); swap_if_greater(0, 2); swap_if_greater(1, 2)
        elif n > 3:
            for i in range(1, n):
                t, j = order[i], i - 1
                while j >= 0 and cmp(ks[t], ks[order[j]]) < 0:
                    order[j + 1] = order[j]
                    j -= 1
                order[j + 1] = t
        for i in order:
            if ks[i]["section"] != g["lastCompletedSection"] + 1:
                return i
        return -1

    def _candidates(self, kart_index):
        vmax = int(cs_int(np.maximum(self.karts["topSpeed"][kart_index], self.karts["reverseSpeed"][kart_index])))   # (int)GetMaxSpeed()
        b = self.p.velocityBucketSize
        out = []
        for i in range(6, vmax, b):
            for j in range(1, 5):
                out.append((i, min(i + b, vmax), j))
        return out

    def _legal(self, g, np_):
        cur = g["karts"][np_]
        cand = self._candidates(np_)
        if not cand:
            return [], [], None
        a_min, a_max, a_lane = (np.array(c, np.int32) for c in zip(*cand))
        n = len(cand)
        ks = np.repeat(np.array([cur]), n)
        ok = ~(self.is_straight(cur["section"]) & ((cur["laneChanges"] + np.abs(a_lane - cur["lane"])) > self.p.maxLaneChanges))   # :346
        radius = self.radius_of_lane(ks["section"], ks["lane"], a_lane)
        k = {f: np.repeat(v[np_], n) for f, v in self.karts.items()}                             # kartAgents[nextPlayer].m_Kart :326
        ok &= ~(self.max_speed_for_radius_and_wear(k, radius, ks["tireAge"].astype(np.float32) / F(10000)) < a_min.astype(np.float32))   # :357
        applied = self.apply_actions(ks, a_min, a_max, a_lane)                                   # :368
        ok &= applied["infeasible"] == 0
        idx = np.flatnonzero(ok)
        return [cand[i] for i in idx], [int(i) for i in idx], applied[idx]

    def next_moves(self, st):
        g = self._rec(st)
        np_ = self.up_next(st)
        if np_ < 0:
            return [], [], -1
        mv, gi, _ = self._legal(g, np_)
        return mv, gi, len(mv)

    def policy_moves(self, st):
        """KartMCTS.cs:252-256: stable OrderBy(time difference).ThenByDescending(max_velocity).ThenBy(|lane difference|).ThenBy(sign * lane)."""
        g = self._rec(st)
        np_ = self.up_next(st)
        if np_ < 0:
            return [], [], -1
        mv, gi, applied = self._legal(g, np_)
        if not mv:
            return [], [], 0
        cur = g["karts"][np_]
        sign_lane = int(self.optimalLane[int(g["lastCompletedSection"]) % self.L])
        sign = 1 if sign_lane == 1 else (-1 if sign_lane == 4 else 0)
        dt = (applied["timeAtSection"].astype(np.int64) - int(cur["timeAtSection"])).astype(np.int32)
        keys = [(int(dt[i]), -mv[i][1], abs(mv[i][2] - int(cur["lane"])), sign * mv[i][2]) for i in range(len(mv))]
        order = sorted(range(len(mv)), key=lambda i: keys[i])                                     # Python's sort is stable, like LINQ's
        return [mv[i] for i in order], [gi[i] for i in order], len(mv)

    def make_move(self, st, a):
        g = self._rec(st)
        np_ = self.up_next(st)
        a = S.action(a)
        new = self.apply_actions(np.array([g["karts"][np_]]), np.array([a.min_velocity], np.int32), np.array([a.max_velocity], np.int32),
                                 np.array([a.lane], np.int32))[0]
        last = int(g["lastCompletedSection"])
        g["karts"][np_] = new
        if all(int(g["karts"][i]["section"]) > last for i in range(int(g["n_karts"]))):
            g["lastCompletedSection"] = last + 1
        return S.hk_game_state.from_buffer_copy(g.tobytes())

    def is_over(self, st):
        g = self._rec(st)
        n = int(g["n_karts"])
        mv, _, cnt = self.next_moves(st)
        if cnt < 0:
            return -1, np.zeros(0, np.float32)
        if cnt == 0:                                                                             # :253-266 (no `else`: longer than n)
            no_move = self.up_next(st)
            scores = []
            for i in range(n):
                if i == no_move or g["karts"][i]["team"] == g["karts"][no_move]["team"]:
                    scores.append(F(0.0))
                scores.append(F(0.5))
            return 1, np.array(scores, np.float32)
        if g["lastCompletedSection"] != g["finalSection"]:
            return 0, np.zeros(0, np.float32)
        if n > 1:                                                                                # :271-310, accumulators never reset
            tp = F(self.p.timePrecision)
            max_score, min_score = tp * F(-1000.0), tp * F(1000.0)
            mult = F(self.p.teamScoreRewardMultiplier)
            team_score = opp_score = F(0.0)
            team_count = opp_count = 0
            raw = []
            with np.errstate(all="ignore"):
                for s in range(n):
                    for o in range(n):
                        t_o = F(int(g["karts"][o]["timeAtSection"]))
                        if s == o:
                            team_score = F(team_score + t_o)
                        elif g["karts"][s]["team"] == g["karts"][o]["team"]:
                            team_score = F(team_score + F(t_o * mult))
                            team_count += 1
                        else:
                            opp_score = F(opp_score + t_o)
                            opp_count += 1
                    score = F(F(opp_score * F(F(F(F(team_count) * mult) + F(1.0)) / F(F(opp_count) * F(1.0)))) - team_score)
                    raw.append(score)
                    # Math.Max(val1, val2): val1 > val2 ? val1 : (IsNaN(val1) ? val1 : val2);  Math.Min likewise with <
                    max_score = max_score if (max_score > score or np.isnan(max_score)) else score
                    min_score = min_score if (min_score < score or np.isnan(min_score)) else score
                out = [F(F(F(F(int(cs_int(r))) - min_score) * F(1.0)) / F(max_score - min_score)) for r in raw]   # foreach (int score ...) :305
            return 1, np.array(out, np.float32)
        mes = int(self.p.maxEpisodeSteps)                                                        # :314: int - int / int, then float
        t = int(g["karts"][0]["timeAtSection"])
        q = abs(t) // mes * (1 if t >= 0 else -1)                                                # C# integer division truncates toward zero
        return 1, np.array([F(mes - q)], np.float32)
