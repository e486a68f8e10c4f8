"""CPU ORACLE of the closed loop for races with up to 4 karts and teams (TEST INFRASTRUCTURE: only tests/ may import this) — the
more-than-two-agents behaviour of HierarchicalKartAgent (Assets/Karting/Scripts/AI/HierarchicalKartAgent.cs) composed from oracle
parts only, nothing from the CUDA library or its host mirrors:
  problem recipe          oracle/np_recipe.py   (SolveLQR :699-1201, written from the C# text; N in 1..4 players after the 8 m filter)
  providers + solve       hk_oracle_lqng.c      (LinearizedBicycle, LQRCheckpointReachAvoidCost in each player's private order on the
                                                 joint state — quirk Q3 —, KartLQR.solveFeedbackLQR) with the REAL player count N
  plant + bookkeeping     hk_oracle_race.c      (hk_oracle_race_step / _plan_fixed, per kart)
  MCTS high level         root state (:180-245) and waypoint hand-off (:366-402) for K karts restated below from the C# text, the
                          sequential search of hk_oracle_mcts.c, the planner's schedule (:85-93, :172-283, :331-353, :660-661)
The solve runs every lqr_every-th step (:317), controls are held in between."""
import numpy as np

from . import np_recipe
from . import oracle as O
from . import structs as S


def _kart_dicts(karts_race):
    return [dict(x=float(k["x"]), z=float(k["z"]), v=float(k["v"]), h=float(k["h"]), section=int(k["section"]), active=bool(k["active"]),
                 team=int(k["team"])) for k in karts_race]


def recipe_agent(track_tables, params, karts_race, plans_race, beliefs_race, ego):
    L = len(track_tables["trig"])
    K = len(karts_race)
    return np_recipe.solve_lqr_recipe(track_tables, _kart_dicts(karts_race), ego, plans_race[ego]["lane"][:L], plans_race[ego]["vel"][:L],
                                      [beliefs_race[ego][o]["lane"][:L] for o in range(K)], [beliefs_race[ego][o]["vel"][:L] for o in range(K)],
                                      bool(params.highModeMcts), int(params.velocityBucketSize), float(params.topSpeed))


def solve_agent(rec, dt, horizon):
    """u0 of the ego from a recipe: providers + solveFeedbackLQR with the game's real player count."""
    N = len(rec["players"])
    n = 4 * N
    A, B = np.zeros((1, N, 4, 4)), np.zeros((1, N, 4, 2))
    Q, q, R = np.zeros((1, N, n, n)), np.zeros((1, N, n)), np.zeros((1, N, 2, 2))
    for i in range(N):
        A[0, i] = O.bicycle_A(dt, rec["x0"][i])
        B[0, i] = O.bicycle_B(dt)
        if N > 1:
            Q[0, i], q[0, i], R[0, i] = O.cost(rec["target"][i], rec["tw"][i], rec["cw"][i], rec["aw"][i], rec["otgt"][i], rec["otw"][i])
        else:                                                            # a lone player: only its own block
            Q[0, i] = np.diag(rec["tw"][i])
            q[0, i] = -rec["target"][i] * rec["tw"][i]
            R[0, i] = np.eye(2) * rec["cw"][i]
    out = O.lqng_solve_batch(A, B, Q, q, R, rec["x0"].reshape(1, n), horizon, full=False)
    return out["u0"][0, :2], int(out["status"][0])


def mcts_root(params, gparams, n_sections, karts_race, plans_race, ego):
    """planWithMCTS's root state (:180-245) for agent `ego` of a K-kart race; returns (hk_game_state, nearby list)."""
    me_sec = int(karts_race[ego]["section"])
    nearby, initial, furthest = [], me_sec, ego
    for a in range(len(karts_race)):                                     # foreach agent in m_envController.Agents :182
        if abs(int(karts_race[a]["section"]) - me_sec) < gparams.sectionWindow:
            nearby.append(a)
            initial = max(initial, int(karts_race[a]["section"]))
            if initial == int(karts_race[a]["section"]):
                furthest = a
    st = S.hk_game_state()
    st.n_karts = len(nearby)
    st.initialSection = st.lastCompletedSection = initial
    st.finalSection = initial + gparams.treeSearchDepth
    for i, a in enumerate(nearby):
        k = karts_race[a]
        t_at = 0
        if int(k["section"]) != initial:                                 # :211-214: int * float * int, (int)
            d = int(plans_race[a]["sectionTimes"][int(k["section"]) % n_sections]) - int(plans_race[furthest]["sectionTimes"][int(k["section"]) % n_sections])
            t_at = int(np.float32(np.float32(np.float32(d) * np.float32(0.02)) * np.float32(gparams.timePrecision)))
        wear = np.float32(np.float32(np.float32(4.0) - np.float32(k["steer"])) / np.float32(3.0))
        st.karts[i] = S.hk_kart_state(player=0, team=int(k["team"]), section=initial, timeAtSection=t_at, min_velocity=0,
                                      max_velocity=min(int(gparams.velocityBucketSize), int(params.topSpeed)), lane=int(k["lane"]),
                                      tireAge=int(np.float32(wear * np.float32(10000))), laneChanges=int(k["laneChanges"]), infeasible=0)
    return st, nearby


def apply_best(n_sections, karts_race, plans_race, beliefs_race, ego, nearby, best):
    """FixedUpdate's hand-off (:366-402): own lanes / velocities beyond the next checkpoint, beliefs about the other karts of the game."""
    sec = int(karts_race[ego]["section"])
    for gs in best:
        for i in range(gs.n_karts):
            ks = gs.karts[i]
            key = ks.section % n_sections
            if nearby[i] == ego:
                if ks.section > sec + (0 if sec == 0 else 1):
                    plans_race[ego]["lane"][key] = ks.lane
                    plans_race[ego]["vel"][key] = ks.max_velocity
            else:
                beliefs_race[ego][nearby[i]]["lane"][key] = ks.lane
                beliefs_race[ego][nearby[i]]["vel"][key] = ks.max_velocity


class PlannerN:
    def __init__(self, game, gparams, n_races, K, iterations, seed=0, first_iterations=0, reuse_cycles=3, apply_delay=0):
        self.game, self.gp, self.K = game, gparams, K
        self.iterations, self.first, self.reuse, self.delay, self.seed = iterations, first_iterations, reuse_cycles, apply_delay, seed
        n = n_races * K
        self.tree = [None] * n
        self.root_valid, self.cycles = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self.pending = [None] * n                                        # (fresh, nearby, best states)
        self.nearby = [None] * n
        self.pending_step = -1


def run_n(races: "O.Races", track_tables, K, lqr_every, karts, plans, beliefs, u_hold, first_step, n_steps, planner: PlannerN | None = None):
    """The loop of hk_raceN_run; arrays are updated in place.  Returns the number of solves with a zero pivot."""
    p = races.params
    n_races, L = karts.shape[0], races.n
    bad = 0
    for step in range(first_step, first_step + n_steps):
        replan = step > 0 and step % p.planEvery == 0
        if replan and not p.highModeMcts:
            races.plan_fixed(karts.reshape(-1), plans.reshape(-1))
        elif planner is not None:
            begin = step == 0 and planner.first > 0
            if (replan and step < planner.gp.maxEpisodeSteps) or begin:
                budget = planner.first if begin else planner.iterations
                seed = planner.seed + (step // p.planEvery) * n_races * K
                for r in range(n_races):
                    snapshot_k, snapshot_p = karts[r].copy(), plans[r].copy()
                    for e in range(K):
                        aid = r * K + e
                        planner.pending[aid] = None
                        if not karts[r, e]["active"]:
                            continue
                        if not (planner.reuse > 0 and planner.root_valid[aid]):
                            st, nearby = mcts_root(p, planner.gp, L, snapshot_k, snapshot_p, e)
                            planner.tree[aid] = O.Tree(planner.game, st, key=seed + aid)
                            planner.nearby[aid] = nearby
                            fresh = 1
                        elif planner.cycles[aid] < planner.reuse:
                            fresh = 0
                        else:
                            continue
                        assert planner.tree[aid].search(budget) in (0, -2)
                        planner.pending[aid] = (fresh, planner.tree[aid].best_states())
                planner.pending_step = step + planner.delay
            if planner.pending_step == step:
                for r in range(n_races):
                    for e in range(K):
                        aid = r * K + e
                        if planner.pending[aid] is None:
                            continue
                        fresh, best = planner.pending[aid]
                        planner.root_valid[aid] = 1
                        planner.cycles[aid] = 1 if fresh else planner.cycles[aid] + 1
                        apply_best(L, karts[r], plans[r], beliefs[r], e, planner.nearby[aid], best)
                        planner.pending[aid] = None
                planner.pending_step = -1
        if step % lqr_every == 0:
            for r in range(n_races):
                for e in range(K):
                    rec = recipe_agent(track_tables, p, karts[r], plans[r], beliefs[r], e)
                    u, status = solve_agent(rec, p.dt, p.horizon)
                    u_hold[r, e] = u
                    bad += status != 0
        before = karts["section"].copy()
        races.step(karts.reshape(-1), plans.reshape(-1), u_hold.reshape(-1, 2), step)
        if planner is not None:
            crossed = (karts["section"] != before).reshape(-1)
            planner.root_valid[crossed] = 0
            planner.cycles[crossed] = 0
    return bad
