/*
 * hk_oracle_game.c — CPU ORACLE (test infrastructure, NOT the product).
 *
 * Plain-C restatement (float32 + int32, compiled with -ffp-contract=off) of the reference's discrete race game and
 * of the MCTS rollout policy, bug-compatible quirks included:
 *   DiscreteKartState.computeTOC / applyAction     Assets/Karting/Scripts/AI/MCTS/KartDiscreteGame.cs:67-122, 127-171
 *   DiscreteGameState.upNext/isOver/nextMoves/makeMove                                   ...:188-243, 251-317, 322-415, 420-446
 *   KartMCTS.simulate (policy ordering + index draw), NextGaussian   Assets/Karting/Scripts/AI/MCTS/KartMCTS.cs:238-278, 204-236
 *   DiscretePositionTracker.{radiusOfLane,distanceToTravel,tireLoad,isStraight,getOptimalLaneSign}
 *                                                  Assets/Karting/Scripts/DiscretePositionTracker.cs:72-88,153-199,235-245
 *   RacingEnvController accessors                  Assets/Karting/Scripts/RacingEnvController.cs:758-784
 *   ArcadeKart.{GetMaxSpeed,getMaxLateralGsForWear,getMaxSpeedForRadiusAndWear}
 *                                                  Assets/Karting/Scripts/KartSystems/ArcadeKart.cs:210,517-520,536-547
 *
 * PARITY UNPINNED by the reference (it has no tests and cannot be compiled here); pinned instead by the
 * surveyor's independently computed known answers (SURVEY.md Appendix D -> tests/golden/game_kat.json).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include "hk_oracle.h"

struct hk_oracle_game {
    int n_sections, n_karts, n_env_karts;
    hk_section* sections;
    hk_kart karts[HK_MAX_KARTS];
    hk_kart env_karts[16];
    hk_game_params p;
};

int hk_oracle_game_create(const hk_section* s, int n_sections, const hk_kart* karts, int n_karts,
                          const hk_kart* env_karts, int n_env_karts, const hk_game_params* p, hk_oracle_game** out)
{
    if (!s || n_sections < 1 || !karts || n_karts < 1 || n_karts > HK_MAX_KARTS || !p || !out) return -1;
    if (env_karts && (n_env_karts < 1 || n_env_karts > 16)) return -1;
    hk_oracle_game* g = (hk_oracle_game*)calloc(1, sizeof(*g));
    g->n_sections = n_sections; g->n_karts = n_karts;
    g->sections = (hk_section*)malloc(sizeof(hk_section) * n_sections);
    memcpy(g->sections, s, sizeof(hk_section) * n_sections);
    memcpy(g->karts, karts, sizeof(hk_kart) * n_karts);
    if (env_karts) { g->n_env_karts = n_env_karts; memcpy(g->env_karts, env_karts, sizeof(hk_kart) * n_env_karts); }
    else { g->n_env_karts = n_karts; memcpy(g->env_karts, karts, sizeof(hk_kart) * n_karts); }
    g->p = *p;
    *out = g;
    return 0;
}
void hk_oracle_game_destroy(hk_oracle_game* g) { if (g) { free(g->sections); free(g); } }

/* C# (int) of a float on Mono/x64 = cvttss2si: NaN and out-of-range give INT_MIN */
static int f2i(float f)
{
    if (!(f > -2147483904.0f && f < 2147483648.0f)) return INT_MIN;
    return (int)f;
}

/* ---- track (DiscretePositionTracker.cs) --------------------------------------------------------------------- */
static const hk_section* sec(const hk_oracle_game* g, int section)
{
    return &g->sections[section % g->n_sections];                   /* RacingEnvController.cs:760 */
}
static int is_straight(const hk_section* s) { return s->insideR == 0.0f; }   /* :197 */
static float lane_radius(const hk_section* s, int lane)                      /* Start() :72-88 */
{
    int k = s->leftTurn ? (lane - 1) : (4 - lane);
    switch (k) {
        case 0: return s->insideR;
        case 1: return s->insideR + s->width * (1.0f / 4.0f);
        case 2: return s->insideR + s->width * (2.0f / 4.0f);
        default: return s->insideR + s->width * (3.0f / 4.0f);
    }
}
static float radius_of_lane(const hk_section* s, int a, int b)               /* :153-158 */
{
    if (is_straight(s)) return 0.0f;
    return (lane_radius(s, a) + lane_radius(s, b)) / 2.0f;
}
static float distance_to_travel(const hk_section* s, int a, int b)           /* :163-175 */
{
    if (is_straight(s)) {
        float widthTraversed = ((float)abs(a - b) * 1.0f / 3.0f) * s->width;
        return sqrtf(widthTraversed * widthTraversed + s->length * s->length);
    } else {
        float avgRad = radius_of_lane(s, a, b);
        return (3.14159274f / 180.0f) * s->turnDeg * avgRad;
    }
}
static float tire_load(const hk_section* s, float velocity, int a, int b)    /* :180-192 */
{
    if (is_straight(s)) return distance_to_travel(s, a, b) * 0.01f;
    float gs = (velocity * velocity) / radius_of_lane(s, a, b);
    return gs * distance_to_travel(s, a, b) * 0.01f;
}
static int optimal_lane_sign(const hk_section* s)                            /* :235-245 */
{
    if (s->optimalLane == 1) return 1;
    if (s->optimalLane == 4) return -1;
    return 0;
}

/* ---- kart (ArcadeKart.cs) ----------------------------------------------------------------------------------- */
static float get_max_speed(const hk_kart* k) { return k->topSpeed > k->reverseSpeed ? k->topSpeed : k->reverseSpeed; } /* :210 */
static float max_gs_for_wear(const hk_kart* k, float wear) { return (1 - wear) * (k->maxGs - k->minGs) + k->minGs; }    /* :517-520 */
float hk_oracle_max_speed_for_radius_and_wear(const hk_kart* k, float radius, float wear)                              /* :536-547 */
{
    if (radius == 0) return k->topSpeed;
    float v = sqrtf(max_gs_for_wear(k, wear) * 9.81f * fabsf(radius));
    if (isinf(v) || isnan(v)) v = k->topSpeed;
    if (v < 0.0001f) v = 0.0001f; else if (v > k->topSpeed) v = k->topSpeed;   /* Mathf.Clamp */
    return v;
}

/* ---- DiscreteKartState --------------------------------------------------------------------------------------- */
static float avg_velocity(int mn, int mx) { return (1.0f * (float)(mn + mx)) / 2.0f; }   /* :58-61 */

float hk_oracle_compute_toc(const hk_kart* k, float distance, float radius, float tireWear, float initV, float finalV) /* :67-122 */
{
    if (finalV > initV && (finalV * finalV - initV * initV) / (2 * k->accel) > distance) return -1.0f;
    if (initV > finalV && (initV * initV - finalV * finalV) / (2 * k->braking) > distance) return -1.0f;
    float ms = hk_oracle_max_speed_for_radius_and_wear(k, radius, tireWear);
    float t1, t3;
    if (ms >= initV) t1 = (ms - initV) / k->accel; else t1 = (initV - ms) / k->braking;
    if (ms >= finalV) t3 = (ms - finalV) / k->braking; else t3 = (finalV - ms) / k->accel;
    float x1 = 0.5f * (initV + ms) * t1;
    float x3 = 0.5f * (finalV + ms) * t3;
    float x2 = distance - x1 - x3;
    float t2 = x2 / ms;
    if ((double)t2 > 0.001) {
        return t1 + t2 + t3;
    } else if (initV <= ms) {
        float maxSpeed = sqrtf((2 * distance * -k->braking * k->accel + -k->braking * initV * initV - k->accel * finalV * finalV)
                               / (-k->accel - k->braking));
        t1 = (maxSpeed - initV) / k->accel;
        t3 = (maxSpeed - finalV) / k->braking;
        return t1 + t3;
    }
    return -1.0f;
}

hk_kart_state hk_oracle_apply_action(const hk_oracle_game* g, const hk_kart_state* s, hk_action a)   /* :127-171 */
{
    const hk_kart* kart = &g->env_karts[s->player];                 /* environment.Agents[player].m_Kart :129 */
    hk_kart_state ns;
    memset(&ns, 0, sizeof(ns));
    ns.team = s->team; ns.player = s->player;
    ns.section = s->section + 1;
    ns.min_velocity = a.min_velocity; ns.max_velocity = a.max_velocity; ns.lane = a.lane;
    if (is_straight(sec(g, s->section)) != is_straight(sec(g, s->section + 1))) ns.laneChanges = 0;
    else if (ns.lane != s->lane) ns.laneChanges = s->laneChanges + abs(ns.lane - s->lane);
    else ns.laneChanges = s->laneChanges;
    float dist = distance_to_travel(sec(g, s->section), s->lane, a.lane);
    float rad = radius_of_lane(sec(g, s->section), s->lane, a.lane);
    /* newState.tireAge is still 0 here => tyre wear 0 (quirk B.6-3) */
    float toc = hk_oracle_compute_toc(kart, dist, rad, (float)ns.tireAge / 10000.0f,
                                      avg_velocity(s->min_velocity, s->max_velocity), avg_velocity(ns.min_velocity, ns.max_velocity));
    int timeUpdate = f2i(toc * (float)g->p.timePrecision);
    if (timeUpdate < 0) ns.infeasible = 1;
    ns.timeAtSection = (int)((unsigned)s->timeAtSection + (unsigned)timeUpdate);   /* unchecked int add */
    float load = tire_load(sec(g, s->section), (float)a.max_velocity, s->lane, a.lane);
    ns.tireAge = f2i(((float)s->tireAge / 10000.0f + load * kart->tireWearFactor) * (float)10000);
    return ns;
}

/* ---- DiscreteGameState --------------------------------------------------------------------------------------- */
static int cmp_kart(const hk_kart_state* a, const hk_kart_state* b)          /* comparison of :191-227 */
{
    if (a->section < b->section) return -1;
    if (a->section > b->section) return 1;
    if (a->timeAtSection < b->timeAtSection) return -1;
    if (a->timeAtSection == b->timeAtSection) {
        float va = avg_velocity(a->min_velocity, a->max_velocity), vb = avg_velocity(b->min_velocity, b->max_velocity);
        if (va > vb) return -1;
        if (va == vb) return 0;
        return 1;
    }
    return 1;
}

int hk_oracle_up_next(const hk_oracle_game* g, const hk_game_state* st)      /* :188-243 */
{
    (void)g;
    int order[HK_MAX_KARTS];
    /* List<T>.Sort(Comparison) = introspective sort (.NET 4.5+ / Mono reference source ArraySortHelper.IntroSort):
     * partitions of <= 16 elements use an insertion sort (stable) EXCEPT sizes 2 and 3, which use exchange networks:
     * size 2: SwapIfGreater(0,1); size 3: SwapIfGreater(0,1), (0,2), (1,2) — the latter is not stable, and ties are
     * common at the root (equal time and bucket), so it is replayed literally.  Indices are sorted to recover the
     * original position (the reference recovers it with ValueType.Equals over all fields incl. the unique name, :232-238). */
    for (int i = 0; i < st->n_karts; ++i) order[i] = i;
#define HK_SWAP_IF_GREATER(a, b) do { if (cmp_kart(&st->karts[order[a]], &st->karts[order[b]]) > 0) { int t_ = order[a]; order[a] = order[b]; order[b] = t_; } } while (0)
    if (st->n_karts == 2) {
        HK_SWAP_IF_GREATER(0, 1);
    } else if (st->n_karts == 3) {
        HK_SWAP_IF_GREATER(0, 1); HK_SWAP_IF_GREATER(0, 2); HK_SWAP_IF_GREATER(1, 2);
    } else {
        for (int i = 1; i < st->n_karts; ++i) {
            int t = order[i], j = i - 1;
            while (j >= 0 && cmp_kart(&st->karts[t], &st->karts[order[j]]) < 0) { order[j + 1] = order[j]; --j; }
            order[j + 1] = t;
        }
    }
#undef HK_SWAP_IF_GREATER
    for (int i = 0; i < st->n_karts; ++i)
        if (st->karts[order[i]].section != st->lastCompletedSection + 1) return order[i];
    return -1;
}

int hk_oracle_next_moves(const hk_oracle_game* g, const hk_game_state* st, hk_action* out, int* gen_index) /* :322-415 */
{
    int np = hk_oracle_up_next(g, st);
    if (np < 0) return -1;                                           /* kartAgents[-1] throws */
    const hk_kart* kart = &g->karts[np];
    const hk_kart_state* cs = &st->karts[np];
    int cnt = 0, gi = 0;
    int vmax = f2i(get_max_speed(kart));
    for (int i = 6; i < vmax; i += g->p.velocityBucketSize)
        for (int j = 1; j < 5; ++j, ++gi) {
            hk_action a;
            a.min_velocity = i;
            a.max_velocity = (i + g->p.velocityBucketSize < vmax) ? i + g->p.velocityBucketSize : vmax;
            a.lane = j;
            /* lane-change rule :346 */
            if (is_straight(sec(g, cs->section)) && cs->laneChanges + abs(a.lane - cs->lane) > g->p.maxLaneChanges) continue;
            float radius = radius_of_lane(sec(g, cs->section), cs->lane, a.lane);
            /* lateral-g limit with the REAL wear :357 */
            if (hk_oracle_max_speed_for_radius_and_wear(kart, radius, (float)cs->tireAge / 10000.0f) < (float)a.min_velocity) continue;
            hk_kart_state applied = hk_oracle_apply_action(g, cs, a);   /* :368 */
            if (applied.infeasible) continue;
            if (out) out[cnt] = a;
            if (gen_index) gen_index[cnt] = gi;
            ++cnt;
        }
    /* the collision filter :385-412 is disabled (`false &&`) and the unfiltered list is returned :414 */
    return cnt;
}

hk_game_state hk_oracle_make_move(const hk_oracle_game* g, const hk_game_state* st, hk_action a, int* err)   /* :420-446 */
{
    hk_game_state ns = *st;
    int np = hk_oracle_up_next(g, st);
    if (np < 0) { if (err) *err = 1; return ns; }
    ns.karts[np] = hk_oracle_apply_action(g, &st->karts[np], a);
    int allAhead = 1;
    for (int i = 0; i < ns.n_karts; ++i) allAhead &= ns.karts[i].section > st->lastCompletedSection;
    if (allAhead) ns.lastCompletedSection += 1;
    return ns;
}

/* returns over flag; scores[] gets up to 2*HK_MAX_KARTS entries, *n_scores their count */
int hk_oracle_is_over(const hk_oracle_game* g, const hk_game_state* st, float* scores, int* n_scores)   /* :251-317 */
{
    int n = 0;
    int nm = hk_oracle_next_moves(g, st, 0, 0);
    if (nm < 0) { *n_scores = 0; return -1; }
    if (nm == 0) {
        int noMove = hk_oracle_up_next(g, st);
        for (int i = 0; i < st->n_karts; ++i) {
            if (i == noMove || st->karts[i].team == st->karts[noMove].team) scores[n++] = 0.0f;
            scores[n++] = 0.5f;                                      /* missing else => list longer than N (quirk B.6-4) */
        }
        *n_scores = n;
        return 1;
    } else if (st->lastCompletedSection != st->finalSection) {
        *n_scores = 0;
        return 0;
    } else if (st->n_karts > 1) {
        float maxScore = (float)g->p.timePrecision * -1000.0f;
        float minScore = (float)g->p.timePrecision * 1000.0f;
        float raw[HK_MAX_KARTS];
        float teamScore = 0.0f, opponentScore = 0.0f;               /* NOT reset per kart (quirk B.6-5) */
        int teamCount = 0, opponentCount = 0;
        for (int s = 0; s < st->n_karts; ++s) {
            for (int o = 0; o < st->n_karts; ++o) {
                if (s == o) {                                        /* s.Equals(o): all fields incl. the unique name */
                    teamScore += (float)st->karts[o].timeAtSection;
                } else if (st->karts[s].team == st->karts[o].team) {
                    teamScore += (float)st->karts[o].timeAtSection * g->p.teamScoreRewardMultiplier;
                    teamCount += 1;
                } else {
                    opponentScore += (float)st->karts[o].timeAtSection;
                    opponentCount += 1;
                }
            }
            float score = opponentScore * (((float)teamCount * g->p.teamScoreRewardMultiplier + 1.0f) / ((float)opponentCount * 1.0f)) - teamScore;
            raw[s] = score;
            maxScore = maxScore > score ? maxScore : score;         /* Math.Max / Math.Min (NaN-propagation unreachable: see DESIGN.md) */
            minScore = minScore < score ? minScore : score;
        }
        for (int s = 0; s < st->n_karts; ++s) {
            int sc = f2i(raw[s]);                                    /* foreach (int score in scores) truncates (quirk B.6-6) */
            scores[n++] = ((float)sc - minScore) * 1.0f / (maxScore - minScore);
        }
        *n_scores = n;
        return 1;
    } else {
        scores[0] = (float)(g->p.maxEpisodeSteps - st->karts[0].timeAtSection / g->p.maxEpisodeSteps);   /* integer division :314 */
        *n_scores = 1;
        return 1;
    }
}

/* KartMCTS.cs:252-256: stable OrderBy(dTime).ThenByDescending(max_velocity).ThenBy(|dLane|).ThenBy(optSign*lane) */
int hk_oracle_policy_moves(const hk_oracle_game* g, const hk_game_state* st, hk_action* out, int* gen_index)
{
    hk_action mv[HK_MAX_ACTIONS]; int gi[HK_MAX_ACTIONS]; int key0[HK_MAX_ACTIONS];
    int cnt = hk_oracle_next_moves(g, st, mv, gi);
    if (cnt <= 0) return cnt;
    int np = hk_oracle_up_next(g, st);
    int optSign = optimal_lane_sign(&g->sections[st->lastCompletedSection % g->n_sections]);   /* :252 */
    for (int k = 0; k < cnt; ++k) {
        hk_game_state ns = hk_oracle_make_move(g, st, mv[k], 0);
        key0[k] = (int)((unsigned)ns.karts[np].timeAtSection - (unsigned)st->karts[np].timeAtSection);
    }
    int order[HK_MAX_ACTIONS];
    for (int k = 0; k < cnt; ++k) order[k] = k;
    for (int i = 1; i < cnt; ++i) {                                  /* stable insertion sort = LINQ OrderBy semantics */
        int t = order[i], j = i - 1;
        for (; j >= 0; --j) {
            int o = order[j], c;
            if (key0[t] != key0[o]) c = key0[t] < key0[o] ? -1 : 1;
            else if (mv[t].max_velocity != mv[o].max_velocity) c = mv[t].max_velocity > mv[o].max_velocity ? -1 : 1;
            else if (abs(mv[t].lane - st->karts[np].lane) != abs(mv[o].lane - st->karts[np].lane))
                c = abs(mv[t].lane - st->karts[np].lane) < abs(mv[o].lane - st->karts[np].lane) ? -1 : 1;
            else if (optSign * mv[t].lane != optSign * mv[o].lane) c = optSign * mv[t].lane < optSign * mv[o].lane ? -1 : 1;
            else c = 0;
            if (c < 0) order[j + 1] = order[j]; else break;
        }
        order[j + 1] = t;
    }
    for (int k = 0; k < cnt; ++k) { if (out) out[k] = mv[order[k]]; if (gen_index) gen_index[k] = gi[order[k]]; }
    return cnt;
}

/* ---- random sources -------------------------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al. 2011): key = (seed lo, seed hi), counter = (c0, c1, c2, c3) */
void hk_oracle_philox4x32_10(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4])
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/*
 * Index distribution of the rollout policy for cnt > 2 legal moves (KartMCTS.cs:266-267 + NextGaussian :218-236):
 * X ~ N(0, sd = cnt/6f) redrawn while outside [-(cnt-1), cnt-1] for at most 10 attempts, else the mean (0);
 * index = RoundToInt(|X|) (half to even).  Closed form as cumulative 32-bit thresholds:
 * a uniform u in [0, 2^32) picks index = #{k : cdf[k] <= u}.  cnt <= 2 is random.Next(cnt) (uniform).
 */
static double phi(double x) { return 0.5 * erfc(-x / 1.4142135623730951); }
void hk_oracle_policy_cdf(int cnt, uint32_t* cdf)
{
    if (cnt <= 2) {
        for (int k = 0; k < cnt; ++k) cdf[k] = (k == cnt - 1) ? 0xFFFFFFFFu : (uint32_t)(((uint64_t)(k + 1) << 32) / (uint64_t)cnt);
        return;
    }
    double sd = (double)((float)cnt / 6.0f);
    double lim = (double)cnt - 1.0;
    double p[HK_MAX_ACTIONS];
    double acc_total = 2.0 * (phi(lim / sd) - 0.5);                 /* P(|X| <= cnt-1) per draw */
    double rej = 1.0 - acc_total;
    /* round-half-even of |X|: ties have probability zero, bins are [k-.5, k+.5] clipped to [0, cnt-1] */
    for (int k = 0; k < cnt; ++k) {
        double lo = k == 0 ? 0.0 : k - 0.5, hi = k + 0.5;
        if (hi > lim) hi = lim;
        p[k] = lo < hi ? 2.0 * (phi(hi / sd) - phi(lo / sd)) : 0.0;
    }
    double geo = 0.0, rp = 1.0;                                      /* 1 + r + ... + r^9 */
    for (int a = 0; a < 10; ++a) { geo += rp; rp *= rej; }
    double c = 0.0;
    for (int k = 0; k < cnt; ++k) {
        double pk = p[k] * geo + (k == 0 ? rp : 0.0);                /* 10 rejections => mean => index 0 */
        c += pk;
        double th = c * 4294967296.0;
        cdf[k] = (k == cnt - 1 || th >= 4294967295.0) ? 0xFFFFFFFFu : (uint32_t)th;
    }
}
int hk_oracle_policy_index(int cnt, const uint32_t* cdf, uint32_t u)
{
    int idx = 0;
    for (int k = 0; k < cnt - 1; ++k) idx += (cdf[k] <= u);
    return idx;
}

/* The reference's own procedure, for the distribution test: polar Box-Muller N(0,1) in double (MathNet Normal.Sample),
 * float arithmetic of NextGaussian(mean, sd, min, max), Mathf.RoundToInt(Mathf.Abs(.)).  rng = xorshift64* uniform. */
static double u01(uint64_t* s)
{
    uint64_t x = *s; x ^= x >> 12; x ^= x << 25; x ^= x >> 27; *s = x;
    return (double)((x * 0x2545F4914F6CDD1DULL) >> 11) * (1.0 / 9007199254740992.0);
}
static double normal_sample(uint64_t* s)
{
    double v1, v2, r;
    do { v1 = 2.0 * u01(s) - 1.0; v2 = 2.0 * u01(s) - 1.0; r = v1 * v1 + v2 * v2; } while (r >= 1.0 || r == 0.0);
    return v1 * sqrt(-2.0 * log(r) / r);
}
int hk_oracle_reference_policy_index(int cnt, uint64_t* rng_state)
{
    if (cnt <= 2) return (int)(u01(rng_state) * cnt);               /* random.Next(cnt) */
    float sd = (float)cnt / 6.0f, mn = -(float)cnt + 1.0f, mx = (float)cnt - 1.0f, x;
    int attempts = 0;
    do { x = 0.0f + (float)normal_sample(rng_state) * sd; attempts += 1; } while ((x < mn || x > mx) && attempts < 10);
    if (attempts == 10 && (x < mn || x > mx)) x = 0.0f;
    return (int)nearbyintf(fabsf(x));                                /* Mathf.RoundToInt: half to even */
}

/*
 * One rollout = KartMCTS.simulate (:238-278) without the tree bookkeeping.
 * mode 0: index from the Philox/CDF sampler shared with the GPU (rollout id, ply) -> bit-reproducible;
 * mode 1: index from the reference's own Gaussian procedure (rng_state).
 * Returns number of plies, or -1 on upNext()==-1.
 */
int hk_oracle_rollout(const hk_oracle_game* g, const hk_game_state* leaf, int mode, uint64_t seed, uint64_t rollout_id,
                      uint64_t* rng_state, hk_action* actions_out, int* choice_out, float* scores, int* n_scores,
                      hk_game_state* terminal)
{
    hk_game_state st = *leaf;
    int ply = 0;
    static uint32_t cdfs[HK_MAX_ACTIONS + 1][HK_MAX_ACTIONS];
    static int cdf_ready = 0;
    if (!cdf_ready) { for (int c = 1; c <= HK_MAX_ACTIONS; ++c) hk_oracle_policy_cdf(c, cdfs[c]); cdf_ready = 1; }
    for (;;) {
        int over = hk_oracle_is_over(g, &st, scores, n_scores);
        if (over < 0) return -1;
        if (over) break;
        hk_action mv[HK_MAX_ACTIONS];
        int cnt = hk_oracle_policy_moves(g, &st, mv, 0);
        int index;
        if (mode == 0) {
            uint32_t r[4];
            hk_oracle_philox4x32_10(seed, (uint32_t)rollout_id, (uint32_t)(rollout_id >> 32), (uint32_t)ply, 0u, r);
            index = hk_oracle_policy_index(cnt, cdfs[cnt], r[0]);
        } else {
            index = hk_oracle_reference_policy_index(cnt, rng_state);
        }
        if (ply < HK_MAX_PLIES) { if (actions_out) actions_out[ply] = mv[index]; if (choice_out) choice_out[ply] = index; }
        st = hk_oracle_make_move(g, &st, mv[index], 0);
        ++ply;
    }
    if (terminal) *terminal = st;
    return ply;
}

/* Leaf-parallel statistics with the same reduction as hk_mcts_rollouts (include/hk_abi.h) */
int hk_oracle_rollouts(const hk_oracle_game* g, const hk_game_state* leaf, int64_t n_rollouts, int mode, uint64_t seed,
                       uint64_t rollout_offset, int64_t* visit, double* reward_sum, int64_t* nan_count, int64_t* plies_sum)
{
    memset(visit, 0, sizeof(int64_t) * HK_MAX_ACTIONS);
    memset(reward_sum, 0, sizeof(double) * HK_MAX_ACTIONS * HK_MAX_KARTS);
    memset(nan_count, 0, sizeof(int64_t) * HK_MAX_ACTIONS);
    int64_t plies = 0;
    uint64_t rng = seed * 0x9E3779B97F4A7C15ULL + 0x1234567ULL; if (!rng) rng = 1;
    const int bucket = g->p.velocityBucketSize;
    for (int64_t r = 0; r < n_rollouts; ++r) {
        hk_action acts[HK_MAX_PLIES]; float scores[2 * HK_MAX_KARTS]; int ns = 0;
        int np = hk_oracle_rollout(g, leaf, mode, seed, rollout_offset + (uint64_t)r, &rng, acts, 0, scores, &ns, 0);
        if (np < 0) return -1;
        if (np == 0) continue;
        plies += np;
        int a = ((acts[0].min_velocity - 6) / bucket) * 4 + acts[0].lane - 1;
        visit[a] += 1;
        int has_nan = 0;
        for (int k = 0; k < leaf->n_karts && k < ns; ++k) has_nan |= isnan(scores[k]);
        if (has_nan) { nan_count[a] += 1; continue; }
        for (int k = 0; k < leaf->n_karts && k < ns; ++k) reward_sum[a * HK_MAX_KARTS + k] += (double)scores[k];
    }
    if (plies_sum) *plies_sum = plies;
    return 0;
}

/* exported wrappers of the track formulas for the known-answer tests */
float hk_oracle_distance_to_travel(const hk_section* s, int a, int b) { return distance_to_travel(s, a, b); }
float hk_oracle_radius_of_lane(const hk_section* s, int a, int b) { return radius_of_lane(s, a, b); }
float hk_oracle_tire_load(const hk_section* s, float velocity, int a, int b) { return tire_load(s, velocity, a, b); }
