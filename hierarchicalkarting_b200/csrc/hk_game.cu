// hk_game.cu — discrete race game + leaf-parallel MCTS rollouts for sm_100a.  Compile with -fmad=false: the game is
// float32/int32 and must be bit-exact with the reference's C# arithmetic; IEEE div/sqrt are nvcc's
// defaults (-prec-div=true -prec-sqrt=true, no -use_fast_math).
//
// Device restatement of (reference files under Assets/Karting/Scripts/):
//   DiscreteKartState.computeTOC / applyAction                 AI/MCTS/KartDiscreteGame.cs:67-122, 127-171
//   DiscreteGameState.upNext / isOver / nextMoves / makeMove   AI/MCTS/KartDiscreteGame.cs:188-243, 251-317, 322-415, 420-446
//   KartMCTS.simulate rollout policy                            AI/MCTS/KartMCTS.cs:238-278 (ordering :256, index draw :266-269)
//   DiscretePositionTracker track formulas                      DiscretePositionTracker.cs:72-88,153-199,235-245
//   ArcadeKart speed limits                                     KartSystems/ArcadeKart.cs:210,517-520,536-547
// One thread plays one rollout; the per-first-action statistics (KartMCTS.processLeaf :124-159 + backpropagate :280-289)
// are reduced in shared memory and flushed with one atomic per block and statistic.
#include "hk_common.cuh"
#include "hk_game.cuh"
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

namespace hk {

// ---- track (DiscretePositionTracker.cs) -------------------------------------------------------------------------
__device__ __forceinline__ const hk_section& sec(const DevGame& g, int section) { return g.sections[section % g.n_sections]; }
__device__ __forceinline__ bool is_straight(const hk_section& s) { return s.insideR == 0.0f; }          // :197
__device__ __forceinline__ float lane_radius(const hk_section& s, int lane)                              // :72-88
{
    const int k = s.leftTurn ? (lane - 1) : (4 - lane);
    const float q = k == 1 ? (1.0f / 4.0f) : (k == 2 ? (2.0f / 4.0f) : (3.0f / 4.0f));
    return k == 0 ? s.insideR : s.insideR + s.width * q;
}
__device__ __forceinline__ float radius_of_lane(const hk_section& s, int a, int b)                       // :153-158
{
    if (is_straight(s)) return 0.0f;
    return (lane_radius(s, a) + lane_radius(s, b)) / 2.0f;
}
__device__ __forceinline__ float distance_to_travel(const hk_section& s, int a, int b)                   // :163-175
{
    if (is_straight(s)) {
        const float w = ((float)abs(a - b) * 1.0f / 3.0f) * s.width;
        return sqrtf(w * w + s.length * s.length);
    }
    return (3.14159274f / 180.0f) * s.turnDeg * radius_of_lane(s, a, b);
}
__device__ __forceinline__ float tire_load(const hk_section& s, float velocity, int a, int b)            // :180-192
{
    if (is_straight(s)) return distance_to_travel(s, a, b) * 0.01f;
    const float gs = (velocity * velocity) / radius_of_lane(s, a, b);
    return gs * distance_to_travel(s, a, b) * 0.01f;
}
__device__ __forceinline__ int optimal_lane_sign(const hk_section& s) { return s.optimalLane == 1 ? 1 : (s.optimalLane == 4 ? -1 : 0); }

// ---- kart (ArcadeKart.cs) -----------------------------------------------------------------------------------------
__device__ __forceinline__ float max_speed_radius_wear(const hk_kart& k, float radius, float wear)      // :536-547, :517-520
{
    if (radius == 0.0f) return k.topSpeed;
    const float gs = (1.0f - wear) * (k.maxGs - k.minGs) + k.minGs;
    float v = sqrtf(gs * 9.81f * fabsf(radius));
    if (isinf(v) || isnan(v)) v = k.topSpeed;
    if (v < 0.0001f) v = 0.0001f; else if (v > k.topSpeed) v = k.topSpeed;
    return v;
}

// C# (int)float on Mono/x64 is cvttss2si: NaN / out of range -> INT_MIN; CUDA would saturate / give 0
__device__ __forceinline__ int f2i(float f)
{
    if (!(f > -2147483904.0f && f < 2147483648.0f)) return INT_MIN;
    return __float2int_rz(f);
}
__device__ __forceinline__ float avg_velocity(int mn, int mx) { return (1.0f * (float)(mn + mx)) / 2.0f; }   // :58-61

__device__ float compute_toc(const hk_kart& k, float distance, float radius, float wear, float initV, float finalV)   // :67-122
{
    if (finalV > initV && (finalV * finalV - initV * initV) / (2.0f * k.accel) > distance) return -1.0f;
    if (initV > finalV && (initV * initV - finalV * finalV) / (2.0f * k.braking) > distance) return -1.0f;
    const float ms = max_speed_radius_wear(k, radius, wear);
    float t1, t3;
    if (ms >= initV) t1 = (ms - initV) / k.accel; else t1 = (initV - ms) / k.braking;
    if (ms >= finalV) t3 = (ms - finalV) / k.braking; else t3 = (finalV - ms) / k.accel;
    const float x1 = 0.5f * (initV + ms) * t1;
    const float x3 = 0.5f * (finalV + ms) * t3;
    const float x2 = distance - x1 - x3;
    const float t2 = x2 / ms;
    if ((double)t2 > 0.001) {
        return t1 + t2 + t3;
    } else if (initV <= ms) {
        const float maxSpeed = sqrtf((2.0f * distance * -k.braking * k.accel + -k.braking * initV * initV - k.accel * finalV * finalV)
                                     / (-k.accel - k.braking));
        t1 = (maxSpeed - initV) / k.accel;
        t3 = (maxSpeed - finalV) / k.braking;
        return t1 + t3;
    }
    return -1.0f;
}

__device__ hk_kart_state apply_action(const DevGame& g, const hk_kart_state& s, const hk_action& a)    // :127-171
{
    const hk_kart& kart = g.env_karts[s.player];                  // environment.Agents[player].m_Kart (:129, quirk B.6-2)
    hk_kart_state ns;
    ns.player = s.player; ns.team = s.team; ns.infeasible = 0;
    ns.section = s.section + 1;
    ns.min_velocity = a.min_velocity; ns.max_velocity = a.max_velocity; ns.lane = a.lane;
    const hk_section& cur = sec(g, s.section);
    if (is_straight(cur) != is_straight(sec(g, s.section + 1))) ns.laneChanges = 0;
    else if (ns.lane != s.lane) ns.laneChanges = s.laneChanges + abs(ns.lane - s.lane);
    else ns.laneChanges = s.laneChanges;
    const float dist = distance_to_travel(cur, s.lane, a.lane);
    const float rad = radius_of_lane(cur, s.lane, a.lane);
    // newState.tireAge is still 0 when computeTOC is called => wear 0 (quirk B.6-3)
    const float toc = compute_toc(kart, dist, rad, 0.0f / 10000.0f, avg_velocity(s.min_velocity, s.max_velocity),
                                  avg_velocity(a.min_velocity, a.max_velocity));
    const int timeUpdate = f2i(toc * (float)g.p.timePrecision);
    if (timeUpdate < 0) ns.infeasible = 1;
    ns.timeAtSection = (int)((unsigned)s.timeAtSection + (unsigned)timeUpdate);
    const float load = tire_load(cur, (float)a.max_velocity, s.lane, a.lane);
    ns.tireAge = f2i(((float)s.tireAge / 10000.0f + load * kart.tireWearFactor) * (float)10000);
    return ns;
}

// ---- DiscreteGameState ----------------------------------------------------------------------------------------------
__device__ __forceinline__ int cmp_kart(const hk_kart_state& a, const hk_kart_state& b)                 // :191-227
{
    if (a.section < b.section) return -1;
    if (a.section > b.section) return 1;
    if (a.timeAtSection < b.timeAtSection) return -1;
    if (a.timeAtSection == b.timeAtSection) {
        const float va = avg_velocity(a.min_velocity, a.max_velocity), vb = avg_velocity(b.min_velocity, b.max_velocity);
        if (va > vb) return -1;
        if (va == vb) return 0;
        return 1;
    }
    return 1;
}

// upNext (:188-243): first kart, in stable (section, time, -avgVelocity) order, that is not yet at lastCompleted+1.
// For 2 karts (one SwapIfGreater) and 4..16 karts (insertion sort) List.Sort is stable, which is equivalent to: among
// karts with section != lastCompleted+1, the minimum under cmp, lowest index on ties.
__device__ int up_next(const hk_game_state& st)
{
    if (st.n_karts == 3) {
        // List.Sort -> IntroSort special-cases 3 elements with a 3-exchange network that is NOT stable
        // (SwapIfGreater(0,1), (0,2), (1,2)); ties are common at the root (equal times and buckets), so replay it.
        int o0 = 0, o1 = 1, o2 = 2, t;
        if (cmp_kart(st.karts[o0], st.karts[o1]) > 0) { t = o0; o0 = o1; o1 = t; }
        if (cmp_kart(st.karts[o0], st.karts[o2]) > 0) { t = o0; o0 = o2; o2 = t; }
        if (cmp_kart(st.karts[o1], st.karts[o2]) > 0) { t = o1; o1 = o2; o2 = t; }
        if (st.karts[o0].section != st.lastCompletedSection + 1) return o0;
        if (st.karts[o1].section != st.lastCompletedSection + 1) return o1;
        if (st.karts[o2].section != st.lastCompletedSection + 1) return o2;
        return -1;
    }
    int best = -1;
    for (int i = 0; i < st.n_karts; ++i) {
        if (st.karts[i].section == st.lastCompletedSection + 1) continue;
        if (best < 0 || cmp_kart(st.karts[i], st.karts[best]) < 0) best = i;
    }
    return best;
}

// nextMoves (:322-415) fused with the rollout policy's sort keys (KartMCTS.cs:256).  For every legal candidate a 64-bit
// key = (dTime asc | max_velocity desc | |dLane| asc | optSign*lane asc | generation index) reproduces the stable
// OrderBy/ThenBy chain; keys[] is indexed by generation index, illegal candidates hold ~0.  Returns the legal count.
__device__ int legal_moves(const DevGame& g, const hk_game_state& st, int np, unsigned long long* keys)
{
    const hk_kart& kart = g.karts[np];
    const hk_kart_state& cs = st.karts[np];
    const hk_section& cur = sec(g, cs.section);
    const int optSign = optimal_lane_sign(g.sections[st.lastCompletedSection % g.n_sections]);            // KartMCTS.cs:252
    const bool straight = is_straight(cur);
    const float wear = (float)cs.tireAge / 10000.0f;
    int cnt = 0, gi = 0;
    for (int v = 6; v < g.vmax; v += g.p.velocityBucketSize) {
        const int vmx = min(v + g.p.velocityBucketSize, g.vmax);
        for (int lane = 1; lane < 5; ++lane, ++gi) {
            unsigned long long key = ~0ull;
            const int dl = abs(lane - cs.lane);
            bool ok = !(straight && cs.laneChanges + dl > g.p.maxLaneChanges);                            // :346
            if (ok) {
                const float radius = radius_of_lane(cur, cs.lane, lane);
                ok = !(max_speed_radius_wear(kart, radius, wear) < (float)v);                             // :357 (real wear)
            }
            if (ok) {
                hk_action a{v, vmx, lane};
                const hk_kart_state ap = apply_action(g, cs, a);                                          // :368
                ok = !ap.infeasible;
                if (ok) {
                    const unsigned dt = (unsigned)(ap.timeAtSection - cs.timeAtSection);                  // >= 0 when feasible
                    key = ((unsigned long long)dt << 22) | ((unsigned long long)(1023 - vmx) << 12) |
                          ((unsigned long long)dl << 10) | ((unsigned long long)(optSign * lane + 4) << 6) | (unsigned long long)gi;
                    ++cnt;
                }
            }
            keys[gi] = key;
        }
    }
    return cnt;
}

// One candidate of legal_moves (same filters, same key): lets a thread block evaluate the candidates of a node in parallel.
__device__ unsigned long long legal_move_key(const DevGame& g, const hk_game_state& st, int np, int gi)
{
    const hk_kart& kart = g.karts[np];
    const hk_kart_state& cs = st.karts[np];
    const hk_section& cur = sec(g, cs.section);
    const int optSign = optimal_lane_sign(g.sections[st.lastCompletedSection % g.n_sections]);            // KartMCTS.cs:252
    const float wear = (float)cs.tireAge / 10000.0f;
    const int v = 6 + (gi >> 2) * g.p.velocityBucketSize, vmx = min(v + g.p.velocityBucketSize, g.vmax), lane = (gi & 3) + 1;
    const int dl = abs(lane - cs.lane);
    if (is_straight(cur) && cs.laneChanges + dl > g.p.maxLaneChanges) return ~0ull;                        // :346
    if (max_speed_radius_wear(kart, radius_of_lane(cur, cs.lane, lane), wear) < (float)v) return ~0ull;    // :357 (real wear)
    const hk_kart_state ap = apply_action(g, cs, hk_action{v, vmx, lane});                                 // :368
    if (ap.infeasible) return ~0ull;
    const unsigned dt = (unsigned)(ap.timeAtSection - cs.timeAtSection);
    return ((unsigned long long)dt << 22) | ((unsigned long long)(1023 - vmx) << 12) | ((unsigned long long)dl << 10) |
           ((unsigned long long)(optSign * lane + 4) << 6) | (unsigned long long)gi;
}

// k-th smallest key (k < cnt): k+1 passes of "smallest key greater than the previous one" — keys are distinct.
__device__ int select_kth(const unsigned long long* keys, int n_cand, int k)
{
    unsigned long long prev = 0;
    bool first = true;
    int idx = 0;
    for (int pass = 0; pass <= k; ++pass) {
        unsigned long long best = ~0ull;
        for (int c = 0; c < n_cand; ++c) {
            const unsigned long long kk = keys[c];
            if ((first || kk > prev) && kk < best) { best = kk; idx = c; }
        }
        prev = best; first = false;
    }
    return idx;
}

__device__ hk_action action_of(const DevGame& g, int gi)
{
    const int v = 6 + (gi >> 2) * g.p.velocityBucketSize;
    return hk_action{v, min(v + g.p.velocityBucketSize, g.vmax), (gi & 3) + 1};
}

__device__ void make_move(const DevGame& g, hk_game_state& st, int np, const hk_action& a)              // :420-446
{
    const int last = st.lastCompletedSection;
    st.karts[np] = apply_action(g, st.karts[np], a);
    bool allAhead = true;
    for (int i = 0; i < st.n_karts; ++i) allAhead &= st.karts[i].section > last;
    if (allAhead) st.lastCompletedSection = last + 1;
}

// isOver (:251-317) given the legal-move count of the state. Returns over flag; scores[0..n_scores) as the reference list.
__device__ bool is_over(const DevGame& g, const hk_game_state& st, int n_moves, int np, float* scores, int& n_scores)
{
    n_scores = 0;
    if (n_moves == 0) {
        for (int i = 0; i < st.n_karts; ++i) {
            if (i == np || st.karts[i].team == st.karts[np].team) scores[n_scores++] = 0.0f;
            scores[n_scores++] = 0.5f;                                // no `else` (quirk B.6-4)
        }
        return true;
    }
    if (st.lastCompletedSection != st.finalSection) return false;
    if (st.n_karts > 1) {
        float maxScore = (float)g.p.timePrecision * -1000.0f;
        float minScore = (float)g.p.timePrecision * 1000.0f;
        float raw[HK_MAX_KARTS];
        float teamScore = 0.0f, opponentScore = 0.0f;                 // never reset (quirk B.6-5)
        int teamCount = 0, opponentCount = 0;
        for (int s = 0; s < st.n_karts; ++s) {
            for (int o = 0; o < st.n_karts; ++o) {
                if (s == o) teamScore += (float)st.karts[o].timeAtSection;
                else if (st.karts[s].team == st.karts[o].team) {
                    teamScore += (float)st.karts[o].timeAtSection * g.p.teamScoreRewardMultiplier;
                    teamCount += 1;
                } else {
                    opponentScore += (float)st.karts[o].timeAtSection;
                    opponentCount += 1;
                }
            }
            const float score = opponentScore * (((float)teamCount * g.p.teamScoreRewardMultiplier + 1.0f) / ((float)opponentCount * 1.0f)) - teamScore;
            raw[s] = score;
            maxScore = maxScore > score ? maxScore : score;
            minScore = minScore < score ? minScore : score;
        }
        for (int s = 0; s < st.n_karts; ++s) {
            const int sc = f2i(raw[s]);                               // foreach (int score in scores) (quirk B.6-6)
            scores[n_scores++] = ((float)sc - minScore) * 1.0f / (maxScore - minScore);
        }
        return true;
    }
    scores[n_scores++] = (float)(g.p.maxEpisodeSteps - st.karts[0].timeAtSection / g.p.maxEpisodeSteps);   // :314
    return true;
}

// ---- Philox4x32-10, counter = (rollout lo, rollout hi, ply, 0), key = seed ------------------------------------------
__device__ __forceinline__ unsigned philox_first(unsigned long long seed, unsigned long long rollout, unsigned ply)
{
    unsigned c0 = (unsigned)rollout, c1 = (unsigned)(rollout >> 32), c2 = ply, c3 = 0u;
    unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}

__device__ __forceinline__ int policy_index(const DevGame& g, int cnt, unsigned u)
{
    // number of thresholds <= u among the cnt - 1 ascending ones: branch-free binary search (6 probes instead of up to 35)
    const unsigned* cdf = g.cdf[cnt];
    const int m = cnt - 1;
    int idx = 0;
#pragma unroll
    for (int s = 32; s; s >>= 1) {
        const int k = idx + s;
        if (k <= m && cdf[k - 1] <= u) idx = k;
    }
    return idx;
}

// ---- table-driven ply (same results as legal_moves + select_kth + make_move, see DevGame) ---------------------------------
struct Tables {
    const int* dt; const unsigned char* order; const float* load; const float* radius; const unsigned long long* lmask;
    const unsigned long long* od;
    int nv, nc;
    __device__ Tables(const DevGame& g)
        : dt(reinterpret_cast<const int*>(g.tables + g.off_dt)), order(g.tables + g.off_order),
          load(reinterpret_cast<const float*>(g.tables + g.off_load)), radius(reinterpret_cast<const float*>(g.tables + g.off_radius)),
          lmask(reinterpret_cast<const unsigned long long*>(g.tables + g.off_lmask)),
          od(reinterpret_cast<const unsigned long long*>(g.tables + g.off_od)), nv(g.nv), nc(g.n_cand) {}
};

// x / velocityBucketSize for x >= 0 without the ~20-instruction integer division in the two bucket sizes the reference uses
__device__ __forceinline__ int div_bucket(int x, int b) { return b == 1 ? x : b == 2 ? (x >> 1) : x / b; }

// velocity level of a kart's bucket if it is one of the action buckets (KartDiscreteGame.cs:329-340), else -1 (e.g. the root's (0, b))
__device__ __forceinline__ int velocity_level(const DevGame& g, int mn, int mx)
{
    const int b = g.p.velocityBucketSize, j = div_bucket(mn - 6, b);
    return (mn >= 6 && mn < g.vmax && (mn - 6) == j * b && mx == min(mn + b, g.vmax)) ? j : -1;
}

// Legal moves of kart np as a bit mask over the RANKS of the pre-sorted candidate list; returns the count.
// sidx (out) = the kart's section index on the track; lcs_idx = lastCompletedSection % n_sections, carried by the caller (an integer
// modulo by a run-time divisor is ~15 instructions, and a ply used to take three)
__device__ __forceinline__ int fast_legal(const DevGame& g, const Tables& tb, const hk_game_state& st, int np, int lvl,
                                          unsigned long long& mask, const unsigned char*& ord, int& sidx, int lcs_idx)
{
    const hk_kart_state& cs = st.karts[np];
    sidx = cs.section % g.n_sections;
    const int type = g.type_of[sidx], l0 = cs.lane - 1;
    const int os = (g.sec_flags[lcs_idx] >> 2) & 3;                                                  // KartMCTS.cs:252
    const float wear = (float)cs.tireAge / 10000.0f;
    const int maxdl = (g.sec_flags[sidx] & 1) ? g.p.maxLaneChanges - cs.laneChanges : 99;             // :346
    const size_t cell = (((size_t)type * 4 + l0) * tb.nv + lvl) * 3 + os;
    ord = tb.order + cell * tb.nc;
    const unsigned long long* lm = tb.lmask + cell * 4 * tb.nv;
    const int b = g.p.velocityBucketSize;
    // Per target lane the speed filter !(maxSpeed < v) (:357) passes the velocity levels up to floor(maxSpeed) (v is an integer;
    // a NaN limit passes all of them), and the statically feasible moves into that lane up to a level are one precomputed mask
    // over the policy's rank order — no scan over the 4 nv candidates.  Branch-free: the four mask loads are issued together.
    unsigned long long mv[4];
    bool okl[4];
#pragma unroll
    for (int l1 = 0; l1 < 4; ++l1) {
        const float ms = max_speed_radius_wear(g.karts[np], g.radius_tab[type * 16 + l0 * 4 + l1], wear);
        const int mi = (int)fminf(ms, 1000.0f);                                                         // fminf(NaN, x) = x
        okl[l1] = abs(l1 - l0) <= maxdl && mi >= 6;
        const int jm = min(div_bucket(max(mi, 6) - 6, b), tb.nv - 1);
        mv[l1] = __ldg(&lm[l1 * tb.nv + jm]);
    }
    mask = 0ull;
#pragma unroll
    for (int l1 = 0; l1 < 4; ++l1) mask |= okl[l1] ? mv[l1] : 0ull;
    return __popcll(mask);
}

__device__ __forceinline__ int nth_set_bit(unsigned long long mask, int n)
{
    // position of the n-th (0-based) set bit: half selection, then a 5-step popcount descent (__fns is a ~50-instruction loop)
    const unsigned lo = (unsigned)mask, hi = (unsigned)(mask >> 32);
    const int clo = __popc(lo);
    const bool up = n >= clo;
    unsigned w = up ? hi : lo;
    int pos = up ? 32 : 0;
    n -= up ? clo : 0;
#pragma unroll
    for (int s = 16; s; s >>= 1) {
        const int c = __popc(w & ((1u << s) - 1u));
        const bool u2 = n >= c;
        n -= u2 ? c : 0; pos += u2 ? s : 0; w = u2 ? (w >> s) : w;
    }
    return pos;
}

// applyAction + makeMove bookkeeping from the tables (:127-171, :420-446)
__device__ __forceinline__ void fast_move(const DevGame& g, const Tables& tb, hk_game_state& st, int np, int lvl, int gi, int sidx, int& lcs_idx)
{
    hk_kart_state& cs = st.karts[np];
    const int type = g.type_of[sidx], l0 = cs.lane - 1, l1 = gi & 3, j = gi >> 2;
    const int b = g.p.velocityBucketSize, v = 6 + j * b;
    const int dtv = __ldg(&tb.dt[(((size_t)type * 4 + l0) * tb.nv + lvl) * tb.nc + gi]);
    const float load = __ldg(&tb.load[((size_t)type * 16 + l0 * 4 + l1) * tb.nv + j]);
    const int last = st.lastCompletedSection;
    if (g.sec_flags[sidx] & 2) cs.laneChanges = 0; else cs.laneChanges += abs(l1 - l0);
    cs.tireAge = f2i(((float)cs.tireAge / 10000.0f + load * g.env_karts[0].tireWearFactor) * (float)10000);
    cs.timeAtSection = (int)((unsigned)cs.timeAtSection + (unsigned)dtv);
    cs.section += 1; cs.min_velocity = v; cs.max_velocity = min(v + b, g.vmax); cs.lane = l1 + 1; cs.infeasible = 0;
    bool allAhead = true;
    for (int i = 0; i < st.n_karts; ++i) allAhead &= st.karts[i].section > last;
    if (allAhead) { st.lastCompletedSection = last + 1; lcs_idx = lcs_idx + 1 == g.n_sections ? 0 : lcs_idx + 1; }
}

// One thread per (type, lane, velocity level): time updates of all candidates, then the policy's static sort order for the
// three possible optimal-lane signs; threads with lvl == 0 also fill the load / radius tables of their (type, lane).
__global__ void build_tables_kernel(DevGame* __restrict__ gg, unsigned char* blob)
{
    const DevGame& g = *gg;
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= g.n_types * 4 * g.nv) return;
    const int lvl = id % g.nv, l0 = (id / g.nv) % 4, type = id / (g.nv * 4);
    int* dt = reinterpret_cast<int*>(blob + g.off_dt) + (size_t)id * g.n_cand;
    unsigned char* order = blob + g.off_order + (size_t)id * 3 * g.n_cand;
    const int b = g.p.velocityBucketSize, sec0 = g.rep_section[type];
    hk_kart_state cs{};
    cs.player = 0; cs.section = sec0; cs.lane = l0 + 1; cs.tireAge = 0;
    cs.min_velocity = 6 + lvl * b; cs.max_velocity = min(cs.min_velocity + b, g.vmax);
    unsigned long long keys[3][HK_MAX_ACTIONS];
    for (int gi = 0; gi < g.n_cand; ++gi) {
        const hk_action a = hk_action{6 + (gi >> 2) * b, min(6 + (gi >> 2) * b + b, g.vmax), (gi & 3) + 1};
        const hk_kart_state ap = apply_action(g, cs, a);
        const int dtv = ap.timeAtSection - cs.timeAtSection;
        dt[gi] = ap.infeasible ? -1 : dtv;
        const int dl = abs(a.lane - cs.lane);
        for (int os = 0; os < 3; ++os)
            keys[os][gi] = ap.infeasible ? ~0ull
                : ((unsigned long long)(unsigned)dtv << 22) | ((unsigned long long)(1023 - a.max_velocity) << 12) |
                  ((unsigned long long)dl << 10) | ((unsigned long long)((os - 1) * a.lane + 4) << 6) | (unsigned long long)gi;
    }
    for (int os = 0; os < 3; ++os) {
        int n_ok = 0;
        for (int gi = 0; gi < g.n_cand; ++gi) n_ok += keys[os][gi] != ~0ull;
        for (int r = 0; r < g.n_cand; ++r) order[os * g.n_cand + r] = r < n_ok ? (unsigned char)select_kth(keys[os], g.n_cand, r) : 255;
        // rank -> (time update, generation index) in one 8-byte entry: the packed playout's move costs one dependent load, not two
        unsigned long long* od = reinterpret_cast<unsigned long long*>(blob + g.off_od) + ((size_t)id * 3 + os) * g.n_cand;
        for (int r = 0; r < g.n_cand; ++r) {
            const int gi = order[os * g.n_cand + r];
            od[r] = gi == 255 ? 0xFFull : (((unsigned long long)(unsigned)dt[gi]) << 8) | (unsigned long long)gi;
        }
    }
    {   // masks over the rank order: moves into lane l1 with velocity level <= j, per optimal-lane sign
        unsigned long long* lm = reinterpret_cast<unsigned long long*>(blob + g.off_lmask) + (size_t)id * 3 * 4 * g.nv;
        for (int os = 0; os < 3; ++os)
            for (int l1 = 0; l1 < 4; ++l1)
                for (int j = 0; j < g.nv; ++j) {
                    unsigned long long m = 0;
                    for (int r = 0; r < g.n_cand; ++r) {
                        const int gi = order[os * g.n_cand + r];
                        if (gi != 255 && (gi & 3) == l1 && (gi >> 2) <= j) m |= 1ull << r;
                    }
                    lm[(os * 4 + l1) * g.nv + j] = m;
                }
    }
    if (lvl == 0) {
        float* load = reinterpret_cast<float*>(blob + g.off_load);
        float* radius = reinterpret_cast<float*>(blob + g.off_radius);
        const hk_section& cur = g.sections[sec0];
        for (int l1 = 0; l1 < 4; ++l1) {
            radius[type * 16 + l0 * 4 + l1] = radius_of_lane(cur, l0 + 1, l1 + 1);
            gg->radius_tab[type * 16 + l0 * 4 + l1] = radius[type * 16 + l0 * 4 + l1];
            for (int j = 0; j < g.nv; ++j)
                load[((size_t)type * 16 + l0 * 4 + l1) * g.nv + j] = tire_load(cur, (float)min(6 + j * b + b, g.vmax), l0 + 1, l1 + 1);
        }
    }
}

// One playout of KartMCTS.simulate (:238-278). Returns plies played; first_gi = generation index of the first action.
// Legal moves of the leaf's own karts in policy order.  A kart that has not moved yet in a rollout still has its leaf state, and
// its legal set depends on nothing else (the collision filter is disabled, KartDiscreteGame.cs:398; the optimal-lane sign is
// that of lastCompletedSection, which only advances once every kart has moved): the sorted list of the slow, directly
// evaluated plies — the root's (0, bucket) velocity bucket is not an action bucket, quirk B.6-1 — is the same for every
// rollout of a leaf, so it is built once per block.
struct RootMoves {
    int cnt[HK_MAX_KARTS];                               // -1: not prepared (table-driven kart)
    unsigned char order[HK_MAX_KARTS][HK_MAX_ACTIONS];   // generation indices in policy order
};

// ---- packed two-kart playout ---------------------------------------------------------------------------------------------------
// Once both karts of a 2-kart game hold action velocity buckets every remaining ply is table-driven, and the state fits in registers:
// per kart (section, timeAtSection, tireAge, misc) with misc = lane-1 | section index on the track << 2 | velocity level << 8 |
// laneChanges << 12.  No local-memory state, no modulo per ply (the section index is carried), upNext() is two compares (at action
// buckets the average velocity is monotone in the level).  Same transitions, draws and terminal scores as the struct-based loop: it
// unpacks into the hk_game_state and lets is_over() score the end.
struct K2 { int sec, time, tire, misc; };

__device__ __forceinline__ bool pack_k2(const DevGame& g, const hk_kart_state& k, K2& o)
{
    const int lvl = k.player == 0 ? velocity_level(g, k.min_velocity, k.max_velocity) : -1;
    if (lvl < 0 || (unsigned)k.laneChanges >= (1u << 19) || (unsigned)(k.lane - 1) > 3u) return false;
    o.sec = k.section; o.time = k.timeAtSection; o.tire = k.tireAge;
    o.misc = (k.lane - 1) | ((k.section % g.n_sections) << 2) | (lvl << 8) | (k.laneChanges << 12);
    return true;
}

__device__ __forceinline__ void unpack_k2(const DevGame& g, const K2& k, hk_kart_state& o)
{
    const int b = g.p.velocityBucketSize, lvl = (k.misc >> 8) & 15, v = 6 + lvl * b;
    o.section = k.sec; o.timeAtSection = k.time; o.tireAge = k.tire; o.lane = (k.misc & 3) + 1; o.laneChanges = k.misc >> 12;
    o.min_velocity = v; o.max_velocity = min(v + b, g.vmax); o.infeasible = 0;
}

// continues a playout from `ply` with both karts packed; returns the total number of plies (or -1: upNext() == -1).
// REC (the sequential search's playouts): every ply also leaves its record word gi | nextMoves().Count << 8 | (upNext() after the move) << 16 in
// rec_plies[ply] — the last field is known one trip later, so a word is written when the next trip has found who moves; -2: more than cap plies.
template <bool REC = false>
__device__ int rollout_packed2(const DevGame& g, const Tables& tb, hk_game_state& st, K2 k0, K2 k1, unsigned long long seed, unsigned long long rid,
                               float* scores, int& n_scores, int& first_gi, int ply, unsigned* rec_plies = nullptr, int cap = 0)
{
    int last = st.lastCompletedSection, lcs_idx = last % g.n_sections;
    const int fin = st.finalSection, b = g.p.velocityBucketSize;
    unsigned pending = 0;
    bool have_pending = false;
    for (;;) {
        // upNext (:188-243): karts not yet at last + 1, minimum (section, time, -avgVelocity), lowest index on ties
        const bool e0 = k0.sec != last + 1, e1 = k1.sec != last + 1;
        if (!e0 && !e1) {
            if (REC && have_pending) rec_plies[ply - 1] = pending | (0xffu << 16);
            return -1;
        }
        const bool one_first = k1.sec < k0.sec || (k1.sec == k0.sec && (k1.time < k0.time || (k1.time == k0.time && ((k1.misc >> 8) & 15) > ((k0.misc >> 8) & 15))));
        const int np = (e0 && e1) ? (one_first ? 1 : 0) : (e0 ? 0 : 1);
        if (REC && have_pending) rec_plies[ply - 1] = pending | ((unsigned)np << 16);
        const K2 k = np ? k1 : k0;
        const int l0 = k.misc & 3, sidx = (k.misc >> 2) & 63, lvl = (k.misc >> 8) & 15, lc = k.misc >> 12;
        const int type = g.type_of[sidx], flags = g.sec_flags[sidx];
        const int os = (g.sec_flags[lcs_idx] >> 2) & 3;                                                   // KartMCTS.cs:252
        const float wear = (float)k.tire / 10000.0f;
        const int maxdl = (flags & 1) ? g.p.maxLaneChanges - lc : 99;                                     // :346
        const size_t cell = (((size_t)type * 4 + l0) * tb.nv + lvl) * 3 + os;
        const unsigned long long* od = tb.od + cell * tb.nc;
        const unsigned long long* lm = tb.lmask + cell * 4 * tb.nv;
        // the four target lanes without branches: the mask loads are issued together (always at a valid index) and selected afterwards
        unsigned long long mv[4];
        bool okl[4];
#pragma unroll
        for (int l1 = 0; l1 < 4; ++l1) {
            const float ms = max_speed_radius_wear(g.karts[np], g.radius_tab[type * 16 + l0 * 4 + l1], wear);
            const int mi = (int)fminf(ms, 1000.0f);                                                     // fminf(NaN, x) = x
            okl[l1] = abs(l1 - l0) <= maxdl && mi >= 6;
            const int jm = min(div_bucket(max(mi, 6) - 6, b), tb.nv - 1);
            mv[l1] = __ldg(&lm[l1 * tb.nv + jm]);
        }
        unsigned long long mask = 0ull;
#pragma unroll
        for (int l1 = 0; l1 < 4; ++l1) mask |= okl[l1] ? mv[l1] : 0ull;
        const int cnt = __popcll(mask);
        if (cnt == 0 || last == fin) {                                                                   // isOver (:251-317) on the struct
            unpack_k2(g, k0, st.karts[0]); unpack_k2(g, k1, st.karts[1]);
            st.lastCompletedSection = last;
            if (is_over(g, st, cnt, np, scores, n_scores)) break;
        }
        const int index = policy_index(g, cnt, philox_first(seed, rid, (unsigned)ply));
        const unsigned long long mvrec = __ldg(&od[nth_set_bit(mask, index)]);
        const int gi = (int)(mvrec & 0xFFull);
        if (REC) {
            if (ply >= cap) return -2;
            pending = (unsigned)gi | ((unsigned)cnt << 8);
            have_pending = true;
        }
        // applyAction + makeMove from the tables (:127-171, :420-446)
        const int l1 = gi & 3, j = gi >> 2;
        const int dtv = (int)(unsigned)(mvrec >> 8);
        const float load = __ldg(&tb.load[((size_t)type * 16 + l0 * 4 + l1) * tb.nv + j]);
        K2 nk;
        const int nlc = (flags & 2) ? 0 : lc + abs(l1 - l0);
        const int nsidx = sidx + 1 == g.n_sections ? 0 : sidx + 1;
        nk.tire = f2i(((float)k.tire / 10000.0f + load * g.env_karts[0].tireWearFactor) * (float)10000);
        nk.time = (int)((unsigned)k.time + (unsigned)dtv);
        nk.sec = k.sec + 1;
        nk.misc = l1 | (nsidx << 2) | (j << 8) | (nlc << 12);
        if (np) k1 = nk; else k0 = nk;
        if (k0.sec > last && k1.sec > last) { last += 1; lcs_idx = lcs_idx + 1 == g.n_sections ? 0 : lcs_idx + 1; }
        if (ply == 0) first_gi = gi;
        ++ply;
    }
    return ply;
}

// The same for games of three and four karts (Duos: BASELINE config 3's game inside the search): NK packed karts in registers, the
// moving kart picked and written back by select chains (an indexed array would live in local memory).  upNext (:188-243) as up_next()
// has it: for four karts the minimum under (section, time, -avgVelocity) among the karts not yet at last + 1, lowest index on ties; for
// three the unstable 3-exchange network of List.Sort.  Used by the sequential search's playouts (seq_playouts_kernel<4>).
template <int NK, bool REC>
__device__ int rollout_packedN(const DevGame& g, const Tables& tb, hk_game_state& st, K2 k0, K2 k1, K2 k2, K2 k3, unsigned long long seed,
                               unsigned long long rid, float* scores, int& n_scores, int ply, unsigned* rec_plies, int cap)
{
    static_assert(NK == 3 || NK == 4, "three or four karts");
    int last = st.lastCompletedSection, lcs_idx = last % g.n_sections;
    const int fin = st.finalSection, b = g.p.velocityBucketSize;
    unsigned pending = 0;
    bool have_pending = false;
    auto cmp = [](const K2& a, const K2& c) -> int {                  // cmp_kart on packed karts (the average velocity is monotone in the level)
        if (a.sec != c.sec) return a.sec < c.sec ? -1 : 1;
        if (a.time != c.time) return a.time < c.time ? -1 : 1;
        const int la = (a.misc >> 8) & 15, lc = (c.misc >> 8) & 15;
        return la > lc ? -1 : (la == lc ? 0 : 1);
    };
    for (;;) {
        int np = -1;
        if (NK == 3) {
            int o0 = 0, o1 = 1, o2 = 2, t;
            auto kk = [&](int i) -> const K2& { return i == 0 ? k0 : i == 1 ? k1 : k2; };
            if (cmp(kk(o0), kk(o1)) > 0) { t = o0; o0 = o1; o1 = t; }
            if (cmp(kk(o0), kk(o2)) > 0) { t = o0; o0 = o2; o2 = t; }
            if (cmp(kk(o1), kk(o2)) > 0) { t = o1; o1 = o2; o2 = t; }
            if (kk(o0).sec != last + 1) np = o0;
            else if (kk(o1).sec != last + 1) np = o1;
            else if (kk(o2).sec != last + 1) np = o2;
        } else {
            K2 best = k0;
            if (k0.sec != last + 1) np = 0;
            if (k1.sec != last + 1 && (np < 0 || cmp(k1, best) < 0)) { np = 1; best = k1; }
            if (k2.sec != last + 1 && (np < 0 || cmp(k2, best) < 0)) { np = 2; best = k2; }
            if (k3.sec != last + 1 && (np < 0 || cmp(k3, best) < 0)) { np = 3; best = k3; }
        }
        if (REC && have_pending) rec_plies[ply - 1] = pending | (((unsigned)np & 0xffu) << 16);
        if (np < 0) return -1;
        const K2 k = np == 0 ? k0 : np == 1 ? k1 : np == 2 ? k2 : k3;
        const int l0 = k.misc & 3, sidx = (k.misc >> 2) & 63, lvl = (k.misc >> 8) & 15, lc = k.misc >> 12;
        const int type = g.type_of[sidx], flags = g.sec_flags[sidx];
        const int os = (g.sec_flags[lcs_idx] >> 2) & 3;                                                   // KartMCTS.cs:252
        const float wear = (float)k.tire / 10000.0f;
        const int maxdl = (flags & 1) ? g.p.maxLaneChanges - lc : 99;                                     // :346
        const size_t cell = (((size_t)type * 4 + l0) * tb.nv + lvl) * 3 + os;
        const unsigned long long* od = tb.od + cell * tb.nc;
        const unsigned long long* lm = tb.lmask + cell * 4 * tb.nv;
        unsigned long long mv[4];
        bool okl[4];
#pragma unroll
        for (int l1 = 0; l1 < 4; ++l1) {
            const float ms = max_speed_radius_wear(g.karts[np], g.radius_tab[type * 16 + l0 * 4 + l1], wear);
            const int mi = (int)fminf(ms, 1000.0f);                                                     // fminf(NaN, x) = x
            okl[l1] = abs(l1 - l0) <= maxdl && mi >= 6;
            const int jm = min(div_bucket(max(mi, 6) - 6, b), tb.nv - 1);
            mv[l1] = __ldg(&lm[l1 * tb.nv + jm]);
        }
        unsigned long long mask = 0ull;
#pragma unroll
        for (int l1 = 0; l1 < 4; ++l1) mask |= okl[l1] ? mv[l1] : 0ull;
        const int cnt = __popcll(mask);
        if (cnt == 0 || last == fin) {                                                                   // isOver (:251-317) on the struct
            unpack_k2(g, k0, st.karts[0]); unpack_k2(g, k1, st.karts[1]); unpack_k2(g, k2, st.karts[2]);
            if (NK == 4) unpack_k2(g, k3, st.karts[3]);
            st.lastCompletedSection = last;
            if (is_over(g, st, cnt, np, scores, n_scores)) break;
        }
        const int index = policy_index(g, cnt, philox_first(seed, rid, (unsigned)ply));
        const unsigned long long mvrec = __ldg(&od[nth_set_bit(mask, index)]);
        const int gi = (int)(mvrec & 0xFFull);
        if (REC) {
            if (ply >= cap) return -2;
            pending = (unsigned)gi | ((unsigned)cnt << 8);
            have_pending = true;
        }
        const int l1 = gi & 3, j = gi >> 2;
        const int dtv = (int)(unsigned)(mvrec >> 8);
        const float load = __ldg(&tb.load[((size_t)type * 16 + l0 * 4 + l1) * tb.nv + j]);
        K2 nk;
        const int nlc = (flags & 2) ? 0 : lc + abs(l1 - l0);
        const int nsidx = sidx + 1 == g.n_sections ? 0 : sidx + 1;
        nk.tire = f2i(((float)k.tire / 10000.0f + load * g.env_karts[0].tireWearFactor) * (float)10000);
        nk.time = (int)((unsigned)k.time + (unsigned)dtv);
        nk.sec = k.sec + 1;
        nk.misc = l1 | (nsidx << 2) | (j << 8) | (nlc << 12);
        if (np == 0) k0 = nk; else if (np == 1) k1 = nk; else if (np == 2) k2 = nk; else k3 = nk;
        if (k0.sec > last && k1.sec > last && k2.sec > last && (NK == 3 || k3.sec > last)) { last += 1; lcs_idx = lcs_idx + 1 == g.n_sections ? 0 : lcs_idx + 1; }
        ++ply;
    }
    return ply;
}

template <bool TRACE>
__device__ int rollout(const DevGame& g, hk_game_state st, unsigned long long seed, unsigned long long rid, float* scores,
                       int& n_scores, int& first_gi, hk_action* act_out, int* choice_out, const RootMoves* root = nullptr)
{
    const Tables tb(g);
    int ply = 0;
    unsigned moved = 0;
    first_gi = -1;
    n_scores = 0;
    int lcs_idx = st.lastCompletedSection % g.n_sections;
    for (;;) {
        if (!TRACE && st.n_karts == 2 && g.tables_ok) {                  // both karts at action buckets: the rest runs packed in registers
            K2 k0, k1;
            if (pack_k2(g, st.karts[0], k0) && pack_k2(g, st.karts[1], k1))
                return rollout_packed2(g, tb, st, k0, k1, seed, rid, scores, n_scores, first_gi, ply);
        }
        const int np = up_next(st);
        if (np < 0) return -1;
        const int lvl = (g.tables_ok && st.karts[np].player == 0) ? velocity_level(g, st.karts[np].min_velocity, st.karts[np].max_velocity) : -1;
        int gi, index;
        if (lvl < 0 && root && !((moved >> np) & 1u) && root->cnt[np] >= 0) {     // leaf kart, list prepared by the block
            const int cnt = root->cnt[np];
            if (is_over(g, st, cnt, np, scores, n_scores)) break;
            index = policy_index(g, cnt, philox_first(seed, rid, (unsigned)ply));
            gi = root->order[np][index];
            make_move(g, st, np, action_of(g, gi));
            lcs_idx = st.lastCompletedSection % g.n_sections;
        } else if (lvl >= 0) {                           // table-driven ply
            unsigned long long mask;
            const unsigned char* ord;
            int sidx;
            const int cnt = fast_legal(g, tb, st, np, lvl, mask, ord, sidx, lcs_idx);
            if (is_over(g, st, cnt, np, scores, n_scores)) break;
            index = policy_index(g, cnt, philox_first(seed, rid, (unsigned)ply));
            gi = __ldg(&ord[nth_set_bit(mask, index)]);
            fast_move(g, tb, st, np, lvl, gi, sidx, lcs_idx);
        } else {                                         // direct evaluation (e.g. the root's (0, bucket) velocity bucket)
            unsigned long long keys[HK_MAX_ACTIONS];
            const int cnt = legal_moves(g, st, np, keys);
            if (is_over(g, st, cnt, np, scores, n_scores)) break;
            index = policy_index(g, cnt, philox_first(seed, rid, (unsigned)ply));
            gi = select_kth(keys, g.n_cand, index);
            make_move(g, st, np, action_of(g, gi));
            lcs_idx = st.lastCompletedSection % g.n_sections;
        }
        moved |= 1u << np;
        if (ply == 0) first_gi = gi;
        if (TRACE && ply < HK_MAX_PLIES) { act_out[ply] = action_of(g, gi); choice_out[ply] = index; }
        ++ply;
    }
    return ply;
}

// ---- kernels ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 6) rollouts_kernel(const DevGame* __restrict__ gg, const hk_game_state* __restrict__ leaves,
                                                       long long rollouts_per_leaf, unsigned long long seed,
                                                       unsigned long long rollout_offset, unsigned long long* visit,
                                                       double* reward_sum, unsigned long long* nan_count,
                                                       unsigned long long* plies_sum, int* error_flag, unsigned long long* next)
{
    __shared__ DevGame g;
    __shared__ unsigned s_visit[HK_MAX_ACTIONS];
    __shared__ unsigned s_nan[HK_MAX_ACTIONS];
    __shared__ double s_reward[HK_MAX_ACTIONS][HK_MAX_KARTS];
    __shared__ unsigned s_plies;
    {
        const int* src = reinterpret_cast<const int*>(gg);
        int* dst = reinterpret_cast<int*>(&g);
        for (int i = threadIdx.x; i < (int)(sizeof(DevGame) / 4); i += blockDim.x) dst[i] = src[i];
        for (int i = threadIdx.x; i < HK_MAX_ACTIONS; i += blockDim.x) {
            s_visit[i] = 0; s_nan[i] = 0;
            for (int k = 0; k < HK_MAX_KARTS; ++k) s_reward[i][k] = 0.0;
        }
        if (threadIdx.x == 0) s_plies = 0;
    }
    __syncthreads();
    const int leaf = blockIdx.y;
    const hk_game_state st = leaves[leaf];
    __shared__ RootMoves s_root;
    __shared__ unsigned long long s_keys[HK_MAX_KARTS][HK_MAX_ACTIONS];
    if (threadIdx.x < HK_MAX_KARTS) {
        const int k = threadIdx.x;
        int cnt = -1;
        if (k < st.n_karts) {
            const int lvl = (g.tables_ok && st.karts[k].player == 0) ? velocity_level(g, st.karts[k].min_velocity, st.karts[k].max_velocity) : -1;
            if (lvl < 0) cnt = legal_moves(g, st, k, s_keys[k]);
        }
        s_root.cnt[k] = cnt;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < HK_MAX_KARTS * HK_MAX_ACTIONS; e += blockDim.x) {
        const int k = e / HK_MAX_ACTIONS, c = e % HK_MAX_ACTIONS;
        if (s_root.cnt[k] > 0 && c < g.n_cand && s_keys[k][c] != ~0ull) {       // keys are distinct: rank = number of smaller keys
            int rank = 0;
            for (int j = 0; j < g.n_cand; ++j) rank += s_keys[k][j] < s_keys[k][c];
            s_root.order[k][rank] = (unsigned char)c;
        }
    }
    __syncthreads();
    // Warps take the leaf's rollouts 32 at a time from a device counter (SMs differ by a few per cent in speed and a launch ends with
    // its slowest warp); which thread plays a rollout does not matter, its Philox counter is its id.  Block-local partial sums in double.
    for (;;) {
        unsigned long long base = 0;
        if ((threadIdx.x & 31) == 0) base = atomicAdd(&next[leaf], 32ull);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= (unsigned long long)rollouts_per_leaf) break;
        const long long r = (long long)base + (threadIdx.x & 31);
        if (r >= rollouts_per_leaf) continue;
        float scores[2 * HK_MAX_KARTS];
        int n_scores, first_gi;
        const unsigned long long rid = rollout_offset + (unsigned long long)leaf * (unsigned long long)rollouts_per_leaf + (unsigned long long)r;
        const int plies = rollout<false>(g, st, seed, rid, scores, n_scores, first_gi, nullptr, nullptr, &s_root);
        if (plies < 0) { atomicExch(error_flag, 1); continue; }
        if (plies == 0) continue;
        atomicAdd(&s_plies, (unsigned)plies);
        atomicAdd(&s_visit[first_gi], 1u);
        bool has_nan = false;
        for (int k = 0; k < st.n_karts && k < n_scores; ++k) has_nan |= isnan(scores[k]);
        if (has_nan) { atomicAdd(&s_nan[first_gi], 1u); continue; }
        for (int k = 0; k < st.n_karts && k < n_scores; ++k)
            if (scores[k] != 0.0f) atomicAdd(&s_reward[first_gi][k], (double)scores[k]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HK_MAX_ACTIONS; i += blockDim.x) {
        if (s_visit[i]) atomicAdd(&visit[(size_t)leaf * HK_MAX_ACTIONS + i], (unsigned long long)s_visit[i]);
        if (s_nan[i]) atomicAdd(&nan_count[(size_t)leaf * HK_MAX_ACTIONS + i], (unsigned long long)s_nan[i]);
        for (int k = 0; k < HK_MAX_KARTS; ++k)
            if (s_reward[i][k] != 0.0) atomicAdd(&reward_sum[((size_t)leaf * HK_MAX_ACTIONS + i) * HK_MAX_KARTS + k], s_reward[i][k]);
    }
    if (threadIdx.x == 0 && s_plies) atomicAdd(&plies_sum[leaf], (unsigned long long)s_plies);
}

// ---- batched tree search: KartMCTS.constructSearchTree + getBestStatesSequence on the device, one CTA per root ------------------
// The GPU form of the reference's own leaf-parallel variant (processLeaf, KartMCTS.cs:124-159), exactly as the host mirrors run it
// (hierarchicalkarting_b200/mcts.py, host/KartMCTS.hpp): every iteration walks the tree with upperConfidenceStrategy (:167-192;
// UCT = total/n + ln(parent.n / n) with the integer division of :164) to a node without children (findLeaf :194-201), creates
// ALL its children (generation order of nextMoves :329-340), plays R rollouts of simulate (:238-278) through each child and
// backpropagates the summed terminal scores (:280-289: every ancestor adds the entry of its own upNext()).  A whole tree lives in
// a per-root slab of global memory; thread 0 does the tree policy, all threads the child creation and the rollouts.  The random
// initial pick of upperConfidenceStrategy (it only matters for exact ties) comes from a Philox counter, so that a host mirror
// given the same stream reproduces the search.
struct TreeNode {
    hk_game_state st;
    double total;                                        // totalValue
    int episodes;                                        // numEpisodes
    int parent, first_child, n_children, upnext, pad_;
};

__device__ __forceinline__ float uct_weight(const TreeNode& parent, const TreeNode& c)                  // KartMCTS.cs:162-165
{
    const int ratio = parent.episodes / c.episodes;      // integer division (quirk B.6-7); the caller has checked c.episodes != 0
    const float lg = ratio > 0 ? (float)log((double)ratio) : -INFINITY;
    return (float)c.total / (float)c.episodes + lg;
}

// upperConfidenceStrategy: index of the chosen child, or -1 when a child without episodes makes UCTWeight divide by zero
// (DivideByZeroException: swallowed by getBestStatesSequence :120)
__device__ int ucs_pick(const TreeNode* nodes, int node, unsigned long long key, unsigned& ctr)
{
    const TreeNode& p = nodes[node];
    const int n = p.n_children, first = p.first_child;
    int best = (int)(philox_first(key, (unsigned long long)ctr++, 0u) % (unsigned)n);
    if (nodes[first + best].episodes == 0) return -1;
    float best_w = uct_weight(p, nodes[first + best]);
    for (int j = 0; j < n; ++j) {
        if (nodes[first + j].episodes == 0) return -1;
        const float w = uct_weight(p, nodes[first + j]);
        if (w > best_w) { best_w = w; best = j; }
    }
    return best;
}

constexpr int TREE_THREADS = 128;

template <int MINB>
__global__ void __launch_bounds__(TREE_THREADS, MINB) tree_search_kernel(const DevGame* __restrict__ gg, const hk_game_state* __restrict__ roots,
                                                                   int iterations, int R, unsigned long long seed, int root_base, int max_nodes,
                                                                   TreeNode* __restrict__ slabs, hk_game_state* __restrict__ best_out,
                                                                   int* __restrict__ n_best_out, int* __restrict__ root_episodes,
                                                                   double* __restrict__ root_values, int* __restrict__ n_nodes_out,
                                                                   int* __restrict__ status_out)
{
    __shared__ DevGame g;
    __shared__ unsigned long long s_keys[HK_MAX_ACTIONS];
    __shared__ unsigned s_vis[HK_MAX_ACTIONS], s_nan[HK_MAX_ACTIONS];
    __shared__ double s_rsum[HK_MAX_ACTIONS][HK_MAX_KARTS];
    __shared__ float s_w[HK_MAX_ACTIONS];
    __shared__ int s_leaf, s_cnt, s_first, s_nnodes, s_stop, s_err;
    __shared__ RootMoves s_root;
    __shared__ unsigned long long s_rkeys[HK_MAX_KARTS][HK_MAX_ACTIONS];
    {
        const int* src = reinterpret_cast<const int*>(gg);
        int* dst = reinterpret_cast<int*>(&g);
        for (int i = threadIdx.x; i < (int)(sizeof(DevGame) / 4); i += blockDim.x) dst[i] = src[i];
    }
    const int root = blockIdx.x;
    TreeNode* nodes = slabs + (size_t)root * max_nodes;
    const unsigned long long rseed = seed + (unsigned long long)(root_base + root);   // rollouts of root r: Philox key seed + r
    const unsigned long long ukey = rseed ^ 0x9E3779B97F4A7C15ull;               // stream of the tie-breaking picks
    unsigned ctr = 0;                                                            // thread 0 only
    if (threadIdx.x == 0) {
        TreeNode& r = nodes[0];
        r.st = roots[root]; r.total = 0.0; r.episodes = 0; r.parent = -1; r.first_child = -1; r.n_children = 0;
        r.upnext = up_next(r.st);
        s_nnodes = 1; s_stop = 0; s_err = 0; s_leaf = 0;
    }
    __syncthreads();
    // The root's children still hold the root's state for every kart but the mover, and a root kart usually stands at the (0, bucket)
    // velocity bucket (quirk B.6-1) that the tables do not cover: its policy-ordered legal list is the same in every rollout of
    // iteration 0 (see RootMoves), so it is built once per tree instead of once per rollout.
    if (threadIdx.x < HK_MAX_KARTS) {
        const int k = threadIdx.x;
        const hk_game_state& st = nodes[0].st;
        int cnt = -1;
        if (k < st.n_karts && k != nodes[0].upnext) {
            const int lvl = (g.tables_ok && st.karts[k].player == 0) ? velocity_level(g, st.karts[k].min_velocity, st.karts[k].max_velocity) : -1;
            if (lvl < 0) cnt = legal_moves(g, st, k, s_rkeys[k]);
        }
        s_root.cnt[k] = cnt;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < HK_MAX_KARTS * HK_MAX_ACTIONS; e += blockDim.x) {
        const int k = e / HK_MAX_ACTIONS, c = e % HK_MAX_ACTIONS;
        if (s_root.cnt[k] > 0 && c < g.n_cand && s_rkeys[k][c] != ~0ull) {
            int rank = 0;
            for (int j = 0; j < g.n_cand; ++j) rank += s_rkeys[k][j] < s_rkeys[k][c];
            s_root.order[k][rank] = (unsigned char)c;
        }
    }
    __syncthreads();
    for (int it = 0; it < iterations; ++it) {
        // findLeaf (:194-201): a node either has all its children or none (processLeaf creates them together).  One level per round:
        // the threads evaluate UCTWeight of the children in parallel (a double-precision log each), thread 0 makes the pick.
        for (;;) {
            const int node = s_leaf;
            const int n = nodes[node].n_children;
            if (n == 0) break;                                                   // uniform: s_leaf and the node are block-wide state
            if ((int)threadIdx.x < n) {
                const TreeNode& c = nodes[nodes[node].first_child + threadIdx.x];
                s_w[threadIdx.x] = c.episodes == 0 ? 0.0f : uct_weight(nodes[node], c);
                if (c.episodes == 0) s_err = 2;                                  // UCTWeight would divide by zero
            }
            __syncthreads();
            if (threadIdx.x == 0 && s_err == 0) {
                int best = (int)(philox_first(ukey, (unsigned long long)ctr++, 0u) % (unsigned)n);
                float best_w = s_w[best];
                for (int j = 0; j < n; ++j)
                    if (s_w[j] > best_w) { best_w = s_w[j]; best = j; }
                s_leaf = nodes[node].first_child + best;
            }
            __syncthreads();
            if (s_err) break;
        }
        const int leaf = s_leaf;
        const int np = nodes[leaf].upnext;
        if (s_err == 0 && np >= 0 && (int)threadIdx.x < g.n_cand) s_keys[threadIdx.x] = legal_move_key(g, nodes[leaf].st, np, threadIdx.x);
        for (int i = threadIdx.x; i < HK_MAX_ACTIONS; i += blockDim.x) {
            s_vis[i] = 0; s_nan[i] = 0;
            for (int k = 0; k < HK_MAX_KARTS; ++k) s_rsum[i][k] = 0.0;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            s_cnt = 0;
            const TreeNode& lf = nodes[leaf];
            if (s_err == 0) {
                if (np < 0) s_err = 1;                                           // ArgumentOutOfRangeException at KartDiscreteGame.cs:326
                else {
                    float scores[2 * HK_MAX_KARTS];
                    int ns, cnt = 0;
                    for (int gi = 0; gi < g.n_cand; ++gi) cnt += s_keys[gi] != ~0ull;
                    if (is_over(g, lf.st, cnt, np, scores, ns)) {                // terminal leaf: simulate() returns at once (:246-249)
                        for (int n = leaf; n >= 0; n = nodes[n].parent) {
                            const int up = nodes[n].upnext;
                            if (up >= 0 && up < ns) nodes[n].total += (double)scores[up];
                            nodes[n].episodes += 1;
                        }
                    } else if (s_nnodes + cnt <= max_nodes) {
                        s_cnt = cnt; s_first = s_nnodes;
                    } else s_stop = 1;                                           // slab full: the search ends here
                }
            }
            s_leaf = 0;                                                          // the next walk starts at the root
        }
        __syncthreads();
        if (s_err || s_stop) break;
        const int cnt = s_cnt, first = s_first;
        if (cnt == 0) { __syncthreads(); continue; }                             // terminal leaf handled above (uniform); s_* are rewritten next
        // children in generation order (initials[j] = new KartMCTSNode(node.state.makeMove(action), node), :142)
        if (threadIdx.x < g.n_cand && s_keys[threadIdx.x] != ~0ull) {
            int rank = 0;
            for (int j = 0; j < (int)threadIdx.x; ++j) rank += s_keys[j] != ~0ull;
            TreeNode& c = nodes[first + rank];
            c.st = nodes[leaf].st;
            make_move(g, c.st, np, action_of(g, threadIdx.x));
            c.total = 0.0; c.episodes = 0; c.parent = leaf; c.first_child = -1; c.n_children = 0;
            c.upnext = up_next(c.st);
        }
        __syncthreads();
        // R rollouts through every child; rollout r of child j is rollout id it * R * HK_MAX_ACTIONS + j * R + r of key rseed
        const unsigned long long offset = (unsigned long long)it * (unsigned long long)R * HK_MAX_ACTIONS;
        for (int idx = threadIdx.x; idx < cnt * R; idx += blockDim.x) {
            const int j = idx / R, r = idx - j * R;
            float scores[2 * HK_MAX_KARTS];
            int ns, first_gi;
            const int plies = rollout<false>(g, nodes[first + j].st, rseed, offset + (unsigned long long)j * R + r, scores, ns, first_gi,
                                             nullptr, nullptr, leaf == 0 ? &s_root : nullptr);
            if (plies < 0) { s_err = 1; continue; }
            if (plies == 0) continue;                                            // the child is terminal: handled below
            atomicAdd(&s_vis[j], 1u);
            const int nk = nodes[first + j].st.n_karts;
            bool has_nan = false;
            for (int k = 0; k < nk && k < ns; ++k) has_nan |= isnan(scores[k]);
            if (has_nan) { atomicAdd(&s_nan[j], 1u); continue; }
            for (int k = 0; k < nk && k < ns; ++k)
                if (scores[k] != 0.0f) atomicAdd(&s_rsum[j][k], (double)scores[k]);
        }
        __syncthreads();
        // backpropagate (:280-289).  What child j adds to every node of its chain — entry upNext() of its reward sums and its
        // episode count; a terminal child its own scores, R times — is formed by thread j, which also settles the child itself;
        // thread 0 then updates the shared ancestors once each, adding the children's contributions in order j (the order of the
        // reference's loop, so every node sees the same sequence of additions).
        if ((int)threadIdx.x < cnt) {
            const int j = threadIdx.x;
            TreeNode& cn = nodes[first + j];
            int count;
            if (s_vis[j] == 0) {                                                 // terminal child: its own scores, R times
                float scores[2 * HK_MAX_KARTS];
                int ns = 0;
                if (cn.upnext >= 0) {
                    int cc = 0;
                    for (int gi = 0; gi < g.n_cand; ++gi) cc += legal_move_key(g, cn.st, cn.upnext, gi) != ~0ull;
                    is_over(g, cn.st, cc, cn.upnext, scores, ns);
                }
                for (int k = 0; k < HK_MAX_KARTS; ++k) s_rsum[j][k] = k < ns ? (double)scores[k] * R : 0.0;
                count = R;
            } else count = (int)(s_vis[j] - s_nan[j]);
            s_vis[j] = (unsigned)count;
            const int up = cn.upnext;
            if (up >= 0 && up < HK_MAX_KARTS) cn.total += s_rsum[j][up];
            cn.episodes += count;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            nodes[leaf].first_child = first; nodes[leaf].n_children = cnt;
            s_nnodes = first + cnt;
            for (int n = leaf; n >= 0; n = nodes[n].parent) {
                const int up = nodes[n].upnext;
                double t = nodes[n].total;
                int e = nodes[n].episodes;
                for (int j = 0; j < cnt; ++j) {
                    if (up >= 0 && up < HK_MAX_KARTS) t += s_rsum[j][up];
                    e += (int)s_vis[j];
                }
                nodes[n].total = t; nodes[n].episodes = e;
            }
        }
        __syncthreads();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // getBestStatesSequence (:108-122): follow upperConfidenceStrategy down, keep the states in which every kart stands at
        // lastCompletedSection
        int nb = 0, node = 0;
        if (s_err != 1) {
            while (nodes[node].n_children > 0 && nb < HK_MCTS_MAX_SEQ) {
                const int j = ucs_pick(nodes, node, ukey, ctr);
                if (j < 0) break;
                node = nodes[node].first_child + j;
                const hk_game_state& st = nodes[node].st;
                bool all = true;
                for (int i = 0; i < st.n_karts; ++i) all &= st.karts[i].section == st.lastCompletedSection;
                if (all) best_out[(size_t)root * HK_MCTS_MAX_SEQ + nb++] = st;
            }
        }
        n_best_out[root] = nb;
        if (n_nodes_out) n_nodes_out[root] = s_nnodes;
        if (status_out) status_out[root] = s_err == 1 ? 1 : 0;
        if (root_episodes)
            for (int j = 0; j < HK_MAX_ACTIONS; ++j) {
                const bool in = j < nodes[0].n_children;
                root_episodes[(size_t)root * HK_MAX_ACTIONS + j] = in ? nodes[nodes[0].first_child + j].episodes : 0;
                if (root_values) root_values[(size_t)root * HK_MAX_ACTIONS + j] = in ? nodes[nodes[0].first_child + j].total : 0.0;
            }
    }
}

#include "hk_mcts_seq.cuh"

__global__ void __launch_bounds__(128) rollouts_trace_kernel(const DevGame* __restrict__ gg, const hk_game_state* __restrict__ leaf,
                                                             long long n_rollouts, unsigned long long seed,
                                                             unsigned long long rollout_offset, int* n_plies_out,
                                                             hk_action* actions_out, int* choice_out, int* n_scores_out,
                                                             float* scores_out)
{
    __shared__ DevGame g;
    {
        const int* src = reinterpret_cast<const int*>(gg);
        int* dst = reinterpret_cast<int*>(&g);
        for (int i = threadIdx.x; i < (int)(sizeof(DevGame) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rollouts) return;
    float scores[2 * HK_MAX_KARTS];
    hk_action acts[HK_MAX_PLIES];
    int choice[HK_MAX_PLIES];
    int n_scores, first_gi;
    const int plies = rollout<true>(g, *leaf, seed, rollout_offset + (unsigned long long)r, scores, n_scores, first_gi, acts, choice);
    n_plies_out[r] = plies;
    n_scores_out[r] = n_scores;
    for (int k = 0; k < 2 * HK_MAX_KARTS; ++k) scores_out[r * 2 * HK_MAX_KARTS + k] = k < n_scores ? scores[k] : 0.0f;
    for (int k = 0; k < HK_MAX_PLIES; ++k) {
        const bool on = k < plies;
        actions_out[r * HK_MAX_PLIES + k] = on ? acts[k] : hk_action{0, 0, 0};
        choice_out[r * HK_MAX_PLIES + k] = on ? choice[k] : -1;
    }
}

// exact-parity replay: one thread per root; see hk_game_replay_batch in include/hk_abi.h
__global__ void __launch_bounds__(128) replay_kernel(const DevGame* __restrict__ gg, int batch, int len,
                                                     const hk_game_state* __restrict__ roots, const hk_action* __restrict__ actions,
                                                     hk_game_state* states_out, int* upnext_out, int* over_out, int* n_scores_out,
                                                     float* scores_out, int* n_moves_out, hk_action* moves_out, int* moves_index_out)
{
    __shared__ DevGame g;
    {
        const int* src = reinterpret_cast<const int*>(gg);
        int* dst = reinterpret_cast<int*>(&g);
        for (int i = threadIdx.x; i < (int)(sizeof(DevGame) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    hk_game_state st = roots[b];
    unsigned long long keys[HK_MAX_ACTIONS];
    for (int k = 0; k <= len; ++k) {
        const size_t o = (size_t)b * (len + 1) + k;
        const int np = up_next(st);
        float scores[2 * HK_MAX_KARTS];
        int n_scores = 0, cnt = 0, over = -1;
        if (np >= 0) {
            cnt = legal_moves(g, st, np, keys);
            over = is_over(g, st, cnt, np, scores, n_scores) ? 1 : 0;
        }
        if (states_out) states_out[o] = st;
        if (upnext_out) upnext_out[o] = np;
        if (over_out) over_out[o] = over;
        if (n_scores_out) n_scores_out[o] = n_scores;
        if (scores_out) for (int j = 0; j < 2 * HK_MAX_KARTS; ++j) scores_out[o * 2 * HK_MAX_KARTS + j] = j < n_scores ? scores[j] : 0.0f;
        if (n_moves_out) n_moves_out[o] = np >= 0 ? cnt : -1;
        if (moves_out || moves_index_out)
            for (int j = 0; j < HK_MAX_ACTIONS; ++j) {
                int gi = -1;
                hk_action a{0, 0, 0};
                if (j < cnt) { gi = select_kth(keys, g.n_cand, j); a = action_of(g, gi); }
                if (moves_out) moves_out[o * HK_MAX_ACTIONS + j] = a;
                if (moves_index_out) moves_index_out[o * HK_MAX_ACTIONS + j] = gi;
            }
        if (k < len) {
            if (np < 0) continue;                                      // makeMove would throw; state stays
            make_move(g, st, np, actions[(size_t)b * len + k]);
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------------
static double phi(double x) { return 0.5 * erfc(-x / 1.4142135623730951); }

// Index distribution of the rollout policy as cumulative 32-bit thresholds (KartMCTS.cs:266-269, NextGaussian :218-236):
// X ~ N(0, cnt/6f) redrawn while |X| > cnt-1 for at most 10 attempts (then the mean 0), index = RoundToInt(|X|).
void policy_cdf_host(int cnt, uint32_t* cdf)
{
    if (cnt <= 2) {                                                    // random.Next(cnt)
        for (int k = 0; k < cnt; ++k) cdf[k] = (k == cnt - 1) ? 0xFFFFFFFFu : (uint32_t)(((uint64_t)(k + 1) << 32) / (uint64_t)cnt);
        return;
    }
    const double sd = (double)((float)cnt / 6.0f), lim = (double)cnt - 1.0;
    const double rej = 1.0 - 2.0 * (phi(lim / sd) - 0.5);
    double geo = 0.0, rp = 1.0;
    for (int a = 0; a < 10; ++a) { geo += rp; rp *= rej; }
    double c = 0.0;
    for (int k = 0; k < cnt; ++k) {
        const double lo = k == 0 ? 0.0 : k - 0.5;
        double hi = k + 0.5;
        if (hi > lim) hi = lim;
        const double pk = (lo < hi ? 2.0 * (phi(hi / sd) - phi(lo / sd)) : 0.0) * geo + (k == 0 ? rp : 0.0);
        c += pk;
        const double th = c * 4294967296.0;
        cdf[k] = (k == cnt - 1 || th >= 4294967295.0) ? 0xFFFFFFFFu : (uint32_t)th;
    }
}

}  // namespace hk

struct hk_game {
    hk::DevGame host;
    hk::DevGame* dev = nullptr;
};

using namespace hk;

extern "C" int hk_policy_cdf(int cnt, uint32_t* cdf_out)
{
    if (cnt < 1 || cnt > HK_MAX_ACTIONS || !cdf_out) { set_error("hk_policy_cdf: cnt must be 1..%d", HK_MAX_ACTIONS); return HK_ERR_INVALID_ARGUMENT; }
    policy_cdf_host(cnt, cdf_out);
    return HK_OK;
}

extern "C" int hk_game_create(const hk_section* sections, int n_sections, const hk_kart* karts, int n_karts,
                              const hk_kart* env_karts, int n_env_karts, const hk_game_params* params, hk_game** out)
{
    if (!sections || !karts || !params || !out || n_sections < 1 || n_sections > HK_MAX_SECTIONS || n_karts < 1 ||
        n_karts > HK_MAX_KARTS || (env_karts && (n_env_karts < 1 || n_env_karts > HK_MAX_ENV_KARTS)) ||
        params->velocityBucketSize < 1) {
        set_error("hk_game_create: invalid argument (sections 1..%d, karts 1..%d, bucket >= 1)", HK_MAX_SECTIONS, HK_MAX_KARTS);
        return HK_ERR_INVALID_ARGUMENT;
    }
    int rc = ensure_device();
    if (rc != HK_OK) return rc;
    hk_game* g = new hk_game();
    std::memset(&g->host, 0, sizeof(DevGame));
    DevGame& d = g->host;
    d.n_sections = n_sections; d.n_karts = n_karts; d.p = *params;
    std::memcpy(d.sections, sections, sizeof(hk_section) * n_sections);
    std::memcpy(d.karts, karts, sizeof(hk_kart) * n_karts);
    if (env_karts) { d.n_env_karts = n_env_karts; std::memcpy(d.env_karts, env_karts, sizeof(hk_kart) * n_env_karts); }
    else { d.n_env_karts = n_karts; std::memcpy(d.env_karts, karts, sizeof(hk_kart) * n_karts); }
    // (int)GetMaxSpeed() of the candidate loop (:329); all karts of one game share constants in the shipped scenes; the
    // generation index space is sized with kart 0 and the per-kart limit is applied through g.vmax of kart 0 only.
    const float ms = karts[0].topSpeed > karts[0].reverseSpeed ? karts[0].topSpeed : karts[0].reverseSpeed;
    d.vmax = (int)ms;
    for (int k = 1; k < n_karts; ++k) {
        const float mk = karts[k].topSpeed > karts[k].reverseSpeed ? karts[k].topSpeed : karts[k].reverseSpeed;
        if ((int)mk != d.vmax) { delete g; set_error("hk_game_create: karts with different (int)GetMaxSpeed() are not supported"); return HK_ERR_INVALID_ARGUMENT; }
    }
    int nv = 0;
    for (int v = 6; v < d.vmax; v += params->velocityBucketSize) ++nv;
    d.n_cand = nv * 4;
    if (d.n_cand > HK_MAX_ACTIONS) { delete g; set_error("hk_game_create: more than %d candidate actions", HK_MAX_ACTIONS); return HK_ERR_INVALID_ARGUMENT; }
    for (int c = 1; c <= HK_MAX_ACTIONS; ++c) policy_cdf_host(c, d.cdf[c]);
    // geometry types, per-section flags and the table layout; the tables themselves are filled on the device
    d.nv = nv; d.n_types = 0; d.tables_ok = 1;
    for (int s = 0; s < n_sections; ++s) {
        int t = -1;
        for (int k = 0; k < d.n_types && t < 0; ++k) {
            const hk_section& r = sections[d.rep_section[k]];
            if (r.insideR == sections[s].insideR && r.length == sections[s].length && r.width == sections[s].width &&
                r.turnDeg == sections[s].turnDeg && (r.leftTurn != 0) == (sections[s].leftTurn != 0)) t = k;
        }
        if (t < 0) {
            if (d.n_types == HK_MAX_TYPES) { d.tables_ok = 0; t = 0; }
            else { t = d.n_types; d.rep_section[d.n_types++] = (unsigned char)s; }
        }
        d.type_of[s] = (unsigned char)t;
        const bool st0 = sections[s].insideR == 0.0f, st1 = sections[(s + 1) % n_sections].insideR == 0.0f;
        const int sign = sections[s].optimalLane == 1 ? 1 : (sections[s].optimalLane == 4 ? -1 : 0);
        d.sec_flags[s] = (unsigned char)((st0 ? 1 : 0) | ((st0 != st1) ? 2 : 0) | ((sign + 1) << 2));
    }
    const size_t cells = (size_t)d.n_types * 4 * nv;
    d.off_dt = 0;
    d.off_order = (int)(cells * d.n_cand * 4);
    d.off_load = (int)((d.off_order + cells * 3 * d.n_cand + 15) & ~(size_t)15);
    d.off_radius = d.off_load + (int)((size_t)d.n_types * 16 * nv * 4);
    d.off_lmask = (d.off_radius + d.n_types * 16 * 4 + 15) & ~15;
    d.off_od = d.off_lmask + (int)(cells * 3 * 4 * nv * 8);
    d.table_bytes = d.off_od + (int)(cells * 3 * d.n_cand * 8);
    unsigned char* blob = nullptr;
    cudaError_t e = cudaMalloc(&g->dev, sizeof(DevGame));
    if (e == cudaSuccess) e = cudaMalloc(&blob, d.table_bytes);
    d.tables = blob;
    if (e == cudaSuccess) e = cudaMemcpy(g->dev, &g->host, sizeof(DevGame), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && d.tables_ok) {
        count_launch(); build_tables_kernel<<<(unsigned)((cells + 63) / 64), 64>>>(g->dev, blob);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
    }
    if (e != cudaSuccess) { set_error("hk_game_create: %s", cudaGetErrorString(e)); if (g->dev) cudaFree(g->dev); if (blob) cudaFree(blob); delete g; return HK_ERR_CUDA; }
    *out = g;
    return HK_OK;
}

extern "C" void hk_game_destroy(hk_game* g)
{
    if (!g) return;
    if (g->dev) cudaFree(g->dev);
    if (g->host.tables) cudaFree(const_cast<unsigned char*>(g->host.tables));
    delete g;
}

static int check_state(const hk_game* g, const hk_game_state* s, const char* who)
{
    if (s->n_karts < 1 || s->n_karts > g->host.n_karts) { set_error("%s: state has %d karts, game has %d", who, s->n_karts, g->host.n_karts); return HK_ERR_INVALID_ARGUMENT; }
    for (int i = 0; i < s->n_karts; ++i)
        if (s->karts[i].player < 0 || s->karts[i].player >= g->host.n_env_karts || s->karts[i].lane < 1 || s->karts[i].lane > 4 ||
            s->karts[i].section < 0) { set_error("%s: kart %d has an invalid player/lane/section", who, i); return HK_ERR_INVALID_ARGUMENT; }
    return HK_OK;
}

extern "C" int hk_game_replay_batch(const hk_game* g, int batch, int len, const hk_game_state* roots, const hk_action* actions,
                                    hk_game_state* states_out, int32_t* upnext_out, int32_t* over_out, int32_t* n_scores_out,
                                    float* scores_out, int32_t* n_moves_out, hk_action* moves_out, int32_t* moves_index_out)
{
    if (!g || batch < 0 || len < 0 || (batch && !roots) || (batch && len && !actions)) { set_error("hk_game_replay_batch: invalid argument"); return HK_ERR_INVALID_ARGUMENT; }
    if (batch == 0) return HK_OK;
    for (int b = 0; b < batch; ++b) { int rc = check_state(g, &roots[b], "hk_game_replay_batch"); if (rc) return rc; }
    for (size_t i = 0; i < (size_t)batch * len; ++i)
        if (actions[i].lane < 1 || actions[i].lane > 4) { set_error("hk_game_replay_batch: action lane out of 1..4"); return HK_ERR_INVALID_ARGUMENT; }
    ThreadCtx* c = ctx();
    if (!c) return HK_ERR_NO_DEVICE;
    const size_t np1 = (size_t)batch * (len + 1);
    const size_t sz[10] = {sizeof(hk_game_state) * batch, sizeof(hk_action) * (size_t)batch * len, sizeof(hk_game_state) * np1, 4 * np1, 4 * np1,
                           4 * np1, 4 * np1 * 2 * HK_MAX_KARTS, 4 * np1, sizeof(hk_action) * np1 * HK_MAX_ACTIONS, 4 * np1 * HK_MAX_ACTIONS};
    size_t off[11]; off[0] = 0;
    for (int i = 0; i < 10; ++i) off[i + 1] = off[i] + ((sz[i] + 255) & ~(size_t)255);
    char* d = (char*)dscratch(c, 0, off[10]);
    if (!d) return HK_ERR_OUT_OF_MEMORY;
    HK_CUDA(cudaMemcpyAsync(d + off[0], roots, sz[0], cudaMemcpyHostToDevice, c->stream));
    if (len) HK_CUDA(cudaMemcpyAsync(d + off[1], actions, sz[1], cudaMemcpyHostToDevice, c->stream));
    count_launch(); replay_kernel<<<(batch + 127) / 128, 128, 0, c->stream>>>(g->dev, batch, len, (const hk_game_state*)(d + off[0]), (const hk_action*)(d + off[1]),
        states_out ? (hk_game_state*)(d + off[2]) : nullptr, upnext_out ? (int*)(d + off[3]) : nullptr, over_out ? (int*)(d + off[4]) : nullptr,
        n_scores_out ? (int*)(d + off[5]) : nullptr, scores_out ? (float*)(d + off[6]) : nullptr, n_moves_out ? (int*)(d + off[7]) : nullptr,
        moves_out ? (hk_action*)(d + off[8]) : nullptr, moves_index_out ? (int*)(d + off[9]) : nullptr);
    HK_CUDA(cudaGetLastError());
    void* outs[8] = {states_out, upnext_out, over_out, n_scores_out, scores_out, n_moves_out, moves_out, moves_index_out};
    for (int i = 0; i < 8; ++i)
        if (outs[i]) HK_CUDA(cudaMemcpyAsync(outs[i], d + off[2 + i], sz[2 + i], cudaMemcpyDeviceToHost, c->stream));
    HK_CUDA(cudaStreamSynchronize(c->stream));
    return HK_OK;
}

static int rollouts_impl(const hk_game* g, const hk_game_state* leaves, int n_leaves, int64_t rollouts_per_leaf, uint64_t seed,
                         uint64_t rollout_offset, int64_t* visit, double* reward_sum, int64_t* nan_count, int64_t* plies_sum)
{
    if (!g || !leaves || n_leaves < 1 || rollouts_per_leaf < 0 || !visit || !reward_sum || !nan_count) { set_error("hk_mcts_rollouts: invalid argument"); return HK_ERR_INVALID_ARGUMENT; }
    for (int l = 0; l < n_leaves; ++l) { int rc = check_state(g, &leaves[l], "hk_mcts_rollouts"); if (rc) return rc; }
    ThreadCtx* c = ctx();
    if (!c) return HK_ERR_NO_DEVICE;
    const size_t nA = (size_t)n_leaves * HK_MAX_ACTIONS;
    const size_t sz[6] = {sizeof(hk_game_state) * n_leaves, 8 * nA, 8 * nA * HK_MAX_KARTS, 8 * nA, 8 * (size_t)n_leaves, 256 + 8 * (size_t)n_leaves};
    size_t off[7]; off[0] = 0;
    for (int i = 0; i < 6; ++i) off[i + 1] = off[i] + ((sz[i] + 255) & ~(size_t)255);
    char* d = (char*)dscratch(c, 0, off[6]);
    if (!d) return HK_ERR_OUT_OF_MEMORY;
    HK_CUDA(cudaMemcpyAsync(d, leaves, sz[0], cudaMemcpyHostToDevice, c->stream));
    HK_CUDA(cudaMemsetAsync(d + off[1], 0, off[6] - off[1], c->stream));
    if (rollouts_per_leaf > 0) {
        int sms = 148, cur_dev = 0;
        cudaGetDevice(&cur_dev);                                          // the device this process is bound to (hk_init), not device 0
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cur_dev);
        long long want = (rollouts_per_leaf + 127) / 128;
        long long cap = (long long)sms * 16 / (n_leaves < sms * 16 ? 1 : 1);          // persistent-ish grid: 16 CTAs of 128 threads per SM
        if (n_leaves > 1) cap = (cap + n_leaves - 1) / n_leaves > 0 ? (cap + n_leaves - 1) / n_leaves : 1;
        long long gx = want < cap ? want : cap;
        if (n_leaves == 1 && want > (long long)sms * 6) {
            // one leaf: a single wave of resident thread blocks (6 per SM at 80 registers) with the same number of rollouts per thread — the grid-stride
            // loop otherwise ends in a partly filled round and a second, partly filled wave of blocks (ncu: `barrier` 1.2 warps per issue cycle)
            const long long resident = (long long)sms * 6, rounds = (want + resident - 1) / resident;
            gx = (want + rounds - 1) / rounds;
        }
        dim3 grid((unsigned)gx, (unsigned)n_leaves);
        count_launch(); rollouts_kernel<<<grid, 128, 0, c->stream>>>(g->dev, (const hk_game_state*)d, rollouts_per_leaf, seed, rollout_offset,
            (unsigned long long*)(d + off[1]), (double*)(d + off[2]), (unsigned long long*)(d + off[3]), (unsigned long long*)(d + off[4]), (int*)(d + off[5]),
            (unsigned long long*)(d + off[5] + 256));
        HK_CUDA(cudaGetLastError());
    }
    int err = 0;
    HK_CUDA(cudaMemcpyAsync(visit, d + off[1], sz[1], cudaMemcpyDeviceToHost, c->stream));
    HK_CUDA(cudaMemcpyAsync(reward_sum, d + off[2], sz[2], cudaMemcpyDeviceToHost, c->stream));
    HK_CUDA(cudaMemcpyAsync(nan_count, d + off[3], sz[3], cudaMemcpyDeviceToHost, c->stream));
    if (plies_sum) HK_CUDA(cudaMemcpyAsync(plies_sum, d + off[4], sz[4], cudaMemcpyDeviceToHost, c->stream));
    HK_CUDA(cudaMemcpyAsync(&err, d + off[5], 4, cudaMemcpyDeviceToHost, c->stream));
    HK_CUDA(cudaStreamSynchronize(c->stream));
    if (err) { set_error("hk_mcts_rollouts: upNext() == -1 reached (KartDiscreteGame.cs:326 would throw)"); return HK_ERR_NO_UPNEXT; }
    return HK_OK;
}

extern "C" int hk_mcts_rollouts(const hk_game* g, const hk_game_state* leaf, int64_t n_rollouts, uint64_t seed, uint64_t rollout_offset,
                                int64_t* visit, double* reward_sum, int64_t* nan_count, int64_t* plies_sum)
{
    return rollouts_impl(g, leaf, 1, n_rollouts, seed, rollout_offset, visit, reward_sum, nan_count, plies_sum);
}

extern "C" int hk_mcts_rollouts_multi(const hk_game* g, const hk_game_state* leaves, int n_leaves, int64_t rollouts_per_leaf, uint64_t seed,
                                      uint64_t rollout_offset, int64_t* visit, double* reward_sum, int64_t* nan_count, int64_t* plies_sum)
{
    return rollouts_impl(g, leaves, n_leaves, rollouts_per_leaf, seed, rollout_offset, visit, reward_sum, nan_count, plies_sum);
}

namespace hk {
const hk_game_params& game_params_of(const hk_game* g) { return g->host.p; }
int game_karts_of(const hk_game* g) { return g->host.n_karts; }

int mcts_search_device(const hk_game* g, const hk_game_state* d_roots, int n_roots, int iterations, int rollouts_per_leaf, uint64_t seed,
                       hk_game_state* d_best, int* d_nbest, int* d_eps, double* d_vals, int* d_nnodes, int* d_status, ThreadCtx* c,
                       cudaStream_t s)
{
    const int max_nodes = 1 + iterations * HK_MAX_ACTIONS;
    // chunks of roots: the tree slabs of one chunk stay below ~2 GB whatever the batch (1,184 thread blocks are resident at a time)
    long long per_chunk = (long long)(2.0e9 / ((double)sizeof(TreeNode) * max_nodes));
    if (per_chunk < 2368) per_chunk = 2368;
    const int chunk = (int)(per_chunk < n_roots ? per_chunk : n_roots);
    TreeNode* slabs = (TreeNode*)dscratch(c, 12, sizeof(TreeNode) * (size_t)chunk * max_nodes);
    if (!slabs) return HK_ERR_OUT_OF_MEMORY;
    HK_CUDA(cudaMemsetAsync(d_best, 0, sizeof(hk_game_state) * (size_t)n_roots * HK_MCTS_MAX_SEQ, s));   // entries past n_best stay zero
    for (int base = 0; base < n_roots; base += chunk) {
        const int nr = base + chunk <= n_roots ? chunk : n_roots - base;
        count_launch();
        static const int minb = getenv("HK_TREE_MINB") ? atoi(getenv("HK_TREE_MINB")) : 8;
#define HK_TREE_LAUNCH(M) tree_search_kernel<M><<<(unsigned)nr, TREE_THREADS, 0, s>>>(g->dev, d_roots + base, iterations, rollouts_per_leaf, seed, base, max_nodes, slabs, \
            d_best + (size_t)base * HK_MCTS_MAX_SEQ, d_nbest + base, d_eps ? d_eps + (size_t)base * HK_MAX_ACTIONS : nullptr,                           \
            d_vals ? d_vals + (size_t)base * HK_MAX_ACTIONS : nullptr, d_nnodes ? d_nnodes + base : nullptr, d_status + base)
        if (minb == 6) HK_TREE_LAUNCH(6); else if (minb == 5) HK_TREE_LAUNCH(5); else if (minb == 7) HK_TREE_LAUNCH(7); else HK_TREE_LAUNCH(8);
#undef HK_TREE_LAUNCH
        HK_CUDA(cudaGetLastError());
    }
    return HK_OK;
}
}  // namespace hk

extern "C" int hk_mcts_search_batch(const hk_game* g, const hk_game_state* roots, int n_roots, int iterations, int rollouts_per_leaf,
                                    uint64_t seed, hk_game_state* best_states, int32_t* n_best, int32_t* root_episodes,
                                    double* root_values, int32_t* n_nodes)
{
    if (!g || !roots || n_roots < 1 || iterations < 0 || rollouts_per_leaf < 1 || !best_states || !n_best) { set_error("hk_mcts_search_batch: invalid argument"); return HK_ERR_INVALID_ARGUMENT; }
    for (int r = 0; r < n_roots; ++r) { int rc = check_state(g, &roots[r], "hk_mcts_search_batch"); if (rc) return rc; }
    ThreadCtx* c = ctx();
    if (!c) return HK_ERR_NO_DEVICE;
    const size_t nA = (size_t)n_roots * HK_MAX_ACTIONS;
    const size_t sz[7] = {sizeof(hk_game_state) * n_roots, sizeof(hk_game_state) * (size_t)n_roots * HK_MCTS_MAX_SEQ, 4 * (size_t)n_roots, 4 * nA, 8 * nA,
                          4 * (size_t)n_roots, 4 * (size_t)n_roots};
    size_t off[8]; off[0] = 0;
    for (int i = 0; i < 7; ++i) off[i + 1] = off[i] + ((sz[i] + 255) & ~(size_t)255);
    char* d = (char*)dscratch(c, 0, off[7]);
    if (!d) return HK_ERR_OUT_OF_MEMORY;
    HK_CUDA(cudaMemcpyAsync(d, roots, sz[0], cudaMemcpyHostToDevice, c->stream));
    int rc = mcts_search_device(g, (const hk_game_state*)d, n_roots, iterations, rollouts_per_leaf, seed, (hk_game_state*)(d + off[1]), (int*)(d + off[2]),
                                (int*)(d + off[3]), (double*)(d + off[4]), (int*)(d + off[5]), (int*)(d + off[6]), c, c->stream);
    if (rc) return rc;
    std::vector<int> st((size_t)n_roots);
    HK_CUDA(cudaMemcpyAsync(best_states, d + off[1], sz[1], cudaMemcpyDeviceToHost, c->stream));
    HK_CUDA(cudaMemcpyAsync(n_best, d + off[2], sz[2], cudaMemcpyDeviceToHost, c->stream));
    if (root_episodes) HK_CUDA(cudaMemcpyAsync(root_episodes, d + off[3], sz[3], cudaMemcpyDeviceToHost, c->stream));
    if (root_values) HK_CUDA(cudaMemcpyAsync(root_values, d + off[4], sz[4], cudaMemcpyDeviceToHost, c->stream));
    if (n_nodes) HK_CUDA(cudaMemcpyAsync(n_nodes, d + off[5], sz[5], cudaMemcpyDeviceToHost, c->stream));
    HK_CUDA(cudaMemcpyAsync(st.data(), d + off[6], sz[6], cudaMemcpyDeviceToHost, c->stream));
    HK_CUDA(cudaStreamSynchronize(c->stream));
    for (int r = 0; r < n_roots; ++r)
        if (st[r]) { set_error("hk_mcts_search_batch: upNext() == -1 reached in the tree of root %d (KartDiscreteGame.cs:326 would throw)", r); return HK_ERR_NO_UPNEXT; }
    return HK_OK;
}

// ---- sequential (faithful) tree search: forest of device-resident trees -----------------------------------------------------------
struct hk_mcts_forest {
    const hk_game* g = nullptr;
    int n_trees = 0, max_nodes = 0;
    hk::SeqTree* trees = nullptr;
    hk_mcts_node* slabs = nullptr;
    unsigned* recs = nullptr;             // fast path: playout records of one chunk of iterations [n_trees][chunk][seq_rec_words(cap)]
    size_t recs_bytes = 0;
    int* remaining = nullptr;             // fast path: -1 while a tree is on it, else the iterations the general kernel still owes it
    int max_plies = 0;                    // largest playout length any root given through the host entry can have
    int* aux = nullptr;                   // fast path: per-tree prefix tables (hk_mcts_seq.cuh, seq_insert_kernel); optional
    long long aux_stride = 0;             // ints per tree
    int aux_levels = 0;
};

namespace hk {
// (float)Math.Log((double)(float)k) of UCTWeight (KartMCTS.cs:164) for the integer ratios a search can produce, computed by the host's
// libm (the same function the oracle calls) so that device trees are bit-equal to the oracle's; one table per process.
constexpr int SEQ_LOG_TABLE = 1 << 16;
static float* g_logtab = nullptr;
static std::mutex g_logtab_mu;
static const float* seq_log_table()
{
    std::lock_guard<std::mutex> lk(g_logtab_mu);
    if (g_logtab) return g_logtab;
    std::vector<float> h((size_t)SEQ_LOG_TABLE);
    for (int k = 0; k < SEQ_LOG_TABLE; ++k) h[k] = k == 0 ? -INFINITY : (float)std::log((double)(float)k);
    float* d = nullptr;
    if (cudaMalloc(&d, sizeof(float) * SEQ_LOG_TABLE) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaMemcpy(d, h.data(), sizeof(float) * SEQ_LOG_TABLE, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); cudaGetLastError(); return nullptr; }
    g_logtab = d;
    return g_logtab;
}

int mcts_seq_search_device(hk_mcts_forest* f, const hk_game_state* d_roots, const int* d_fresh, int iterations, uint64_t seed,
                           hk_game_state* d_best, int* d_nbest, int* d_nnodes, int* d_status, cudaStream_t s, bool clear_best, int max_plies)
{
    const float* lt = seq_log_table();
    if (!lt) { set_error("hk_mcts_forest_search: log table allocation failed"); return HK_ERR_OUT_OF_MEMORY; }
    if (d_best && clear_best) HK_CUDA(cudaMemsetAsync(d_best, 0, sizeof(hk_game_state) * (size_t)f->n_trees * HK_MCTS_MAX_SEQ, s));   // entries past n_best stay zero
    // One warp per block, 32 trees per warp.  Fewer trees per warp (HK_SEQ_LANES = 16 / 8 / 4: more, less divergent warps) was measured and
    // is SLOWER — 32,768 trees x 512 iterations: 69 / 107 / 136 / 208 ms of host call at 32 / 16 / 8 / 4 — a warp's time is set by its
    // longest lane (dependent node-chain latency), not by the sum of its lanes' paths, so halving the lanes only doubles the warps.
    static const int lanes_env = getenv("HK_SEQ_LANES") ? atoi(getenv("HK_SEQ_LANES")) : 0;
    int lanes = 32;
    if (lanes_env == 8 || lanes_env == 16 || lanes_env == 32 || lanes_env == 4) lanes = lanes_env;
    const unsigned tree_blocks = (unsigned)((f->n_trees + lanes - 1) / lanes);
    const char* fast_env = getenv("HK_SEQ_FAST");                    // read per call: tests run both paths in one process
    const bool fast = !(fast_env && atoi(fast_env) == 0);
    if (!fast || iterations == 0) {                                   // the general kernel alone: init, every iteration, best states
        count_launch();
        seq_search_kernel<<<tree_blocks, 32, 0, s>>>(f->g->dev, f->trees, f->slabs, f->max_nodes, f->n_trees, iterations, nullptr, seed, 0, d_roots, d_fresh,
                                                     lt, SEQ_LOG_TABLE, d_best, d_nbest, d_nnodes, d_status, lanes, true, true);
        HK_CUDA(cudaGetLastError());
        return HK_OK;
    }
    // fast path (hk_mcts_seq.cuh): init; per chunk of iterations the playouts of every (tree, iteration) in parallel and their insertion in
    // order; then the general kernel for trees whose root became fully expanded on the way, and getBestStatesSequence for all
    const int cap = max_plies > 0 && max_plies < HK_MAX_PLIES ? max_plies : HK_MAX_PLIES;
    const size_t words = seq_rec_words(cap);
    long long chunk = (long long)(768.0e6 / ((double)f->n_trees * words * 4));
    chunk = chunk < 16 ? 16 : chunk > 512 ? 512 : chunk;
    if (chunk > iterations) chunk = iterations;
    const size_t need = (size_t)f->n_trees * (size_t)chunk * words * 4;
    if (need > f->recs_bytes) {
        HK_CUDA(cudaStreamSynchronize(s));
        if (f->recs) cudaFree(f->recs);
        f->recs = nullptr; f->recs_bytes = 0;
        HK_CUDA(cudaMalloc(&f->recs, need));
        f->recs_bytes = need;
    }
    if (!f->remaining) HK_CUDA(cudaMalloc(&f->remaining, sizeof(int) * (size_t)f->n_trees));
    HK_CUDA(cudaMemsetAsync(f->remaining, 0xff, sizeof(int) * (size_t)f->n_trees, s));
    count_launch();
    seq_search_kernel<<<tree_blocks, 32, 0, s>>>(f->g->dev, f->trees, f->slabs, f->max_nodes, f->n_trees, 0, nullptr, seed, 0, d_roots, d_fresh, lt,
                                                 SEQ_LOG_TABLE, nullptr, nullptr, nullptr, nullptr, lanes, true, false);
    HK_CUDA(cudaGetLastError());
    if (f->aux) {                                                     // fresh trees start with an empty prefix table (and the right to use it)
        const unsigned bx = (unsigned)((f->aux_stride + 1023) / 1024 < 16 ? (f->aux_stride + 1023) / 1024 : 16);
        count_launch();
        for (int t0 = 0; t0 < f->n_trees; t0 += 65535) {              // gridDim.y limit
            const int nt = f->n_trees - t0 < 65535 ? f->n_trees - t0 : 65535;
            seq_aux_clear_kernel<<<dim3(bx, (unsigned)nt), 256, 0, s>>>(f->trees + t0, f->aux + (size_t)t0 * f->aux_stride, f->aux_stride, nt, d_fresh ? d_fresh + t0 : nullptr);
        }
        HK_CUDA(cudaGetLastError());
    }
    // Playouts and insertion alternate on the caller's stream.  Running the playouts of chunk c + 1 on a second stream beside the insertion
    // of chunk c (two record buffers) was measured and does not pay: 61.6 ms per 32,768 x 512 call against 48.8 ms (the playout grid's blocks
    // crowd the 1,024 one-warp insertion blocks off the SMs), 49.6 ms with the playout stream at the lowest priority.
    for (int base = 0; base < iterations; base += (int)chunk) {
        const int count = base + chunk <= iterations ? (int)chunk : iterations - base;
        const long long threads = (long long)f->n_trees * count;
        count_launch();
        static const bool packn = !(getenv("HK_SEQ_PACKN") && atoi(getenv("HK_SEQ_PACKN")) == 0);     // measurement knob: struct-based playouts for 3-4 karts
        if (f->g->host.n_karts >= 3 && packn)
            seq_playouts_kernel<4><<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(f->g->dev, f->trees, f->n_trees, d_fresh, f->remaining, count, base, cap, f->recs);
        else
            seq_playouts_kernel<2><<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(f->g->dev, f->trees, f->n_trees, d_fresh, f->remaining, count, base, cap, f->recs);
        HK_CUDA(cudaGetLastError());
        count_launch();
        seq_insert_kernel<<<(unsigned)((f->n_trees + 31) / 32), 32, 0, s>>>(f->trees, f->slabs, f->max_nodes, f->n_trees, d_fresh, f->remaining, count, count,
                                                                          iterations - base - count, cap, f->recs, f->aux, f->aux_stride, f->aux_levels,
                                                                          f->g->host.n_cand);
        HK_CUDA(cudaGetLastError());
    }
    count_launch();
    seq_search_kernel<<<tree_blocks, 32, 0, s>>>(f->g->dev, f->trees, f->slabs, f->max_nodes, f->n_trees, 0, f->remaining, seed, 0, d_roots, d_fresh, lt,
                                                 SEQ_LOG_TABLE, d_best, d_nbest, d_nnodes, d_status, lanes, false, true);
    HK_CUDA(cudaGetLastError());
    return HK_OK;
}
}  // namespace hk

extern "C" int hk_mcts_forest_create(const hk_game* g, int n_trees, int max_nodes_per_tree, hk_mcts_forest** out)
{
    if (!g || n_trees < 1 || max_nodes_per_tree < 1 || !out) { set_error("hk_mcts_forest_create: invalid argument"); return HK_ERR_INVALID_ARGUMENT; }
    int rc = ensure_device();
    if (rc != HK_OK) return rc;
    hk_mcts_forest* f = new hk_mcts_forest();
    f->g = g; f->n_trees = n_trees; f->max_nodes = max_nodes_per_tree;
    cudaError_t e = cudaMalloc(&f->trees, sizeof(SeqTree) * (size_t)n_trees);
    if (e == cudaSuccess) e = cudaMalloc(&f->slabs, sizeof(hk_mcts_node) * (size_t)n_trees * max_nodes_per_tree);
    if (e == cudaSuccess) e = cudaMemset(f->trees, 0, sizeof(SeqTree) * (size_t)n_trees);
    if (e == cudaSuccess && !(getenv("HK_SEQ_AUX") && atoi(getenv("HK_SEQ_AUX")) == 0)) {
        // prefix tables over the first 3 (2, 1) plies, as many levels as fit 2 GiB for this forest; doing without is only slower
        const long long nc = g->host.n_cand;
        long long stride = 0;
        int levels = 0;
        const long long budget = getenv("HK_SEQ_AUX_LEVELS") ? 1ll << 40 : 2ll << 30;
        const int want = getenv("HK_SEQ_AUX_LEVELS") ? atoi(getenv("HK_SEQ_AUX_LEVELS")) : 3;
        for (int l = 1; l <= want && l <= 3; ++l) {
            const long long st = l == 1 ? nc : l == 2 ? nc + nc * nc : nc + nc * nc + nc * nc * nc;
            if (st * 4 * n_trees <= budget) { stride = st; levels = l; }
        }
        if (levels && cudaMalloc(&f->aux, (size_t)stride * 4 * (size_t)n_trees) == cudaSuccess) { f->aux_stride = stride; f->aux_levels = levels; }
        else { f->aux = nullptr; cudaGetLastError(); }
    }
    if (e != cudaSuccess) {
        set_error("hk_mcts_forest_create: %s (%d trees x %d nodes x 32 B)", cudaGetErrorString(e), n_trees, max_nodes_per_tree);
        cudaGetLastError();
        if (f->trees) cudaFree(f->trees);
        if (f->slabs) cudaFree(f->slabs);
        delete f;
        return e == cudaErrorMemoryAllocation ? HK_ERR_OUT_OF_MEMORY : HK_ERR_CUDA;
    }
    *out = f;
    return HK_OK;
}

extern "C" void hk_mcts_forest_destroy(hk_mcts_forest* f)
{
    if (!f) return;
    if (f->trees) cudaFree(f->trees);
    if (f->slabs) cudaFree(f->slabs);
    if (f->recs) cudaFree(f->recs);
    if (f->remaining) cudaFree(f->remaining);
    if (f->aux) cudaFree(f->aux);
    delete f;
}

extern "C" int hk_mcts_forest_search(hk_mcts_forest* f, const hk_game_state* roots, const int32_t* fresh, int iterations, uint64_t seed,
                                     hk_game_state* best_states, int32_t* n_best, int32_t* n_nodes, int32_t* status)
{
    if (!f || iterations < 0 || !best_states || !n_best) { set_error("hk_mcts_forest_search: invalid argument"); return HK_ERR_INVALID_ARGUMENT; }
    const int n = f->n_trees;
    bool any_fresh = false;
    for (int r = 0; r < n; ++r) {
        if (fresh && fresh[r] <= 0) continue;
        any_fresh = true;
        if (!roots) { set_error("hk_mcts_forest_search: roots is NULL but tree %d is fresh", r); return HK_ERR_INVALID_ARGUMENT; }
        int rc = check_state(f->g, &roots[r], "hk_mcts_forest_search");
        if (rc) return rc;
        long long pl = 0;
        for (int k = 0; k < roots[r].n_karts; ++k) { const long long dd = (long long)roots[r].finalSection - roots[r].karts[k].section; pl += dd > 0 ? dd : 0; }
        if (pl > HK_MAX_PLIES) { set_error("hk_mcts_forest_search: a playout of tree %d could take %lld plies (> %d)", r, pl, HK_MAX_PLIES); return HK_ERR_INVALID_ARGUMENT; }
        if ((int)pl > f->max_plies) f->max_plies = (int)pl;
    }
    ThreadCtx* c = ctx();
    if (!c) return HK_ERR_NO_DEVICE;
    const size_t sz[6] = {sizeof(hk_game_state) * (size_t)n, 4 * (size_t)n, sizeof(hk_game_state) * (size_t)n * HK_MCTS_MAX_SEQ, 4 * (size_t)n, 4 * (size_t)n, 4 * (size_t)n};
    size_t off[7]; off[0] = 0;
    for (int i = 0; i < 6; ++i) off[i + 1] = off[i] + ((sz[i] + 255) & ~(size_t)255);
    char* d = (char*)dscratch(c, 0, off[6]);
    if (!d) return HK_ERR_OUT_OF_MEMORY;
    if (any_fresh) HK_CUDA(cudaMemcpyAsync(d, roots, sz[0], cudaMemcpyHostToDevice, c->stream));
    if (fresh) HK_CUDA(cudaMemcpyAsync(d + off[1], fresh, sz[1], cudaMemcpyHostToDevice, c->stream));
    int rc = mcts_seq_search_device(f, (const hk_game_state*)d, fresh ? (const int*)(d + off[1]) : nullptr, iterations, seed,
                                    (hk_game_state*)(d + off[2]), (int*)(d + off[3]), (int*)(d + off[4]), (int*)(d + off[5]), c->stream, true, f->max_plies);
    if (rc) { cudaStreamSynchronize(c->stream); return rc; }
    std::vector<int> st((size_t)n);
    cudaError_t e = cudaMemcpyAsync(best_states, d + off[2], sz[2], cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(n_best, d + off[3], sz[3], cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && n_nodes) e = cudaMemcpyAsync(n_nodes, d + off[4], sz[4], cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(st.data(), d + off[5], sz[5], cudaMemcpyDeviceToHost, c->stream);
    cudaError_t e2 = cudaStreamSynchronize(c->stream);                 // also on the error path: nothing may stay in flight on the caller's buffers
    if (e == cudaSuccess) e = e2;
    if (e != cudaSuccess) { set_error("hk_mcts_forest_search: %s", cudaGetErrorString(e)); return HK_ERR_CUDA; }
    if (status) std::memcpy(status, st.data(), sizeof(int) * (size_t)n);
    for (int r = 0; r < n; ++r)
        if (st[r] == 1) { set_error("hk_mcts_forest_search: upNext() == -1 reached in tree %d (KartDiscreteGame.cs:326 would throw)", r); return HK_ERR_NO_UPNEXT; }
    return HK_OK;
}

extern "C" int hk_mcts_forest_nodes(const hk_mcts_forest* f, int tree, hk_mcts_node* nodes_out, int max_nodes, int32_t* n_nodes_out)
{
    if (!f || tree < 0 || tree >= f->n_trees || max_nodes < 0 || (max_nodes && !nodes_out)) { set_error("hk_mcts_forest_nodes: invalid argument"); return HK_ERR_INVALID_ARGUMENT; }
    ThreadCtx* c = ctx();
    if (!c) return HK_ERR_NO_DEVICE;
    SeqTree hdr;
    HK_CUDA(cudaMemcpy(&hdr, f->trees + tree, sizeof(SeqTree), cudaMemcpyDeviceToHost));
    if (n_nodes_out) *n_nodes_out = hdr.n_nodes;
    const int n = hdr.n_nodes < max_nodes ? hdr.n_nodes : max_nodes;
    if (n > 0) HK_CUDA(cudaMemcpy(nodes_out, f->slabs + (size_t)tree * f->max_nodes, sizeof(hk_mcts_node) * (size_t)n, cudaMemcpyDeviceToHost));
    return HK_OK;
}

extern "C" int hk_mcts_search_seq_batch(const hk_game* g, const hk_game_state* roots, int n_roots, int iterations, uint64_t seed,
                                        hk_game_state* best_states, int32_t* n_best, int32_t* root_gen, int32_t* root_episodes,
                                        float* root_values, int32_t* n_nodes)
{
    if (!g || !roots || n_roots < 1 || iterations < 0 || !best_states || !n_best) { set_error("hk_mcts_search_seq_batch: invalid argument"); return HK_ERR_INVALID_ARGUMENT; }
    long long plies = 1;
    for (int r = 0; r < n_roots; ++r) {
        int rc = check_state(g, &roots[r], "hk_mcts_search_seq_batch");
        if (rc) return rc;
        long long p = 0;
        for (int k = 0; k < roots[r].n_karts; ++k) { const long long dd = (long long)roots[r].finalSection - roots[r].karts[k].section; p += dd > 0 ? dd : 0; }
        if (p > plies) plies = p;
    }
    if (plies > HK_MAX_PLIES) { set_error("hk_mcts_search_seq_batch: a playout could take %lld plies (> %d)", plies, HK_MAX_PLIES); return HK_ERR_INVALID_ARGUMENT; }
    const long long max_nodes = 1 + (long long)iterations * plies;
    if (max_nodes > (1ll << 30)) { set_error("hk_mcts_search_seq_batch: tree too large"); return HK_ERR_INVALID_ARGUMENT; }
    hk_mcts_forest* f = nullptr;
    int rc = hk_mcts_forest_create(g, n_roots, (int)max_nodes, &f);
    if (rc) return rc;
    rc = hk_mcts_forest_search(f, roots, nullptr, iterations, seed, best_states, n_best, n_nodes, nullptr);
    if (rc == HK_OK && (root_gen || root_episodes || root_values)) {
        // the root's children in insertion order, gathered on the device (one thread per tree), three copies back
        const size_t cnt = (size_t)n_roots * HK_MAX_ACTIONS;
        int* d_gen = nullptr;
        cudaError_t e = cudaMalloc(&d_gen, cnt * 12);
        if (e == cudaSuccess) {
            int* d_ep = d_gen + cnt;
            float* d_val = reinterpret_cast<float*>(d_ep + cnt);
            count_launch();
            seq_root_children_kernel<<<(unsigned)((n_roots + 127) / 128), 128>>>(f->slabs, f->max_nodes, n_roots, d_gen, d_ep, d_val);
            e = cudaGetLastError();
            if (e == cudaSuccess && root_gen) e = cudaMemcpy(root_gen, d_gen, cnt * 4, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess && root_episodes) e = cudaMemcpy(root_episodes, d_ep, cnt * 4, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess && root_values) e = cudaMemcpy(root_values, d_val, cnt * 4, cudaMemcpyDeviceToHost);
            cudaFree(d_gen);
        }
        if (e != cudaSuccess) { cudaGetLastError(); set_error("hk_mcts_search_seq_batch: %s", cudaGetErrorString(e)); rc = e == cudaErrorMemoryAllocation ? HK_ERR_OUT_OF_MEMORY : HK_ERR_CUDA; }
    }
    hk_mcts_forest_destroy(f);
    return rc;
}

extern "C" int hk_mcts_rollouts_trace(const hk_game* g, const hk_game_state* leaf, int64_t n_rollouts, uint64_t seed, uint64_t rollout_offset,
                                      int32_t* n_plies_out, hk_action* actions_out, int32_t* choice_out, int32_t* n_scores_out, float* scores_out)
{
    if (!g || !leaf || n_rollouts < 0 || !n_plies_out || !actions_out || !choice_out || !n_scores_out || !scores_out) { set_error("hk_mcts_rollouts_trace: invalid argument"); return HK_ERR_INVALID_ARGUMENT; }
    int rc = check_state(g, leaf, "hk_mcts_rollouts_trace");
    if (rc) return rc;
    if (n_rollouts == 0) return HK_OK;
    ThreadCtx* c = ctx();
    if (!c) return HK_ERR_NO_DEVICE;
    const size_t n = (size_t)n_rollouts;
    const size_t sz[6] = {sizeof(hk_game_state), 4 * n, sizeof(hk_action) * n * HK_MAX_PLIES, 4 * n * HK_MAX_PLIES, 4 * n, 4 * n * 2 * HK_MAX_KARTS};
    size_t off[7]; off[0] = 0;
    for (int i = 0; i < 6; ++i) off[i + 1] = off[i] + ((sz[i] + 255) & ~(size_t)255);
    char* d = (char*)dscratch(c, 0, off[6]);
    if (!d) return HK_ERR_OUT_OF_MEMORY;
    HK_CUDA(cudaMemcpyAsync(d, leaf, sz[0], cudaMemcpyHostToDevice, c->stream));
    count_launch(); rollouts_trace_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(g->dev, (const hk_game_state*)d, n_rollouts, seed, rollout_offset,
        (int*)(d + off[1]), (hk_action*)(d + off[2]), (int*)(d + off[3]), (int*)(d + off[4]), (float*)(d + off[5]));
    HK_CUDA(cudaGetLastError());
    void* outs[5] = {n_plies_out, actions_out, choice_out, n_scores_out, scores_out};
    for (int i = 0; i < 5; ++i) HK_CUDA(cudaMemcpyAsync(outs[i], d + off[1 + i], sz[1 + i], cudaMemcpyDeviceToHost, c->stream));
    HK_CUDA(cudaStreamSynchronize(c->stream));
    return HK_OK;
}
