// hk_race.cu — the closed loop without PhysX for many independent 2-kart races, sm_100a.
//
// Per step and per agent: the LQNG problem recipe of HierarchicalKartAgent.SolveLQR, raycast-free branches
// (reference: Assets/Karting/Scripts/AI/HierarchicalKartAgent.cs:699-1197), device-side assembly + solve (hk_lqng.cu),
// the actuator map (:1206-1224) on the kinematic model the planners assume (MPC/KartMPCDynamics.cs:55-70) and the
// checkpoint bookkeeping of OnTriggerEnter (:611-662; lane by DiscretePositionTracker.CalculateLane :116-148); planFixed
// (:145-166) as the high level.  One thread per agent; kart records are 64-byte AoS (a warp reads 2 KB contiguous), plans are
// touched at two keys per agent and step.  Compiled with -fmad=false: the float32 expressions of the recipe (distances,
// Mathf.Atan2, Mathf.Pow weights) must round exactly like the reference's scalar code, an FMA would move avoid weights by
// one float ulp (6e-8), far above the 1e-9 parity bar.  Unity's Mathf.X(float) is (float)Math.X(double).
#include "hk_common.cuh"
#include "hk_game.cuh"
#include <cmath>
#include <vector>

namespace hk {

struct DevTrack {
    int n;
    hk_section sec[HK_MAX_SECTIONS];
    double trig[HK_MAX_SECTIONS][2];
    double fwd[HK_MAX_SECTIONS][2];
    double lane[HK_MAX_SECTIONS][4][2];
};

}  // namespace hk

struct hk_track {
    hk::DevTrack* dev;
    int n;
};

namespace hk {

__device__ __forceinline__ float mathf_atan2(float y, float x) { return (float)atan2((double)y, (double)x); }
__device__ __forceinline__ float magnitude2(float dx, float dz) { return (float)sqrt((double)(dx * dx + dz * dz)); }
__device__ __forceinline__ float wrap2pi_f(float a) { return a < 0 ? a + 2 * 3.14159274f : a; }
__device__ __forceinline__ double angle_difference(double a1, double a2) { return atan2(sin(a2 - a1), cos(a2 - a1)); }   // HKA:1341-1344
__device__ __forceinline__ bool is_straight(const DevTrack* t, int section) { return t->sec[section % t->n].insideR == 0.0f; }

__device__ __forceinline__ void plan_target(const DevTrack* t, const hk_race_params& p, const int8_t* lanes, const float* vels, int idx,
                                            double& x, double& z, double& vel)
{
    const double max_speed = (double)p.topSpeed;                    // GetMaxSpeed(), ArcadeKart.cs:210
    const int ln = lanes[idx];
    if (ln != 0) {
        x = t->lane[idx][ln - 1][0];
        z = t->lane[idx][ln - 1][1];
        const double v = (double)vels[idx] + (p.highModeMcts ? p.velocityBucketSize * 2 : 0);
        vel = max_speed < v ? max_speed : v;
    } else {
        x = t->trig[idx][0];
        z = t->trig[idx][1];
        vel = max_speed;
    }
}

// One thread per problem b = 2 race + ego; player 0 = ego, player 1 = the other kart (HKA:702).
// packed != nullptr: the description goes out as ONE 44-double record per problem (x0 | target | tw | cw | aw | otgt | otw, the layout of
// hk_lqng_assemble_solve_packed) that the solve kernel stages with one bulk copy; else into the seven arrays.
// cs != nullptr: also (cos h, sin h) of both players, [problem][2][2] — what lqng_trig_kernel would compute from the record (same double-precision
// functions; their code does not depend on this file's -fmad=false)
struct RecipeOut { double *x0, *target, *tw, *cw, *aw, *otgt, *otw, *cs; };   // bases such that base + b * width addresses problem b

__device__ __forceinline__ RecipeOut recipe_out(int b, double* x0_, double* target_, double* tw_, double* cw_, double* aw_, double* otgt_, double* otw_,
                                                double* packed, double* cs)
{
    double* const rec = packed ? packed + (size_t)b * 44 : nullptr;     // biased so that the indexing (base + b * width) lands in the record
    return RecipeOut{packed ? rec - (size_t)b * 8 : x0_, packed ? rec + 8 - (size_t)b * 8 : target_, packed ? rec + 16 - (size_t)b * 8 : tw_,
                     packed ? rec + 24 - (size_t)b * 2 : cw_, packed ? rec + 26 - (size_t)b * 4 : aw_, packed ? rec + 30 - (size_t)b * 8 : otgt_,
                     packed ? rec + 38 - (size_t)b * 6 : otw_, cs};
}

// player i of problem b, first part: state, target, own weights; returns the player's target point / speed for the other player's second part
__device__ __forceinline__ void race_recipe_player_a(const DevTrack* __restrict__ t, const hk_race_params& p, int b, int i, const hk_race_kart* karts,
                                                     const hk_race_plan* plans, const RecipeOut& o, double& tlx, double& tlz, double& vel)
{
    const int e = b & 1;
    const hk_race_kart* pair = karts + (b - e);
    const hk_race_plan* plan = plans + b;
    const hk_race_kart k = pair[i == 0 ? e : 1 - e];
    const int8_t* lanes = i == 0 ? plan->lane : plan->oppLane;       // own plan / belief about the other (:745-817)
    const float* vels = i == 0 ? plan->vel : plan->oppVel;
    double* x = o.x0 + (size_t)b * 8 + i * 4;
    x[0] = k.x; x[1] = k.z; x[2] = k.v; x[3] = k.h;                  // :730-736
    if (o.cs) { o.cs[(size_t)b * 4 + 2 * i] = cos(k.h); o.cs[(size_t)b * 4 + 2 * i + 1] = sin(k.h); }
    const int s = k.section + 1;                                     // :745
    const int idx = s % t->n, idx2 = (s + 1) % t->n;
    double nlx, nlz, nvel;
    plan_target(t, p, lanes, vels, idx, tlx, tlz, vel);
    plan_target(t, p, lanes, vels, idx2, nlx, nlz, nvel);
    const bool stopped = (float)k.v <= 5.0f;                         // :808
    double tx = tlx, tz = tlz, tv = stopped ? 0.0 : vel;
    const float d_t = magnitude2((float)(tlx - k.x), (float)(tlz - k.z));
    const bool near = d_t <= (is_straight(t, k.section) ? 10.5f : 7.5f);               // :823
    const float d_c = magnitude2((float)(t->trig[idx][0] - k.x), (float)(t->trig[idx][1] - k.z));
    const bool follow = near && (d_c <= 4.0f);                       // :877-890, centre-line distance stand-in
    const double h0 = k.h;
    double th;
    if (follow) {
        const double f6w = (double)wrap2pi_f(mathf_atan2((float)(nlz - k.z), (float)(nlx - k.x)));
        th = h0 - angle_difference(h0, f6w);                         // :887
        tx = nlx; tz = nlz;
        if (!stopped) tv = nvel;
    } else {
        const double f1w = (double)wrap2pi_f(mathf_atan2((float)(tlz - k.z), (float)(tlx - k.x)));
        if (near) {
            const double f2w = (double)wrap2pi_f(mathf_atan2((float)(nlz - tlz), (float)(nlx - tlx)));
            double blend = f1w - angle_difference(f2w, f1w) * (double)0.4f;             // :896
            if (blend < 0) blend += 2 * (double)3.14159274f;
            th = h0 - angle_difference(h0, blend);                   // :898
        } else {
            th = h0 - angle_difference(h0, f1w);                     // :921
        }
    }
    double* tg = o.target + (size_t)b * 8 + i * 4;
    tg[0] = tx; tg[1] = tz; tg[2] = tv; tg[3] = th;
    const double vmax1 = k.v > 1.0 ? k.v : 1.0;                      // own target weights, 2-agent branch (:930-962)
    const double w_xz = stopped ? 0.3 * 3.1 : 0.3 * 3.1 / vmax1;
    double* w = o.tw + (size_t)b * 8 + i * 4;
    w[0] = w_xz; w[1] = w_xz; w[2] = stopped ? -2.0 : 5e-4; w[3] = p.highModeMcts ? 3.5 : 1.9;
    o.cw[(size_t)b * 2 + i] = 0.115;                                 // :1192-1196
}

// player i of problem b, second part: avoid + opponent-target weights (:964-1190); (otx, otz, ovel) = the OTHER player's target from its first part
__device__ __forceinline__ void race_recipe_player_b(const hk_race_params& p, int b, int i, const hk_race_kart* karts, const RecipeOut& o, double otx,
                                                     double otz, double ovel)
{
    const int e = b & 1;
    const hk_race_kart* pair = karts + (b - e);
    const hk_race_kart ki = pair[i == 0 ? e : 1 - e];
    const hk_race_kart ko = pair[i == 0 ? 1 - e : e];
    const float mult = i == 0 ? (p.highModeMcts ? 1.0f : 0.45f) : 1.3f;   // k == this ? (Fixed ? 0.45f : 1.0f) : 1.3f, :999-1002
    const float dist = magnitude2((float)(ko.x - ki.x), (float)(ko.z - ki.z));
    const bool far = dist > 8 || !ko.active;                          // ... .magnitude > 8 || !o.is_active, :1010
    const float w32 = 1.0f / ((float)pow((double)dist, (double)1.5f) * mult);          // 1f/(Mathf.Pow(d,1.5f)*mult), :1019
    const double w = far ? 0.0 : (double)w32;
    o.aw[(size_t)b * 4 + i * 2 + 0] = w;
    o.aw[(size_t)b * 4 + i * 2 + 1] = w;
    double* og = o.otgt + (size_t)b * 8 + i * 4;
    og[0] = otx; og[1] = otz; og[2] = ovel; og[3] = 0.0;
    const double vmax1 = ki.v > 1.0 ? ki.v : 1.0;
    const double wxz = (p.highModeMcts ? 0.2 : 0.1) / vmax1;         // :1089-1091
    double* ow = o.otw + (size_t)b * 6 + i * 3;
    ow[0] = far ? 0.0 : wxz; ow[1] = far ? 0.0 : wxz; ow[2] = far ? 0.0 : 0.08;
}

__device__ __forceinline__ void race_recipe_body(const DevTrack* __restrict__ t, const hk_race_params& p, int b, const hk_race_kart* karts,
                                                 const hk_race_plan* plans, double* x0_, double* target_, double* tw_, double* cw_, double* aw_,
                                                 double* otgt_, double* otw_, double* packed, double* cs)
{
    const RecipeOut o = recipe_out(b, x0_, target_, tw_, cw_, aw_, otgt_, otw_, packed, cs);
    double tlx[2], tlz[2], vel[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) race_recipe_player_a(t, p, b, i, karts, plans, o, tlx[i], tlz[i], vel[i]);
#pragma unroll
    for (int i = 0; i < 2; ++i) race_recipe_player_b(p, b, i, karts, o, tlx[1 - i], tlz[1 - i], vel[1 - i]);
}

__global__ void race_recipe_kernel(const DevTrack* __restrict__ t, hk_race_params p, int n_problems, const hk_race_kart* __restrict__ karts,
                                   const hk_race_plan* __restrict__ plans, double* x0_, double* target_, double* tw_, double* cw_, double* aw_,
                                   double* otgt_, double* otw_, double* packed = nullptr, double* cs = nullptr)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_problems) return;
    race_recipe_body(t, p, b, karts, plans, x0_, target_, tw_, cw_, aw_, otgt_, otw_, packed, cs);
}

// planFixed (:145-166), one thread per agent
__global__ void race_plan_fixed_kernel(const DevTrack* __restrict__ t, hk_race_params p, int n_karts, const hk_race_kart* __restrict__ karts,
                                       hk_race_plan* plans)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_karts || !karts[k].active) return;
    const int s = karts[k].section;
    const int hi = s + p.treeSearchDepth < 1000 ? s + p.treeSearchDepth : 1000;
    for (int i = s + 1; i < hi + 1; ++i) {
        const int key = i % t->n;
        if (plans[k].lane[key] == 0) {
            plans[k].lane[key] = (int8_t)t->sec[(i - 1) % t->n].optimalLane;   // getOptimalNextLane
            plans[k].vel[key] = p.topSpeed;                                     // GetMaxSpeed()
        }
    }
}

// actuator map (:1206-1224) + kinematic plant (KartMPCDynamics.cs:55-70) + OnTriggerEnter bookkeeping (:611-662).
// u has `u_stride` doubles per kart (2: plain controls; 4: the LQNG u0 record of problem = kart, ego first).
__device__ __forceinline__ void race_step_body(const DevTrack* __restrict__ t, const hk_race_params& p, int i, int episode_step, const double* __restrict__ u,
                                               int u_stride, const int* __restrict__ lqng_status, unsigned long long* status_count, hk_race_kart* karts,
                                               hk_race_plan* plans, int* __restrict__ root_valid, int* __restrict__ cycles)
{
    if (lqng_status && lqng_status[i] != 0) atomicAdd(status_count, 1ull);
    hk_race_kart k = karts[i];
    if (!k.active) return;
    const double u0 = u[(size_t)i * u_stride], u1 = u[(size_t)i * u_stride + 1];
    const float max_ang = k.steer * 0.4f;                            // getMaxAngularVelocity, ArcadeKart.cs:505-510
    float ang = (float)u1;
    ang = ang < -max_ang ? -max_ang : (ang > max_ang ? max_ang : ang);              // Mathf.Clamp :1206
    bool accel = false, brake = false;
    if (u0 < 0) brake = true;
    else if (u0 > 0) accel = true;
    else ang = 0.0f;
    const float steering = ang / (0.4f * k.steer);                   // m_Steering :1224
    const float turning_power = steering * k.steer * (fabsf((float)k.v) > 0.5f ? 1.0f : 0.0f);   // ArcadeKart.cs:406
    const double omega = (double)(turning_power * 0.4f);
    const double x_old = k.x, z_old = k.z;
    k.x = x_old + p.dt * k.v * cos(k.h);
    k.z = z_old + p.dt * k.v * sin(k.h);
    // Unity's yaw is left-handed: a positive TurnInput turns the kart clockwise seen from above, i.e. the solver's heading
    // atan2(forward.z, forward.x) DEcreases.  SolveLQR compensates by mirroring the target heading about the current one
    // (h0 - AngleDifference(h0, target) = 2 h0 - target, :887,:898,:921), so the stand-in must turn like Unity does.
    double h = k.h - p.dt * omega;
    const double TWO_PI = 6.283185307179586;
    if (h < 0) h += TWO_PI;
    if (h >= TWO_PI) h -= TWO_PI;
    double v = k.v;
    if (accel) { v += p.dt * (double)p.accel; if (v > (double)p.topSpeed) v = (double)p.topSpeed; }
    else if (brake) { v -= p.dt * (double)p.braking; if (v < 0) v = 0; }
    else { v -= p.dt * (double)p.coastingDrag; if (v < 0) v = 0; }  // MoveTowards(v, 0, dt CoastingDrag), ArcadeKart.cs:431
    k.h = h; k.v = v;
    // did the kart enter the trigger of checkpoint section+1 during this step?
    const int index = k.section + 1, c = index % t->n;
    const double fx = t->fwd[c][0], fz = t->fwd[c][1];
    const double s_old = (x_old - t->trig[c][0]) * fx + (z_old - t->trig[c][1]) * fz;
    const double s_new = (k.x - t->trig[c][0]) * fx + (k.z - t->trig[c][1]) * fz;
    const double lat = -(k.x - t->trig[c][0]) * fz + (k.z - t->trig[c][1]) * fx;
    if (s_old < 0 && s_new >= 0 && fabs(lat) <= (double)p.gateHalfWidth) {
        int lane_new = 1;                                            // CalculateLane: first minimum
        float best = 0;
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const float d = magnitude2((float)(k.x - t->lane[c][l][0]), (float)(k.z - t->lane[c][l][1]));
            if (l == 0 || d < best) { best = d; lane_new = l + 1; }
        }
        hk_race_plan* pl = plans + i;
        const int pln = pl->lane[c];
        if (pln != 0) {                                              // KartAgent.cs:226-239, InitCheckpointIndex = 0
            const float d = magnitude2((float)(k.x - t->lane[c][pln - 1][0]), (float)(k.z - t->lane[c][pln - 1][1])) - 1.3f;
            pl->avgLaneDiff = ((d > 0.0f ? d : 0.0f) + pl->avgLaneDiff * (float)(index - 1)) / (float)index;
            pl->avgVelDiff = (((float)k.v - pl->vel[c]) + pl->avgVelDiff * (float)(index - 1)) / (float)index;
        }
        pl->lane[c] = 0;                                             // m_UpcomingLanes.Remove (:631-632)
        pl->vel[c] = 0.0f;
        pl->sectionTimes[c] = episode_step;                          // :650
        if (c == 0 && index / t->n >= 1 && index / t->n <= HK_MAX_LAPS) pl->lapStep[index / t->n - 1] = episode_step;
        const int dl = abs(k.lane - lane_new);
        const bool st_old = is_straight(t, k.section), st_new = is_straight(t, index);
        if (k.laneChanges + dl > p.maxLaneChanges && st_old) k.illegalLaneChanges += 1;   // :638-642
        if (st_old != st_new) k.laneChanges = 0;                     // :643-646
        else if (k.lane != lane_new) k.laneChanges += dl;            // :647-650
        k.section = index;
        k.lane = lane_new;
        k.sectionStep = episode_step;
        if (k.section == p.goalSection) k.active = 0;                // ReachGoalSection (:652-655)
        if (root_valid) { root_valid[i] = 0; cycles[i] = 0; }        // currentRoot = null; CyclesRootProcessed = 0 (:660-661)
    }
    karts[i] = k;
}

__global__ void race_step_kernel(const DevTrack* __restrict__ t, hk_race_params p, int n_karts, int episode_step, const double* __restrict__ u,
                                 int u_stride, const int* __restrict__ lqng_status, unsigned long long* status_count, hk_race_kart* karts,
                                 hk_race_plan* plans, int* __restrict__ root_valid = nullptr, int* __restrict__ cycles = nullptr)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_karts) return;
    race_step_body(t, p, i, episode_step, u, u_stride, lqng_status, status_count, karts, plans, root_valid, cycles);
}

// n_sub consecutive steps with the same (held) controls in one launch: the loops with a solve every lqr_every-th step (Duos) have nothing
// between two plant steps unless a planning event falls there.  The LQNG status is counted once, with the first step.
__global__ void race_steps_kernel(const DevTrack* __restrict__ t, hk_race_params p, int n_karts, int first_step, int n_sub, const double* __restrict__ u,
                                  int u_stride, const int* __restrict__ lqng_status, unsigned long long* status_count, hk_race_kart* karts,
                                  hk_race_plan* plans, int* __restrict__ root_valid, int* __restrict__ cycles)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_karts) return;
    for (int k = 0; k < n_sub; ++k)
        race_step_body(t, p, i, first_step + k, u, u_stride, k == 0 ? lqng_status : nullptr, status_count, karts, plans, root_valid, cycles);
}

// The step of one FixedUpdate and the problem description of the NEXT one in one launch (2-kart races; no planning event between the two):
// the two karts of a race are neighbouring lanes of a warp, so the partner's new state is there after a __syncwarp.  Two launches and the
// (cos h, sin h) kernel less per step of the loop.
__global__ void race_step_recipe_kernel(const DevTrack* __restrict__ t, hk_race_params p, int n_karts, int episode_step, const double* __restrict__ u,
                                        int u_stride, const int* __restrict__ lqng_status, unsigned long long* status_count, hk_race_kart* karts,
                                        hk_race_plan* plans, int* __restrict__ root_valid, int* __restrict__ cycles, double* packed, double* cs)
{
    // two threads per agent: thread (b, i) builds player i of agent b's problem (the float math of a player is a long dependent chain and
    // the batch is small); the four threads of a race are neighbouring lanes, thread (b, 0) also does agent b's plant step
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = gid >> 1, i = gid & 1;
    const bool live = b < n_karts;                                   // n_karts is even and blockDim a multiple of 4: races do not straddle warps
    if (live && i == 0) race_step_body(t, p, b, episode_step, u, u_stride, lqng_status, status_count, karts, plans, root_valid, cycles);
    __syncwarp();                                                    // both karts' new states are visible to the race's four threads
    double tlx = 0.0, tlz = 0.0, vel = 0.0;
    RecipeOut o{};
    if (live) {
        o = recipe_out(b, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, packed, cs);
        race_recipe_player_a(t, p, b, i, karts, plans, o, tlx, tlz, vel);
    }
    const double otx = __shfl_xor_sync(0xffffffffu, tlx, 1), otz = __shfl_xor_sync(0xffffffffu, tlz, 1), ovel = __shfl_xor_sync(0xffffffffu, vel, 1);
    if (live) race_recipe_player_b(p, b, i, karts, o, otx, otz, ovel);
}

static int check_track(const hk_track* t, const hk_race_params* p, const char* who)
{
    if (!t || !t->dev || !p) { set_error("%s: null track / params", who); return HK_ERR_INVALID_ARGUMENT; }
    if (p->planEvery <= 0 || p->horizon < 0 || p->horizon > HK_MAX_HORIZON || p->treeSearchDepth < 0) {
        set_error("%s: invalid parameters", who);
        return HK_ERR_INVALID_ARGUMENT;
    }
    return HK_OK;
}

}  // namespace hk

using namespace hk;

extern "C" int hk_track_create(const hk_section* sections, const double* trigger_xz, const double* forward_xz, const double* lane_xz,
                               int n_sections, hk_track** out)
{
    if (!sections || !trigger_xz || !forward_xz || !lane_xz || !out || n_sections < 1 || n_sections > HK_MAX_SECTIONS) {
        set_error("hk_track_create: invalid argument (1..%d sections)", HK_MAX_SECTIONS);
        return HK_ERR_INVALID_ARGUMENT;
    }
    ThreadCtx* c = ctx();
    if (!c) return HK_ERR_NO_DEVICE;
    DevTrack* h = new DevTrack();
    h->n = n_sections;
    for (int i = 0; i < n_sections; ++i) {
        h->sec[i] = sections[i];
        for (int a = 0; a < 2; ++a) { h->trig[i][a] = trigger_xz[i * 2 + a]; h->fwd[i][a] = forward_xz[i * 2 + a]; }
        for (int l = 0; l < 4; ++l)
            for (int a = 0; a < 2; ++a) h->lane[i][l][a] = lane_xz[(i * 4 + l) * 2 + a];
    }
    DevTrack* d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(DevTrack));
    if (e == cudaSuccess) e = cudaMemcpy(d, h, sizeof(DevTrack), cudaMemcpyHostToDevice);
    delete h;
    if (e != cudaSuccess) { set_error("hk_track_create: %s", cudaGetErrorString(e)); if (d) cudaFree(d); return HK_ERR_CUDA; }
    *out = new hk_track{d, n_sections};
    return HK_OK;
}

extern "C" void hk_track_destroy(hk_track* t)
{
    if (!t) return;
    if (t->dev) cudaFree(t->dev);
    delete t;
}

extern "C" int hk_race_recipe(const hk_track* t, const hk_race_params* p, int n_races, const hk_race_kart* karts, const hk_race_plan* plans,
                              double* x0, double* target, double* tw, double* cw, double* aw, double* otgt, double* otw)
{
    int rc = check_track(t, p, "hk_race_recipe");
    if (rc) return rc;
    if (n_races < 0 || (n_races > 0 && (!karts || !plans || !x0 || !target || !tw || !cw || !aw || !otgt || !otw))) {
        set_error("hk_race_recipe: invalid argument");
        return HK_ERR_INVALID_ARGUMENT;
    }
    if (n_races == 0) return HK_OK;
    ThreadCtx* c = ctx();
    if (!c) return HK_ERR_NO_DEVICE;
    const size_t nb = (size_t)2 * n_races;
    const size_t out_elems = nb * (8 + 8 + 8 + 2 + 4 + 8 + 6);
    char* d = (char*)dscratch(c, 8, nb * (sizeof(hk_race_kart) + sizeof(hk_race_plan)) + out_elems * sizeof(double));
    if (!d) return HK_ERR_OUT_OF_MEMORY;
    hk_race_kart* dk = (hk_race_kart*)d;
    hk_race_plan* dp = (hk_race_plan*)(dk + nb);
    double* o = (double*)(dp + nb);
    double *dx0 = o, *dtg = dx0 + nb * 8, *dtw = dtg + nb * 8, *dcw = dtw + nb * 8, *daw = dcw + nb * 2, *dot = daw + nb * 4, *dow = dot + nb * 8;
    HK_CUDA(cudaMemcpyAsync(dk, karts, nb * sizeof(hk_race_kart), cudaMemcpyHostToDevice, c->stream));
    HK_CUDA(cudaMemcpyAsync(dp, plans, nb * sizeof(hk_race_plan), cudaMemcpyHostToDevice, c->stream));
    count_launch();
    race_recipe_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, c->stream>>>(t->dev, *p, (int)nb, dk, dp, dx0, dtg, dtw, dcw, daw, dot, dow);
    HK_CUDA(cudaGetLastError());
    double* dst[7] = {x0, target, tw, cw, aw, otgt, otw};
    double* src[7] = {dx0, dtg, dtw, dcw, daw, dot, dow};
    const size_t per[7] = {8, 8, 8, 2, 4, 8, 6};
    for (int i = 0; i < 7; ++i) HK_CUDA(cudaMemcpyAsync(dst[i], src[i], nb * per[i] * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    HK_CUDA(cudaStreamSynchronize(c->stream));
    return HK_OK;
}

extern "C" int hk_race_plan_fixed(const hk_track* t, const hk_race_params* p, int n_karts, const hk_race_kart* karts, hk_race_plan* plans)
{
    int rc = check_track(t, p, "hk_race_plan_fixed");
    if (rc) return rc;
    if (n_karts < 0 || (n_karts > 0 && (!karts || !plans))) { set_error("hk_race_plan_fixed: invalid argument"); return HK_ERR_INVALID_ARGUMENT; }
    if (n_karts == 0) return HK_OK;
    ThreadCtx* c = ctx();
    if (!c) return HK_ERR_NO_DEVICE;
    char* d = (char*)dscratch(c, 8, (size_t)n_karts * (sizeof(hk_race_kart) + sizeof(hk_race_plan)));
    if (!d) return HK_ERR_OUT_OF_MEMORY;
    hk_race_kart* dk = (hk_race_kart*)d;
    hk_race_plan* dp = (hk_race_plan*)(dk + n_karts);
    HK_CUDA(cudaMemcpyAsync(dk, karts, (size_t)n_karts * sizeof(hk_race_kart), cudaMemcpyHostToDevice, c->stream));
    HK_CUDA(cudaMemcpyAsync(dp, plans, (size_t)n_karts * sizeof(hk_race_plan), cudaMemcpyHostToDevice, c->stream));
    count_launch();
    race_plan_fixed_kernel<<<(n_karts + 127) / 128, 128, 0, c->stream>>>(t->dev, *p, n_karts, dk, dp);
    HK_CUDA(cudaGetLastError());
    HK_CUDA(cudaMemcpyAsync(plans, dp, (size_t)n_karts * sizeof(hk_race_plan), cudaMemcpyDeviceToHost, c->stream));
    HK_CUDA(cudaStreamSynchronize(c->stream));
    return HK_OK;
}

extern "C" int hk_race_step(const hk_track* t, const hk_race_params* p, int n_karts, int episode_step, const double* u, hk_race_kart* karts,
                            hk_race_plan* plans)
{
    int rc = check_track(t, p, "hk_race_step");
    if (rc) return rc;
    if (n_karts < 0 || (n_karts > 0 && (!karts || !plans || !u))) { set_error("hk_race_step: invalid argument"); return HK_ERR_INVALID_ARGUMENT; }
    if (n_karts == 0) return HK_OK;
    ThreadCtx* c = ctx();
    if (!c) return HK_ERR_NO_DEVICE;
    char* d = (char*)dscratch(c, 8, (size_t)n_karts * (sizeof(hk_race_kart) + sizeof(hk_race_plan) + 2 * sizeof(double)));
    if (!d) return HK_ERR_OUT_OF_MEMORY;
    hk_race_kart* dk = (hk_race_kart*)d;
    hk_race_plan* dp = (hk_race_plan*)(dk + n_karts);
    double* du = (double*)(dp + n_karts);
    HK_CUDA(cudaMemcpyAsync(dk, karts, (size_t)n_karts * sizeof(hk_race_kart), cudaMemcpyHostToDevice, c->stream));
    HK_CUDA(cudaMemcpyAsync(dp, plans, (size_t)n_karts * sizeof(hk_race_plan), cudaMemcpyHostToDevice, c->stream));
    HK_CUDA(cudaMemcpyAsync(du, u, (size_t)n_karts * 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    count_launch();
    race_step_kernel<<<(n_karts + 127) / 128, 128, 0, c->stream>>>(t->dev, *p, n_karts, episode_step, du, 2, nullptr, nullptr, dk, dp);
    HK_CUDA(cudaGetLastError());
    HK_CUDA(cudaMemcpyAsync(karts, dk, (size_t)n_karts * sizeof(hk_race_kart), cudaMemcpyDeviceToHost, c->stream));
    HK_CUDA(cudaMemcpyAsync(plans, dp, (size_t)n_karts * sizeof(hk_race_plan), cudaMemcpyDeviceToHost, c->stream));
    HK_CUDA(cudaStreamSynchronize(c->stream));
    return HK_OK;
}

namespace hk {

// planWithMCTS's root state for every agent (HierarchicalKartAgent.cs:180-245): agent id = 2 race + ego; the karts within
// sectionWindow sections of the ego, all placed at the furthest one's section with the time they trail it by (float32 product, :211-214),
// velocity bucket (0, bucket) (quirk B.6-1), player 0 (B.6-2), tyre age from the steering stat.  nearby[id][i] = race-local agent of
// game kart i, -1 = none.  Same arithmetic as race.mcts_root_state / mcts_root_states_batch.
// With a planner (fresh != null) the kernel also takes planWithMCTS's decision (:175, :265): fresh[id] = 1 new tree (currentRoot == null),
// 0 continue the tree (CyclesRootProcessed < reuse_cycles), -1 no search (inactive agent, or the root has been reused enough); a
// continued tree keeps the root and the `nearby` map it was built with.
__global__ void race_mcts_root_kernel(const DevTrack* __restrict__ tr, hk_race_params p, int section_window, int time_precision, int n_agents,
                                      const hk_race_kart* __restrict__ karts, const hk_race_plan* __restrict__ plans,
                                      hk_game_state* __restrict__ roots, int* __restrict__ nearby,
                                      const int* __restrict__ root_valid = nullptr, const int* __restrict__ cycles = nullptr, int reuse_cycles = 0,
                                      int* __restrict__ fresh = nullptr, bool build_all = false)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_agents) return;
    if (fresh) {
        int f = 1;
        if (!karts[id].active) f = -1;                                       // !m_envController.inactiveAgents.Contains(this) (:331)
        else if (reuse_cycles > 0 && root_valid[id]) f = cycles[id] < reuse_cycles ? 0 : -1;   // :265 (reuse_cycles == 0: always a new tree)
        fresh[id] = f;
        if (f != 1 && !build_all) return;                                    // the leaf-parallel search runs over every root of the batch
    }
    const int r2 = id & ~1, e = id & 1, L = tr->n;
    const int sec0 = karts[r2].section, sec1 = karts[r2 + 1].section, sec_e = e ? sec1 : sec0, sec_o = e ? sec0 : sec1;
    const bool near = abs(sec_o - sec_e) < section_window;
    const int initial = near ? max(sec0, sec1) : sec_e;
    const int furthest = near ? (sec1 >= sec0 ? 1 : 0) : e;
    hk_game_state st;
    st.n_karts = near ? 2 : 1;
    st.initialSection = initial; st.lastCompletedSection = initial; st.finalSection = initial + p.treeSearchDepth;
    for (int slot = 0; slot < HK_MAX_KARTS; ++slot) {
        hk_kart_state ks = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        const bool valid = slot < 2 && (near || slot == 0);
        if (valid) {
            const int a = near ? slot : e;
            const hk_race_kart& k = karts[r2 + a];
            int t_at = 0;
            if (k.section != initial) {
                const int d = plans[r2 + a].sectionTimes[k.section % L] - plans[r2 + furthest].sectionTimes[k.section % L];
                t_at = (int)(((float)d * 0.02f) * (float)time_precision);
            }
            const float wear = (4.0f - k.steer) / 3.0f;                          // (maxSteer - steer) / (maxSteer - minSteer), ArcadeKart.cs:300-306
            ks.player = 0; ks.team = a; ks.section = initial; ks.timeAtSection = t_at; ks.min_velocity = 0;
            ks.max_velocity = min(p.velocityBucketSize, (int)p.topSpeed); ks.lane = k.lane; ks.tireAge = (int)(wear * 10000.0f);
            ks.laneChanges = k.laneChanges; ks.infeasible = 0;
            nearby[id * 2 + slot] = a;
        } else if (slot < 2) nearby[id * 2 + slot] = -1;
        st.karts[slot] = ks;
    }
    roots[id] = st;
}

// The waypoint hand-off of FixedUpdate (HierarchicalKartAgent.cs:366-402) from getBestStatesSequence: the ego's own lanes / velocities
// for sections beyond the next checkpoint, and its belief about the other kart's.  Same as race.apply_best_states(_batch).
// With a planner the kernel is the moment the background thread's result lands (:250-253, :271-273): only agents whose search was
// started (fresh >= 0) take part; bestStates is replaced, currentRoot is the searched tree, CyclesRootProcessed = 1 after a new tree
// and += 1 after a continued one (read at landing time: a checkpoint crossing in between has reset it).
__global__ void race_mcts_apply_kernel(const DevTrack* __restrict__ tr, int n_agents, const hk_race_kart* __restrict__ karts,
                                       const int* __restrict__ nearby, const hk_game_state* __restrict__ best, const int* __restrict__ n_best,
                                       hk_race_plan* __restrict__ plans, const int* __restrict__ fresh = nullptr, int* __restrict__ root_valid = nullptr,
                                       int* __restrict__ cycles = nullptr)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_agents) return;
    if (fresh) {
        if (fresh[id] < 0) return;
        root_valid[id] = 1;
        cycles[id] = fresh[id] ? 1 : cycles[id] + 1;
    }
    const int e = id & 1, L = tr->n, sec = karts[id].section, bound = sec + (sec == 0 ? 0 : 1);
    hk_race_plan& pl = plans[id];
    for (int k = 0; k < n_best[id]; ++k) {
        const hk_game_state& gs = best[(size_t)id * HK_MCTS_MAX_SEQ + k];
        for (int slot = 0; slot < gs.n_karts && slot < 2; ++slot) {
            const hk_kart_state& ks = gs.karts[slot];
            const int who = nearby[id * 2 + slot], key = ks.section % L;
            if (who == e) {
                if (ks.section > bound) { pl.lane[key] = (int8_t)ks.lane; pl.vel[key] = (float)ks.max_velocity; }
            } else if (who >= 0) { pl.oppLane[key] = (int8_t)ks.lane; pl.oppVel[key] = (float)ks.max_velocity; }
        }
    }
}

}  // namespace hk

// ---- the MCTS high level of many races: planner state that survives between hk_race_run_planned calls -------------------------------
struct hk_race_planner {
    const hk_game* game = nullptr;
    hk_race_mcts_params mp{};
    int n_agents = 0, karts_per_race = 2;
    bool n_layout = false;                // made by hk_raceN_planner_create: `nearby` has HK_MAX_KARTS entries per agent
    hk_mcts_forest* forest = nullptr;     // mode 0: every agent's currentRoot
    char* dev = nullptr;                  // one allocation: roots | best | nearby | n_best | fresh | root_valid | cycles | status
    hk_game_state *roots = nullptr, *best = nullptr;
    int *nearby = nullptr, *n_best = nullptr, *fresh = nullptr, *root_valid = nullptr, *cycles = nullptr, *status = nullptr;
    int pending_step = -1;                // step at which the result of the search in flight lands (-1: none)
};

extern "C" void hk_race_planner_destroy(hk_race_planner* pl)
{
    if (!pl) return;
    if (pl->forest) hk_mcts_forest_destroy(pl->forest);
    if (pl->dev) cudaFree(pl->dev);
    delete pl;
}

extern "C" int hk_race_planner_create(const hk_game* game, const hk_race_mcts_params* mp, int n_races, hk_race_planner** out)
{
    if (!game || !mp || !out || n_races < 1 || mp->iterations < 0 || mp->first_iterations < 0 || mp->reuse_cycles < 0 || mp->apply_delay < 0 ||
        (mp->mode != 0 && mp->mode != 1) || (mp->mode == 1 && mp->rollouts_per_leaf < 1)) {
        set_error("hk_race_planner_create: invalid argument (mode 0 | 1, budgets >= 0, rollouts_per_leaf >= 1 in mode 1)");
        return HK_ERR_INVALID_ARGUMENT;
    }
    if (game_karts_of(game) < 2) { set_error("hk_race_planner_create: the game must have at least 2 karts"); return HK_ERR_INVALID_ARGUMENT; }
    int rc = ensure_device();
    if (rc != HK_OK) return rc;
    hk_race_planner* pl = new hk_race_planner();
    pl->game = game; pl->mp = *mp; pl->n_agents = 2 * n_races;
    const size_t nb = (size_t)pl->n_agents;
    if (mp->mode == 0) {
        // a tree lives for one new search plus (normally) reuse_cycles - 1 continued ones — one more is reserved, see hk_race_mcts_params;
        // a playout of a 2-kart root takes <= 2 depth plies
        const long long big = mp->first_iterations > mp->iterations ? mp->first_iterations : mp->iterations;
        const long long life = big + (long long)mp->reuse_cycles * mp->iterations;
        const long long max_nodes = mp->max_tree_nodes > 0 ? mp->max_tree_nodes : 1 + life * 2 * game_params_of(game).treeSearchDepth;
        if (max_nodes > (1ll << 30) || 2 * game_params_of(game).treeSearchDepth > HK_MAX_PLIES) { delete pl; set_error("hk_race_planner_create: trees too large"); return HK_ERR_INVALID_ARGUMENT; }
        rc = hk_mcts_forest_create(game, pl->n_agents, (int)max_nodes, &pl->forest);
        if (rc) { delete pl; return rc; }
    }
    const size_t sz[8] = {sizeof(hk_game_state) * nb, sizeof(hk_game_state) * nb * HK_MCTS_MAX_SEQ, 4 * nb * 2, 4 * nb, 4 * nb, 4 * nb, 4 * nb, 4 * nb};
    size_t off[9]; off[0] = 0;
    for (int i = 0; i < 8; ++i) off[i + 1] = off[i] + ((sz[i] + 255) & ~(size_t)255);
    cudaError_t e = cudaMalloc(&pl->dev, off[8]);
    if (e == cudaSuccess) e = cudaMemset(pl->dev, 0, off[8]);
    if (e != cudaSuccess) { set_error("hk_race_planner_create: %s", cudaGetErrorString(e)); cudaGetLastError(); hk_race_planner_destroy(pl); return HK_ERR_OUT_OF_MEMORY; }
    pl->roots = (hk_game_state*)pl->dev; pl->best = (hk_game_state*)(pl->dev + off[1]); pl->nearby = (int*)(pl->dev + off[2]);
    pl->n_best = (int*)(pl->dev + off[3]); pl->fresh = (int*)(pl->dev + off[4]); pl->root_valid = (int*)(pl->dev + off[5]);
    pl->cycles = (int*)(pl->dev + off[6]); pl->status = (int*)(pl->dev + off[7]);
    *out = pl;
    return HK_OK;
}

// Synchronises every stream the calling thread's context owns: an error return must not leave copies in flight on the caller's buffers.
static void drain(ThreadCtx* c) { drain_ctx(c); }
#define HK_CUDA_DRAIN(call)                                                                             \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            hk::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);  \
            drain(c);                                                                                   \
            return e_ == cudaErrorMemoryAllocation ? HK_ERR_OUT_OF_MEMORY : HK_ERR_CUDA;                \
        }                                                                                               \
    } while (0)

// on_device: karts / plans (and u_last, then [agent][4]: the LQNG u0 record, ego controls first) are DEVICE pointers and `user_stream` the stream
// to work on (hk_race_run_device); else host arrays that are copied in and out around the loop.
static int race_run_impl(const hk_track* t, const hk_race_params* p, hk_race_planner* pl, int n_races, int first_step, int n_steps,
                         hk_race_kart* karts, hk_race_plan* plans, double* u_last, int64_t* lqng_status_nonzero, bool on_device = false,
                         void* user_stream = nullptr)
{
    int rc = check_track(t, p, "hk_race_run");
    if (rc) return rc;
    if (n_races < 0 || n_steps < 0 || first_step < 0 || (n_races > 0 && (!karts || !plans))) {
        set_error("hk_race_run: invalid argument");
        return HK_ERR_INVALID_ARGUMENT;
    }
    if (pl && (pl->n_layout || pl->n_agents != 2 * n_races || !p->highModeMcts || pl->mp.apply_delay >= p->planEvery)) {
        set_error("hk_race_run_planned: planner made for %d races, highModeMcts must be 1, apply_delay < planEvery", pl->n_agents / 2);
        return HK_ERR_INVALID_ARGUMENT;
    }
    if (lqng_status_nonzero) *lqng_status_nonzero = 0;
    if (n_races == 0 || n_steps == 0) return HK_OK;
    ThreadCtx* c = ctx();
    if (!c) return HK_ERR_NO_DEVICE;
    const size_t nb = (size_t)2 * n_races;                           // agents = problems per step
    const size_t compact = nb * (8 + 8 + 8 + 2 + 4 + 8 + 6);
    char* d = (char*)dscratch(c, 8, nb * (sizeof(hk_race_kart) + sizeof(hk_race_plan) + sizeof(int)) + (compact + nb * 8) * sizeof(double) + 64);
    if (!d) return HK_ERR_OUT_OF_MEMORY;
    hk_race_kart* dk = on_device ? karts : (hk_race_kart*)d;
    hk_race_plan* dp = on_device ? plans : (hk_race_plan*)((hk_race_kart*)d + nb);
    double* o = (double*)((hk_race_plan*)((hk_race_kart*)d + nb) + nb);
    double *dx0 = o, *dtg = dx0 + nb * 8, *dtw = dtg + nb * 8, *dcw = dtw + nb * 8, *daw = dcw + nb * 2, *dot = daw + nb * 4, *dow = dot + nb * 8;
    double* du = (on_device && u_last) ? u_last : dow + nb * 6;
    double* dcs = dow + nb * 6 + nb * 4;                                       // (cos h, sin h) of both players of every problem
    unsigned long long* dcount = (unsigned long long*)(dcs + nb * 4);
    int* dst = (int*)(dcount + 1);
    cudaStream_t s = (on_device && user_stream) ? (cudaStream_t)user_stream : c->stream;
    if (!on_device) {
        HK_CUDA_DRAIN(cudaMemcpyAsync(dk, karts, nb * sizeof(hk_race_kart), cudaMemcpyHostToDevice, s));
        HK_CUDA_DRAIN(cudaMemcpyAsync(dp, plans, nb * sizeof(hk_race_plan), cudaMemcpyHostToDevice, s));
    }
    HK_CUDA_DRAIN(cudaMemsetAsync(dcount, 0, sizeof(unsigned long long), s));
    const unsigned blocks = (unsigned)((nb + 127) / 128);
    int searches = 0;
    static const bool fuse = !(getenv("HK_RACE_FUSE") && atoi(getenv("HK_RACE_FUSE")) == 0);       // measurement knob
    bool have_recipe = false;                                        // the description of `step` was written by the previous step's launch
    for (int step = first_step; step < first_step + n_steps; ++step) {
        // HKA:331-353 (0.5 Hz): planFixed or planWithMCTS by mode; MCTS agents also plan when the episode begins (:85-93, T = 1.5)
        const bool replan = step > 0 && step % p->planEvery == 0;
        if (replan && !p->highModeMcts) {
            count_launch();
            race_plan_fixed_kernel<<<blocks, 128, 0, s>>>(t->dev, *p, (int)nb, dk, dp);
        } else if (pl) {                                                 // MCTS mode without a planner: the caller plans between two runs
            const hk_game_params& gp = game_params_of(pl->game);
            const hk_race_mcts_params& mp = pl->mp;
            const bool begin = step == 0 && mp.first_iterations > 0;
            if ((replan && step < gp.maxEpisodeSteps) || begin) {        // episodeSteps < maxEpisodeSteps (:331)
                const int budget = begin ? mp.first_iterations : mp.iterations;
                // tree of agent a searched by the event of step s: key = seed + (s / planEvery) n_agents + a — a function of the absolute
                // step, so that a loop advanced in blocks draws the streams of one long call
                const uint64_t seed = mp.seed + (uint64_t)(step / p->planEvery) * (uint64_t)nb;
                count_launch();
                race_mcts_root_kernel<<<blocks, 128, 0, s>>>(t->dev, *p, gp.sectionWindow, gp.timePrecision, (int)nb, dk, dp, pl->roots, pl->nearby,
                                                            pl->root_valid, pl->cycles, mp.mode == 0 ? mp.reuse_cycles : 0, pl->fresh, mp.mode == 1);
                HK_CUDA_DRAIN(cudaGetLastError());
                if (mp.mode == 0) rc = mcts_seq_search_device(pl->forest, pl->roots, pl->fresh, budget, seed, pl->best, pl->n_best, nullptr, pl->status, s, false,
                                                              pl->karts_per_race * gp.treeSearchDepth);
                else rc = mcts_search_device(pl->game, pl->roots, (int)nb, budget, mp.rollouts_per_leaf, seed, pl->best, pl->n_best, nullptr, nullptr, nullptr, pl->status, c, s);
                if (rc) { drain(c); return rc; }
                pl->pending_step = step + mp.apply_delay;
                ++searches;
            }
            if (pl->pending_step == step) {                              // the background thread's result lands (:250-253, :271-273), then FixedUpdate hands it off (:366-402)
                count_launch();
                race_mcts_apply_kernel<<<blocks, 128, 0, s>>>(t->dev, (int)nb, dk, pl->nearby, pl->best, pl->n_best, dp, pl->fresh, pl->root_valid, pl->cycles);
                HK_CUDA_DRAIN(cudaGetLastError());
                pl->pending_step = -1;
            }
        }
        if (!have_recipe) {
            count_launch();
            race_recipe_kernel<<<blocks, 128, 0, s>>>(t->dev, *p, (int)nb, dk, dp, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, dx0, dcs);
            HK_CUDA_DRAIN(cudaGetLastError());
        }
        rc = lqng_assemble_launch_packed((int)nb, 2, p->horizon, p->dt, dx0, du, dst, s, 9, nullptr, dcs);   // dx0: 44-double records (the seven arrays' space)
        if (rc) { drain(c); return rc; }
        // the next step's description rides with this step's plant update unless a planning event (planFixed, a search, a result landing)
        // comes between the two
        const int nx = step + 1;
        have_recipe = fuse && nx < first_step + n_steps && nx % p->planEvery != 0 && !(pl && pl->pending_step == nx);
        count_launch();
        if (have_recipe)
            race_step_recipe_kernel<<<(unsigned)((2 * nb + 127) / 128), 128, 0, s>>>(t->dev, *p, (int)nb, step, du, 4, dst, dcount, dk, dp, pl ? pl->root_valid : nullptr,
                                                          pl ? pl->cycles : nullptr, dx0, dcs);
        else
            race_step_kernel<<<blocks, 128, 0, s>>>(t->dev, *p, (int)nb, step, du, 4, dst, dcount, dk, dp, pl ? pl->root_valid : nullptr, pl ? pl->cycles : nullptr);
        HK_CUDA_DRAIN(cudaGetLastError());
    }
    if (!on_device) {
        HK_CUDA_DRAIN(cudaMemcpyAsync(karts, dk, nb * sizeof(hk_race_kart), cudaMemcpyDeviceToHost, s));
        HK_CUDA_DRAIN(cudaMemcpyAsync(plans, dp, nb * sizeof(hk_race_plan), cudaMemcpyDeviceToHost, s));
    }
    unsigned long long count = 0;
    HK_CUDA_DRAIN(cudaMemcpyAsync(&count, dcount, sizeof(count), cudaMemcpyDeviceToHost, s));
    if (u_last && !on_device) {
        // u0 records are [problem][4] (ego controls first): compact to [race][2 agents][2] on the host side of the copy
        double* hu = (double*)hscratch(c, 2, nb * 4 * sizeof(double));
        if (!hu) { drain(c); return HK_ERR_OUT_OF_MEMORY; }
        HK_CUDA_DRAIN(cudaMemcpyAsync(hu, du, nb * 4 * sizeof(double), cudaMemcpyDeviceToHost, s));
        HK_CUDA_DRAIN(cudaStreamSynchronize(s));
        for (size_t b = 0; b < nb; ++b) { u_last[b * 2] = hu[b * 4]; u_last[b * 2 + 1] = hu[b * 4 + 1]; }
    }
    std::vector<int> mst;
    if (pl && searches) { mst.resize(nb); HK_CUDA_DRAIN(cudaMemcpyAsync(mst.data(), pl->status, 4 * nb, cudaMemcpyDeviceToHost, s)); }
    HK_CUDA_DRAIN(cudaStreamSynchronize(s));                           // also for device state: the count and the trees' status come back
    if (lqng_status_nonzero) *lqng_status_nonzero = (int64_t)count;
    for (size_t a = 0; a < mst.size(); ++a)
        if (mst[a] == 1) { set_error("hk_race_run_planned: upNext() == -1 reached in the tree of agent %zu (KartDiscreteGame.cs:326 would throw)", a); return HK_ERR_NO_UPNEXT; }
    return HK_OK;
}

extern "C" int hk_race_run(const hk_track* t, const hk_race_params* p, int n_races, int first_step, int n_steps, hk_race_kart* karts,
                           hk_race_plan* plans, double* u_last, int64_t* lqng_status_nonzero)
{
    return race_run_impl(t, p, nullptr, n_races, first_step, n_steps, karts, plans, u_last, lqng_status_nonzero);
}

extern "C" int hk_race_run_device(const hk_track* t, const hk_race_params* p, hk_race_planner* planner, int n_races, int first_step, int n_steps,
                                  hk_race_kart* d_karts, hk_race_plan* d_plans, double* d_u, int64_t* lqng_status_nonzero, void* cuda_stream)
{
    return race_run_impl(t, p, planner, n_races, first_step, n_steps, d_karts, d_plans, d_u, lqng_status_nonzero, true, cuda_stream);
}

extern "C" int hk_race_run_planned(const hk_track* t, const hk_race_params* p, hk_race_planner* planner, int n_races, int first_step, int n_steps,
                                   hk_race_kart* karts, hk_race_plan* plans, double* u_last, int64_t* lqng_status_nonzero)
{
    if (!planner) { set_error("hk_race_run_planned: planner is NULL"); return HK_ERR_INVALID_ARGUMENT; }
    return race_run_impl(t, p, planner, n_races, first_step, n_steps, karts, plans, u_last, lqng_status_nonzero);
}

extern "C" int hk_race_planner_state(const hk_race_planner* planner, int32_t* root_valid, int32_t* cycles, int32_t* tree_status)
{
    if (planner && tree_status) HK_CUDA(cudaMemcpy(tree_status, planner->status, 4 * (size_t)planner->n_agents, cudaMemcpyDeviceToHost));
    if (!planner) { set_error("hk_race_planner_state: planner is NULL"); return HK_ERR_INVALID_ARGUMENT; }
    if (root_valid) HK_CUDA(cudaMemcpy(root_valid, planner->root_valid, 4 * (size_t)planner->n_agents, cudaMemcpyDeviceToHost));
    if (cycles) HK_CUDA(cudaMemcpy(cycles, planner->cycles, 4 * (size_t)planner->n_agents, cudaMemcpyDeviceToHost));
    return HK_OK;
}

extern "C" int hk_race_run_mcts(const hk_track* t, const hk_race_params* p, const hk_game* game, int iterations, int rollouts_per_leaf,
                                uint64_t seed, int n_races, int first_step, int n_steps, hk_race_kart* karts, hk_race_plan* plans,
                                double* u_last, int64_t* lqng_status_nonzero)
{
    if (!game || iterations < 0 || rollouts_per_leaf < 0 || !p || !p->highModeMcts) { set_error("hk_race_run_mcts: needs a game, highModeMcts = 1, rollouts_per_leaf >= 0"); return HK_ERR_INVALID_ARGUMENT; }
    if (n_races <= 0 || n_steps <= 0) return race_run_impl(t, p, nullptr, n_races, first_step, n_steps, karts, plans, u_last, lqng_status_nonzero);
    hk_race_mcts_params mp{};
    mp.mode = rollouts_per_leaf > 0 ? 1 : 0; mp.iterations = iterations; mp.first_iterations = 0; mp.rollouts_per_leaf = rollouts_per_leaf;
    mp.reuse_cycles = 0; mp.apply_delay = 0; mp.seed = seed; mp.max_tree_nodes = 0; mp.pad_ = 0;
    hk_race_planner* pl = nullptr;
    int rc = hk_race_planner_create(game, &mp, n_races, &pl);
    if (rc) return rc;
    rc = race_run_impl(t, p, pl, n_races, first_step, n_steps, karts, plans, u_last, lqng_status_nonzero);
    hk_race_planner_destroy(pl);
    return rc;
}

// =====================================================================================================================================
// Races with up to 4 karts and teams (hk_raceN_*, include/hk_abi.h): what SolveLQR / planWithMCTS do when the environment has more
// than two agents (HierarchicalKartAgent.cs:702-725, 930-1003, 1004-1190; 182-233).  One thread per agent id = K race + e.
// =====================================================================================================================================
namespace hk {

struct PlanView { const int8_t* lane; const float* vel; };

// Four threads per agent, one per player slot of the 4-player layout (dummy players zero, cw = 1): each repeats the cheap prelude (who takes
// part) and builds one player's description — the float math of a player (atan2f, pow) is a long dependent chain and the batch is small.
__global__ void raceN_recipe_kernel(const DevTrack* __restrict__ t, hk_race_params p, int K, int n_agents, const hk_race_kart* __restrict__ karts,
                                    const hk_race_plan* __restrict__ plans, const hk_race_belief* __restrict__ beliefs, int* __restrict__ n_players,
                                    int* __restrict__ players_out, double* x0, double* target, double* tw, double* cw, double* aw, double* otgt,
                                    double* otw, double* rec2 = nullptr, double* cs2 = nullptr, int* __restrict__ n2 = nullptr)
{
    // rec2 (the loop): a game of one or two players is written straight as the 44-double record of the 2-kart kernel (+ its (cos h, sin h)
    // pairs, (0, 0) marking a dummy player) and NOT in the 4-player layout; a game of three or four gets an all-dummy record there (the
    // 2-kart launch covers every slot) and the 4-player layout as before.  n2 = players of the record (0 for the all-dummy one).
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int id = gid >> 2, i = gid & 3;
    if (id >= n_agents) return;
    const int e = id % K;
    const hk_race_kart* rk = karts + (id - e);
    const bool fixed = !p.highModeMcts;
    // [this] + teamAgents + otherAgents in environment order (:702), kept if within 8 m of the ego when the environment has > 2 agents (:709-721)
    int all[HK_MAX_KARTS], na = 0;
    all[na++] = e;
    for (int a = 0; a < K; ++a) if (a != e && rk[a].team == rk[e].team) all[na++] = a;
    for (int a = 0; a < K; ++a) if (rk[a].team != rk[e].team) all[na++] = a;
    int act[HK_MAX_KARTS], N = 0, nearby = -1;
    if (K > 2) {
        for (int i = 0; i < na; ++i) {
            const hk_race_kart& k = rk[all[i]];
            if (magnitude2((float)(k.x - rk[e].x), (float)(k.z - rk[e].z)) < 8) { nearby += 1; act[N++] = all[i]; }
        }
    } else {
        for (int i = 0; i < na; ++i) act[N++] = all[i];
    }
    nearby = nearby > 1 ? nearby : 1;                                    // Math.Max(nearbyAgents, 1) :726
    if (i == 0) {
        n_players[id] = N;
        for (int j = 0; j < HK_MAX_KARTS; ++j) players_out[id * 4 + j] = j < N ? act[j] : -1;
        if (n2) n2[id] = N <= 2 ? N : 0;
    }
    const bool small = rec2 != nullptr && N <= 2;
    if (rec2 && i < 2 && (N > 2 || i >= N)) {                            // dummy player i of the 2-kart record: zero weights, control weight 1
        double* r = rec2 + (size_t)id * 44;
        for (int c = 0; c < 4; ++c) { r[i * 4 + c] = 0.0; r[8 + i * 4 + c] = 0.0; r[16 + i * 4 + c] = 0.0; r[30 + i * 4 + c] = 0.0; }
        r[24 + i] = 1.0; r[26 + i * 2] = 0.0; r[26 + i * 2 + 1] = 0.0;
        for (int c = 0; c < 3; ++c) r[38 + i * 3 + c] = 0.0;
        cs2[(size_t)id * 4 + 2 * i] = 0.0; cs2[(size_t)id * 4 + 2 * i + 1] = 0.0;
    }
    if (small && i >= N) return;                                         // nothing of this game goes to the 4-player layout
    unsigned in_game = 0;
    for (int i = 0; i < N; ++i) in_game |= 1u << act[i];
    const double max_speed = (double)p.topSpeed;
    auto view = [&](int k) -> PlanView {                                 // own plan (k == this) or the ego's belief about kart k
        if (k == e) return PlanView{plans[id].lane, plans[id].vel};
        const hk_race_belief& b = beliefs[(size_t)id * K + k];
        return PlanView{b.lane, b.vel};
    };
    auto plan_xzv = [&](int k, int idx, double& x, double& z, double& v) {
        const PlanView pv = view(k);
        const int ln = pv.lane[idx];
        if (ln != 0) {
            x = t->lane[idx][ln - 1][0]; z = t->lane[idx][ln - 1][1];
            const double vv = (double)pv.vel[idx] + (p.highModeMcts ? p.velocityBucketSize * 2 : 0);
            v = max_speed < vv ? max_speed : vv;
        } else { x = t->trig[idx][0]; z = t->trig[idx][1]; v = max_speed; }
    };
    {
        const size_t o4 = ((size_t)id * 4 + i) * 4;
        double* r2 = small ? rec2 + (size_t)id * 44 : nullptr;
        double* xo = small ? r2 + i * 4 : x0 + o4; double* tg = small ? r2 + 8 + i * 4 : target + o4; double* w = small ? r2 + 16 + i * 4 : tw + o4;
        double* awi = small ? r2 + 26 + i * 2 : aw + ((size_t)id * 4 + i) * 6;
        double* ogi = small ? r2 + 30 + i * 4 : otgt + ((size_t)id * 4 + i) * 12;
        double* owi = small ? r2 + 38 + i * 3 : otw + ((size_t)id * 4 + i) * 9;
        double* cwp = small ? r2 + 24 + i : cw + (size_t)id * 4 + i;
        for (int c = 0; c < (small ? 2 : 6); ++c) awi[c] = 0.0;           // a 2-kart record has one private slot per player
        for (int c = 0; c < (small ? 4 : 12); ++c) ogi[c] = 0.0;
        for (int c = 0; c < (small ? 3 : 9); ++c) owi[c] = 0.0;
        if (i >= N) {                                                    // dummy player
            for (int c = 0; c < 4; ++c) { xo[c] = 0.0; tg[c] = 0.0; w[c] = 0.0; }
            *cwp = 1.0;
            return;
        }
        const int kI = act[i];
        const hk_race_kart k = rk[kI];
        xo[0] = k.x; xo[1] = k.z; xo[2] = k.v; xo[3] = k.h;              // :730-736
        if (small) { cs2[(size_t)id * 4 + 2 * i] = cos(k.h); cs2[(size_t)id * 4 + 2 * i + 1] = sin(k.h); }
        const int s = k.section + 1;                                     // :745
        const int idx = s % t->n, idx2 = (s + 1) % t->n;
        double lx, lz, vel, nlx, nlz, nvel;
        plan_xzv(kI, idx, lx, lz, vel);
        plan_xzv(kI, idx2, nlx, nlz, nvel);
        const bool stopped = (float)k.v <= 5.0f;                         // :808
        double tx = lx, tz = lz, tv = stopped ? 0.0 : vel;
        const float d_t = magnitude2((float)(lx - k.x), (float)(lz - k.z));
        const bool near = d_t <= (is_straight(t, k.section) ? 10.5f : 7.5f);           // :823
        const float d_c = magnitude2((float)(t->trig[idx][0] - k.x), (float)(t->trig[idx][1] - k.z));
        const bool follow = near && (d_c <= 4.0f);                       // :877-890, centre-line distance stand-in
        const double h0 = k.h;
        double th;
        if (follow) {
            const double f6w = (double)wrap2pi_f(mathf_atan2((float)(nlz - k.z), (float)(nlx - k.x)));
            th = h0 - angle_difference(h0, f6w);                         // :887
            tx = nlx; tz = nlz;
            if (!stopped) tv = nvel;
        } else {
            const double f1w = (double)wrap2pi_f(mathf_atan2((float)(lz - k.z), (float)(lx - k.x)));
            if (near) {
                const double f2w = (double)wrap2pi_f(mathf_atan2((float)(nlz - lz), (float)(nlx - lx)));
                double blend = f1w - angle_difference(f2w, f1w) * (double)0.4f;         // :896
                if (blend < 0) blend += 2 * (double)3.14159274f;
                th = h0 - angle_difference(h0, blend);                   // :898
            } else {
                th = h0 - angle_difference(h0, f1w);                     // :921
            }
        }
        tg[0] = tx; tg[1] = tz; tg[2] = tv; tg[3] = th;
        const double vmax1 = k.v > 1.0 ? k.v : 1.0;                      // own target weights :928-962
        w[3] = N > 2 ? (fixed ? 2.5 : 3.5) * nearby : (fixed ? 1.9 : 3.5);
        const double wxz = stopped ? nearby * 0.3 * 3.1 : nearby * 0.3 * 3.1 / vmax1;
        w[0] = wxz; w[1] = wxz; w[2] = stopped ? (double)(nearby * -2) : nearby * 5e-4;
        *cwp = N > 2 ? (fixed ? 0.135 : 0.25) : 0.115;                                  // :1192-1196
        float mult;                                                      // :977-1003
        if (K > 2 && N > 2) mult = kI == e ? (fixed ? 0.55f : 1.0f) / nearby : 1.7f / nearby;
        else mult = kI == e ? (fixed ? 0.45f : 1.0f) : 1.3f;
        int slot = 0, nearby_opponents = 0;
        for (int pass = 0; pass < 2; ++pass) {                           // player k's otherAgents (:1004-1096), then its teamAgents (:1099-1190)
            for (int o = 0; o < K; ++o) {
                const bool mate = rk[o].team == k.team;
                if (o == kI ? true : (pass == 0 ? mate : !mate)) continue;
                if (!((in_game >> o) & 1u)) continue;                    // !actualAllPlayers.Contains(o)
                const hk_race_kart ko = rk[o];
                const float dist = magnitude2((float)(ko.x - k.x), (float)(ko.z - k.z));
                const bool off = dist > 8 || !ko.active;
                const float m_eff = pass == 0 ? mult : mult / 2.0f;       // multiplier2 :1113
                const float w32 = 1.0f / ((float)pow((double)dist, (double)1.5f) * m_eff);
                const double wa = off ? 0.0 : (double)w32;
                awi[slot * 2] = wa; awi[slot * 2 + 1] = wa;
                if (pass == 0 && !off) nearby_opponents += 1;
                double ox, oz, ov;
                plan_xzv(o, (ko.section + 1) % t->n, ox, oz, ov);
                ogi[slot * 4] = ox; ogi[slot * 4 + 1] = oz; ogi[slot * 4 + 2] = pass == 0 ? ov : max_speed; ogi[slot * 4 + 3] = 0.0;
                double oxz = 0.0, ovw = 0.0;
                if (pass == 0) {
                    if (!off) {
                        if (N > 2) { oxz = (fixed ? 0.1 : 0.2) / (vmax1 * nearby); ovw = 0.08 / nearby; }   // :1083-1085
                        else { oxz = (fixed ? 0.1 : 0.2) / vmax1; ovw = 0.08; }                            // :1089-1091
                    }
                } else if (!off && nearby_opponents >= 1) {
                    if (N > 2) oxz = -(fixed ? 0.0 : 3e-5) / (vmax1 * nearby);                             // :1178-1180
                    else oxz = -(fixed ? 1e-4 : 2e-4) / vmax1;                                             // :1184-1186
                }
                owi[slot * 3] = oxz; owi[slot * 3 + 1] = oxz; owi[slot * 3 + 2] = ovw;
                ++slot;
            }
        }
    }
}

// planWithMCTS's root state (:180-245) for K-kart races: every agent of the race within sectionWindow sections of the ego, in
// environment order, placed at the furthest one's section; team = getTeamID.  nearby [id][HK_MAX_KARTS].
__global__ void raceN_mcts_root_kernel(const DevTrack* __restrict__ tr, hk_race_params p, int K, int section_window, int time_precision, int n_agents,
                                       const hk_race_kart* __restrict__ karts, const hk_race_plan* __restrict__ plans,
                                       hk_game_state* __restrict__ roots, int* __restrict__ nearby, const int* __restrict__ root_valid,
                                       const int* __restrict__ cycles, int reuse_cycles, int* __restrict__ fresh, bool build_all)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_agents) return;
    int f = 1;
    if (!karts[id].active) f = -1;
    else if (reuse_cycles > 0 && root_valid[id]) f = cycles[id] < reuse_cycles ? 0 : -1;
    fresh[id] = f;
    if (f != 1 && !build_all) return;
    const int e = id % K, r0 = id - e, L = tr->n;
    const int me_sec = karts[id].section;
    int initial = me_sec, furthest = e, n = 0, near[HK_MAX_KARTS];
    for (int a = 0; a < K; ++a)                                          // :182-193
        if (abs(karts[r0 + a].section - me_sec) < section_window) {
            near[n++] = a;
            if (karts[r0 + a].section > initial) initial = karts[r0 + a].section;
            if (initial == karts[r0 + a].section) furthest = a;
        }
    hk_game_state st;
    st.n_karts = n; st.initialSection = initial; st.lastCompletedSection = initial; st.finalSection = initial + p.treeSearchDepth;
    for (int slot = 0; slot < HK_MAX_KARTS; ++slot) {
        hk_kart_state ks = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (slot < n) {
            const int a = near[slot];
            const hk_race_kart& k = karts[r0 + a];
            int t_at = 0;
            if (k.section != initial) {
                const int d = plans[r0 + a].sectionTimes[k.section % L] - plans[r0 + furthest].sectionTimes[k.section % L];
                t_at = (int)(((float)d * 0.02f) * (float)time_precision);
            }
            const float wear = (4.0f - k.steer) / 3.0f;
            ks.player = 0; ks.team = k.team; ks.section = initial; ks.timeAtSection = t_at; ks.min_velocity = 0;
            ks.max_velocity = min(p.velocityBucketSize, (int)p.topSpeed); ks.lane = k.lane; ks.tireAge = (int)(wear * 10000.0f);
            ks.laneChanges = k.laneChanges; ks.infeasible = 0;
            nearby[id * HK_MAX_KARTS + slot] = a;
        } else nearby[id * HK_MAX_KARTS + slot] = -1;
        st.karts[slot] = ks;
    }
    roots[id] = st;
}

// the waypoint hand-off (:366-402) for K-kart races: own lanes / velocities, and the belief tables about every other kart of the game
__global__ void raceN_mcts_apply_kernel(const DevTrack* __restrict__ tr, int K, int n_agents, const hk_race_kart* __restrict__ karts,
                                        const int* __restrict__ nearby, const hk_game_state* __restrict__ best, const int* __restrict__ n_best,
                                        hk_race_plan* __restrict__ plans, hk_race_belief* __restrict__ beliefs, const int* __restrict__ fresh,
                                        int* __restrict__ root_valid, int* __restrict__ cycles)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_agents) return;
    if (fresh[id] < 0) return;
    root_valid[id] = 1;
    cycles[id] = fresh[id] ? 1 : cycles[id] + 1;
    const int e = id % K, L = tr->n, sec = karts[id].section, bound = sec + (sec == 0 ? 0 : 1);
    hk_race_plan& pl = plans[id];
    for (int k = 0; k < n_best[id]; ++k) {
        const hk_game_state& gs = best[(size_t)id * HK_MCTS_MAX_SEQ + k];
        for (int slot = 0; slot < gs.n_karts && slot < HK_MAX_KARTS; ++slot) {
            const hk_kart_state& ks = gs.karts[slot];
            const int who = nearby[id * HK_MAX_KARTS + slot], key = ks.section % L;
            if (who == e) {
                if (ks.section > bound) { pl.lane[key] = (int8_t)ks.lane; pl.vel[key] = (float)ks.max_velocity; }
            } else if (who >= 0) {
                hk_race_belief& b = beliefs[(size_t)id * K + who];
                b.lane[key] = (int8_t)ks.lane; b.vel[key] = (float)ks.max_velocity;
            }
        }
    }
}

}  // namespace hk

static int raceN_check(const hk_track* t, const hk_race_params* p, int K, int n_races, const char* who)
{
    int rc = check_track(t, p, who);
    if (rc) return rc;
    if (K < 2 || K > HK_MAX_KARTS || n_races < 0) { set_error("%s: karts_per_race must be 2..%d", who, HK_MAX_KARTS); return HK_ERR_INVALID_ARGUMENT; }
    return HK_OK;
}

extern "C" int hk_raceN_recipe(const hk_track* t, const hk_race_params* p, int K, int n_races, const hk_race_kart* karts, const hk_race_plan* plans,
                               const hk_race_belief* beliefs, int32_t* n_players, int32_t* players, double* x0, double* target, double* tw,
                               double* cw, double* aw, double* otgt, double* otw)
{
    int rc = raceN_check(t, p, K, n_races, "hk_raceN_recipe");
    if (rc) return rc;
    if (n_races > 0 && (!karts || !plans || !beliefs || !n_players || !players || !x0 || !target || !tw || !cw || !aw || !otgt || !otw)) {
        set_error("hk_raceN_recipe: invalid argument");
        return HK_ERR_INVALID_ARGUMENT;
    }
    if (n_races == 0) return HK_OK;
    ThreadCtx* c = ctx();
    if (!c) return HK_ERR_NO_DEVICE;
    const size_t nb = (size_t)K * n_races;
    const size_t per[7] = {16, 16, 16, 4, 24, 48, 36};
    size_t out_elems = 0;
    for (size_t v : per) out_elems += v;
    const size_t in_bytes = nb * (sizeof(hk_race_kart) + sizeof(hk_race_plan)) + nb * K * sizeof(hk_race_belief);
    char* d = (char*)dscratch(c, 8, in_bytes + nb * (5 * sizeof(int) + out_elems * sizeof(double)) + 256);
    if (!d) return HK_ERR_OUT_OF_MEMORY;
    hk_race_kart* dk = (hk_race_kart*)d;
    hk_race_plan* dp = (hk_race_plan*)(dk + nb);
    hk_race_belief* db = (hk_race_belief*)(dp + nb);
    double* o = (double*)(((uintptr_t)(db + nb * K) + 15) & ~(uintptr_t)15);
    double* dout[7];
    for (int i = 0; i < 7; ++i) { dout[i] = o; o += nb * per[i]; }
    int* dn = (int*)o; int* dpl = dn + nb;
    cudaStream_t s = c->stream;
    HK_CUDA_DRAIN(cudaMemcpyAsync(dk, karts, nb * sizeof(hk_race_kart), cudaMemcpyHostToDevice, s));
    HK_CUDA_DRAIN(cudaMemcpyAsync(dp, plans, nb * sizeof(hk_race_plan), cudaMemcpyHostToDevice, s));
    HK_CUDA_DRAIN(cudaMemcpyAsync(db, beliefs, nb * K * sizeof(hk_race_belief), cudaMemcpyHostToDevice, s));
    count_launch();
    raceN_recipe_kernel<<<(unsigned)((nb * 4 + 127) / 128), 128, 0, s>>>(t->dev, *p, K, (int)nb, dk, dp, db, dn, dpl, dout[0], dout[1], dout[2], dout[3],
                                                                   dout[4], dout[5], dout[6]);
    HK_CUDA_DRAIN(cudaGetLastError());
    double* hout[7] = {x0, target, tw, cw, aw, otgt, otw};
    for (int i = 0; i < 7; ++i) HK_CUDA_DRAIN(cudaMemcpyAsync(hout[i], dout[i], nb * per[i] * sizeof(double), cudaMemcpyDeviceToHost, s));
    HK_CUDA_DRAIN(cudaMemcpyAsync(n_players, dn, nb * sizeof(int), cudaMemcpyDeviceToHost, s));
    HK_CUDA_DRAIN(cudaMemcpyAsync(players, dpl, nb * 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
    HK_CUDA_DRAIN(cudaStreamSynchronize(s));
    return HK_OK;
}

// Games of one or two players (after the 8 m filter most are) do not need the 4-player frame: raceN_recipe_kernel writes them as the
// 44-double record of the 2-kart kernel — player 1 of a one-player game a decoupled dummy (zero weights, control weight 1, (cos h, sin h) =
// (0, 0): the kernel assembles A = I, B = 0) — and lqng_mma4_kernel is gated to the games of three and four.  A game of three or four
// gets an all-dummy record there (the 2-kart launch covers every slot; its answer for those slots is not used).
__global__ void raceN_merge2_kernel(int n_agents, const int* __restrict__ n_players, const double* __restrict__ u2, const int* __restrict__ st2,
                                    double* __restrict__ u, int* __restrict__ st)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_agents || n_players[b] > 2) return;
    u[(size_t)b * 8] = u2[(size_t)b * 4]; u[(size_t)b * 8 + 1] = u2[(size_t)b * 4 + 1];
    st[b] = st2[b];
}

extern "C" int hk_raceN_planner_create(const hk_game* game, const hk_race_mcts_params* mp, int K, int n_races, hk_race_planner** out)
{
    if (!game || !mp || !out || n_races < 1 || K < 2 || K > HK_MAX_KARTS || mp->iterations < 0 || mp->first_iterations < 0 || mp->reuse_cycles < 0 ||
        mp->apply_delay < 0 || (mp->mode != 0 && mp->mode != 1) || (mp->mode == 1 && mp->rollouts_per_leaf < 1)) {
        set_error("hk_raceN_planner_create: invalid argument");
        return HK_ERR_INVALID_ARGUMENT;
    }
    if (game_karts_of(game) < K) { set_error("hk_raceN_planner_create: the game has fewer karts than a race"); return HK_ERR_INVALID_ARGUMENT; }
    int rc = ensure_device();
    if (rc != HK_OK) return rc;
    hk_race_planner* pl = new hk_race_planner();
    pl->game = game; pl->mp = *mp; pl->n_agents = K * n_races; pl->karts_per_race = K; pl->n_layout = true;
    const size_t nb = (size_t)pl->n_agents;
    const int depth = game_params_of(game).treeSearchDepth;
    if (mp->mode == 0) {
        const long long big = mp->first_iterations > mp->iterations ? mp->first_iterations : mp->iterations;
        const long long life = big + (long long)mp->reuse_cycles * mp->iterations;
        const long long max_nodes = mp->max_tree_nodes > 0 ? mp->max_tree_nodes : 1 + life * K * depth;
        if (max_nodes > (1ll << 30) || K * depth > HK_MAX_PLIES) { delete pl; set_error("hk_raceN_planner_create: trees too large"); return HK_ERR_INVALID_ARGUMENT; }
        rc = hk_mcts_forest_create(game, pl->n_agents, (int)max_nodes, &pl->forest);
        if (rc) { delete pl; return rc; }
    }
    const size_t sz[8] = {sizeof(hk_game_state) * nb, sizeof(hk_game_state) * nb * HK_MCTS_MAX_SEQ, 4 * nb * HK_MAX_KARTS, 4 * nb, 4 * nb, 4 * nb, 4 * nb, 4 * nb};
    size_t off[9]; off[0] = 0;
    for (int i = 0; i < 8; ++i) off[i + 1] = off[i] + ((sz[i] + 255) & ~(size_t)255);
    cudaError_t e = cudaMalloc(&pl->dev, off[8]);
    if (e == cudaSuccess) e = cudaMemset(pl->dev, 0, off[8]);
    if (e != cudaSuccess) { set_error("hk_raceN_planner_create: %s", cudaGetErrorString(e)); cudaGetLastError(); hk_race_planner_destroy(pl); return HK_ERR_OUT_OF_MEMORY; }
    pl->roots = (hk_game_state*)pl->dev; pl->best = (hk_game_state*)(pl->dev + off[1]); pl->nearby = (int*)(pl->dev + off[2]);
    pl->n_best = (int*)(pl->dev + off[3]); pl->fresh = (int*)(pl->dev + off[4]); pl->root_valid = (int*)(pl->dev + off[5]);
    pl->cycles = (int*)(pl->dev + off[6]); pl->status = (int*)(pl->dev + off[7]);
    *out = pl;
    return HK_OK;
}

// on_device: karts / plans / beliefs are DEVICE pointers, u_hold the device array [agent][8] (the 4-player u0 record, ego controls first) and
// `user_stream` the stream to work on (hk_raceN_run_device); else host arrays (u_hold [agent][2]) copied in and out around the loop.
static int raceN_run_impl(const hk_track* t, const hk_race_params* p, hk_race_planner* pl, int K, int lqr_every, int n_races, int first_step,
                          int n_steps, hk_race_kart* karts, hk_race_plan* plans, hk_race_belief* beliefs, double* u_hold,
                          int64_t* lqng_status_nonzero, bool on_device, void* user_stream)
{
    int rc = raceN_check(t, p, K, n_races, "hk_raceN_run");
    if (rc) return rc;
    if (n_steps < 0 || first_step < 0 || lqr_every < 1 || (n_races > 0 && (!karts || !plans || !beliefs || !u_hold))) { set_error("hk_raceN_run: invalid argument"); return HK_ERR_INVALID_ARGUMENT; }
    if (pl && (!pl->n_layout || pl->n_agents != K * n_races || pl->karts_per_race != K || !p->highModeMcts || pl->mp.apply_delay >= p->planEvery)) {
        set_error("hk_raceN_run: planner made for another batch, highModeMcts must be 1, apply_delay < planEvery");
        return HK_ERR_INVALID_ARGUMENT;
    }
    if (lqng_status_nonzero) *lqng_status_nonzero = 0;
    if (n_races == 0 || n_steps == 0) return HK_OK;
    ThreadCtx* c = ctx();
    if (!c) return HK_ERR_NO_DEVICE;
    const size_t nb = (size_t)K * n_races;
    const size_t per[7] = {16, 16, 16, 4, 24, 48, 36};
    size_t out_elems = 8;                                              // + the u0 record of the 4-player frame
    for (size_t v : per) out_elems += v;
    const size_t in_bytes = nb * (sizeof(hk_race_kart) + sizeof(hk_race_plan)) + nb * K * sizeof(hk_race_belief);
    static const bool split = !(getenv("HK_RACEN_SPLIT") && atoi(getenv("HK_RACEN_SPLIT")) == 0);   // measurement knob: every game in the 4-player frame
    static const bool multi = !(getenv("HK_RACEN_MULTISTEP") && atoi(getenv("HK_RACEN_MULTISTEP")) == 0);   // measurement knob: one launch per step
    char* d = (char*)dscratch(c, 8, in_bytes + nb * (8 * sizeof(int) + (out_elems + 52) * sizeof(double)) + 512);
    if (!d) return HK_ERR_OUT_OF_MEMORY;
    hk_race_kart* dk = on_device ? karts : (hk_race_kart*)d;
    hk_race_plan* dp = on_device ? plans : (hk_race_plan*)((hk_race_kart*)d + nb);
    hk_race_belief* db = on_device ? beliefs : (hk_race_belief*)((hk_race_plan*)((hk_race_kart*)d + nb) + nb);
    double* o = (double*)(((uintptr_t)((hk_race_belief*)((hk_race_plan*)((hk_race_kart*)d + nb) + nb) + nb * K) + 15) & ~(uintptr_t)15);
    double* dout[7];
    for (int i = 0; i < 7; ++i) { dout[i] = o; o += nb * per[i]; }
    double* du = on_device ? u_hold : o; o += nb * 8;
    unsigned long long* dcount = (unsigned long long*)o;
    double* drec2 = (double*)(dcount + 2); double* du2 = drec2 + nb * 44;                // 2-kart records and answers of the small games
    double* dcs2 = du2 + nb * 4;
    int* dn = (int*)(dcs2 + nb * 4); int* dpl = dn + nb; int* dst = dpl + nb * 4; int* dn2 = dst + nb; int* dst2 = dn2 + nb;
    cudaStream_t s = (on_device && user_stream) ? (cudaStream_t)user_stream : c->stream;
    if (!on_device) {
        HK_CUDA_DRAIN(cudaMemcpyAsync(dk, karts, nb * sizeof(hk_race_kart), cudaMemcpyHostToDevice, s));
        HK_CUDA_DRAIN(cudaMemcpyAsync(dp, plans, nb * sizeof(hk_race_plan), cudaMemcpyHostToDevice, s));
        HK_CUDA_DRAIN(cudaMemcpyAsync(db, beliefs, nb * K * sizeof(hk_race_belief), cudaMemcpyHostToDevice, s));
        HK_CUDA_DRAIN(cudaMemsetAsync(du, 0, nb * 8 * sizeof(double), s));
        HK_CUDA_DRAIN(cudaMemcpy2DAsync(du, 8 * sizeof(double), u_hold, 2 * sizeof(double), 2 * sizeof(double), nb, cudaMemcpyHostToDevice, s));
    }
    HK_CUDA_DRAIN(cudaMemsetAsync(dcount, 0, 2 * sizeof(unsigned long long), s));
    const unsigned blocks = (unsigned)((nb + 127) / 128);
    int searches = 0;
    for (int step = first_step; step < first_step + n_steps; ++step) {
        const bool replan = step > 0 && step % p->planEvery == 0;
        if (replan && !p->highModeMcts) {
            count_launch();
            race_plan_fixed_kernel<<<blocks, 128, 0, s>>>(t->dev, *p, (int)nb, dk, dp);
        } else if (pl) {
            const hk_game_params& gp = game_params_of(pl->game);
            const hk_race_mcts_params& mp = pl->mp;
            const bool begin = step == 0 && mp.first_iterations > 0;
            if ((replan && step < gp.maxEpisodeSteps) || begin) {
                const int budget = begin ? mp.first_iterations : mp.iterations;
                const uint64_t seed = mp.seed + (uint64_t)(step / p->planEvery) * (uint64_t)nb;
                count_launch();
                raceN_mcts_root_kernel<<<blocks, 128, 0, s>>>(t->dev, *p, K, gp.sectionWindow, gp.timePrecision, (int)nb, dk, dp, pl->roots, pl->nearby,
                                                             pl->root_valid, pl->cycles, mp.mode == 0 ? mp.reuse_cycles : 0, pl->fresh, mp.mode == 1);
                HK_CUDA_DRAIN(cudaGetLastError());
                if (mp.mode == 0) rc = mcts_seq_search_device(pl->forest, pl->roots, pl->fresh, budget, seed, pl->best, pl->n_best, nullptr, pl->status, s, false,
                                                              pl->karts_per_race * gp.treeSearchDepth);
                else rc = mcts_search_device(pl->game, pl->roots, (int)nb, budget, mp.rollouts_per_leaf, seed, pl->best, pl->n_best, nullptr, nullptr, nullptr, pl->status, c, s);
                if (rc) { drain(c); return rc; }
                pl->pending_step = step + mp.apply_delay;
                ++searches;
            }
            if (pl->pending_step == step) {
                count_launch();
                raceN_mcts_apply_kernel<<<blocks, 128, 0, s>>>(t->dev, K, (int)nb, dk, pl->nearby, pl->best, pl->n_best, dp, db, pl->fresh, pl->root_valid, pl->cycles);
                HK_CUDA_DRAIN(cudaGetLastError());
                pl->pending_step = -1;
            }
        }
        const bool solve = step % lqr_every == 0;                        // 50 Hz with 2 agents, every 4th step with more (:317)
        if (solve) {
            count_launch();
            raceN_recipe_kernel<<<(unsigned)((nb * 4 + 127) / 128), 128, 0, s>>>(t->dev, *p, K, (int)nb, dk, dp, db, dn, dpl, dout[0], dout[1], dout[2], dout[3], dout[4], dout[5], dout[6],
                                                                                 split ? drec2 : nullptr, split ? dcs2 : nullptr, split ? dn2 : nullptr);
            HK_CUDA_DRAIN(cudaGetLastError());
            if (split) {                                                   // the small games: records and (cos h, sin h) pairs come from the recipe kernel
                rc = lqng_assemble_launch_packed((int)nb, 2, p->horizon, p->dt, drec2, du2, dst2, s, 10, dn2, dcs2);
                if (rc) { drain(c); return rc; }
            }
            rc = lqng_assemble_launch((int)nb, 4, p->horizon, p->dt, dout[0], dout[1], dout[2], dout[3], dout[4], dout[5], dout[6], du, dst, s, 9, dn, split ? 3 : 0);
            if (rc) { drain(c); return rc; }
            if (split) {
                count_launch();
                raceN_merge2_kernel<<<blocks, 128, 0, s>>>((int)nb, dn, du2, dst2, du, dst);
                HK_CUDA_DRAIN(cudaGetLastError());
            }
        }
        // the steps up to the next solve / planning event ride in the same launch (held controls, no kernel in between)
        int n_sub = 1;
        while (multi && step + n_sub < first_step + n_steps && (step + n_sub) % lqr_every != 0 && (step + n_sub) % p->planEvery != 0 &&
               !(pl && pl->pending_step == step + n_sub))
            ++n_sub;
        count_launch();
        race_steps_kernel<<<blocks, 128, 0, s>>>(t->dev, *p, (int)nb, step, n_sub, du, 8, solve ? dst : nullptr, dcount, dk, dp, pl ? pl->root_valid : nullptr,
                                                 pl ? pl->cycles : nullptr);
        HK_CUDA_DRAIN(cudaGetLastError());
        step += n_sub - 1;
    }
    if (!on_device) {
        HK_CUDA_DRAIN(cudaMemcpyAsync(karts, dk, nb * sizeof(hk_race_kart), cudaMemcpyDeviceToHost, s));
        HK_CUDA_DRAIN(cudaMemcpyAsync(plans, dp, nb * sizeof(hk_race_plan), cudaMemcpyDeviceToHost, s));
        HK_CUDA_DRAIN(cudaMemcpyAsync(beliefs, db, nb * K * sizeof(hk_race_belief), cudaMemcpyDeviceToHost, s));
        HK_CUDA_DRAIN(cudaMemcpy2DAsync(u_hold, 2 * sizeof(double), du, 8 * sizeof(double), 2 * sizeof(double), nb, cudaMemcpyDeviceToHost, s));
    }
    unsigned long long count = 0;
    HK_CUDA_DRAIN(cudaMemcpyAsync(&count, dcount, sizeof(count), cudaMemcpyDeviceToHost, s));
    std::vector<int> mst;
    if (pl && searches) { mst.resize(nb); HK_CUDA_DRAIN(cudaMemcpyAsync(mst.data(), pl->status, 4 * nb, cudaMemcpyDeviceToHost, s)); }
    HK_CUDA_DRAIN(cudaStreamSynchronize(s));
    if (lqng_status_nonzero) *lqng_status_nonzero = (int64_t)count;
    for (size_t a = 0; a < mst.size(); ++a)
        if (mst[a] == 1) { set_error("hk_raceN_run: upNext() == -1 reached in the tree of agent %zu (KartDiscreteGame.cs:326 would throw)", a); return HK_ERR_NO_UPNEXT; }
    return HK_OK;
}

extern "C" int hk_raceN_run(const hk_track* t, const hk_race_params* p, hk_race_planner* pl, int K, int lqr_every, int n_races, int first_step,
                            int n_steps, hk_race_kart* karts, hk_race_plan* plans, hk_race_belief* beliefs, double* u_hold,
                            int64_t* lqng_status_nonzero)
{
    return raceN_run_impl(t, p, pl, K, lqr_every, n_races, first_step, n_steps, karts, plans, beliefs, u_hold, lqng_status_nonzero, false, nullptr);
}

extern "C" int hk_raceN_run_device(const hk_track* t, const hk_race_params* p, hk_race_planner* pl, int K, int lqr_every, int n_races, int first_step,
                                   int n_steps, hk_race_kart* d_karts, hk_race_plan* d_plans, hk_race_belief* d_beliefs, double* d_u_hold,
                                   int64_t* lqng_status_nonzero, void* cuda_stream)
{
    return raceN_run_impl(t, p, pl, K, lqr_every, n_races, first_step, n_steps, d_karts, d_plans, d_beliefs, d_u_hold, lqng_status_nonzero, true, cuda_stream);
}
