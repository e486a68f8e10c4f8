/*
 * hk_oracle_race.c — CPU ORACLE (test infrastructure, NOT the product) for the closed loop without PhysX
 * (SURVEY.md §8f ranks 1-2, §3.3).  Plain C restatement of
 *   - the LQNG problem recipe of HierarchicalKartAgent.SolveLQR, raycast-free branches
 *     (Assets/Karting/Scripts/AI/HierarchicalKartAgent.cs:699-1197; initial states :730-736, targets :745-817, target-heading
 *     rules :821-823, :877-890, :896-898, :919-923, own weights :930-962, avoid weights :999-1023, opponent target
 *     weights :1089-1091, control weight :1192-1196),
 *   - planFixed (:145-166),
 *   - the actuator map (:1206-1224) on the kinematic model the planners assume (MPC/KartMPCDynamics.cs:55-70),
 *   - the checkpoint bookkeeping of OnTriggerEnter (:611-662) with DiscretePositionTracker.CalculateLane (:116-148).
 * Unity's Mathf.X(float...) is (float)Math.X(double...): float32-typed expressions are evaluated in double on the
 * float32 inputs and rounded once.  PARITY UNPINNED by the reference (no tests, no .NET here): pinned instead by the
 * independently written numpy recipe of hierarchicalkarting_b200/scenarios.py (tests/test_oracle_cpu.py).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "hk_oracle.h"

static const float PI_F = 3.14159274f;                     /* Mathf.PI */

static double angle_difference(double a1, double a2)       /* HierarchicalKartAgent.cs:1341-1344 */
{
    return atan2(sin(a2 - a1), cos(a2 - a1));
}
static float mathf_atan2(float y, float x) { return (float)atan2((double)y, (double)x); }
static float magnitude2(float dx, float dz) { return (float)sqrt((double)(dx * dx + dz * dz)); }   /* Vector3.magnitude, dy = 0 */
static float wrap2pi_f(float a) { return a < 0 ? a + 2 * PI_F : a; }

typedef struct {
    int n;
    const hk_section* sec;
    const double *trig, *fwd, *lane;
} track_view;

static int is_straight(const track_view* t, int section) { return t->sec[section % t->n].insideR == 0.0f; }

/* target point / velocity of kart k for checkpoint index idx, from a lane plan (0 = key absent => Trigger, max speed) */
static void plan_target(const track_view* t, const hk_race_params* p, const int8_t* lanes, const float* vels, int idx,
                        double* xz, double* vel)
{
    const double max_speed = (double)p->topSpeed;          /* GetMaxSpeed(), ArcadeKart.cs:210 (TopSpeed > ReverseSpeed) */
    if (lanes[idx] != 0) {
        xz[0] = t->lane[(idx * 4 + lanes[idx] - 1) * 2];
        xz[1] = t->lane[(idx * 4 + lanes[idx] - 1) * 2 + 1];
        const double v = (double)vels[idx] + (p->highModeMcts ? p->velocityBucketSize * 2 : 0);
        *vel = max_speed < v ? max_speed : v;
    } else {
        xz[0] = t->trig[idx * 2];
        xz[1] = t->trig[idx * 2 + 1];
        *vel = max_speed;
    }
}

/* One problem: ego = karts[e], other = karts[1 - e].  Outputs: x0[2][4], target[2][4], tw[2][4], cw[2], aw[2][1][2],
 * otgt[2][1][4], otw[2][1][3]. */
void hk_oracle_race_recipe_one(const hk_section* sections, const double* trig, const double* fwd, const double* lane, int n_sections,
                               const hk_race_params* p, const hk_race_kart* karts, const hk_race_plan* plans, int e,
                               double* x0, double* target, double* tw, double* cw, double* aw, double* otgt, double* otw)
{
    const track_view t = {n_sections, sections, trig, fwd, lane};
    double tl[2][2], vel[2];
    for (int i = 0; i < 2; ++i) {
        const hk_race_kart* k = &karts[i == 0 ? e : 1 - e];
        const int8_t* lanes = i == 0 ? plans[e].lane : plans[e].oppLane;      /* own plan / belief about the other (:745-817) */
        const float* vels = i == 0 ? plans[e].vel : plans[e].oppVel;
        x0[i * 4 + 0] = k->x; x0[i * 4 + 1] = k->z; x0[i * 4 + 2] = k->v; x0[i * 4 + 3] = k->h;     /* :730-736 */
        const int s = k->section + 1;                                           /* :745 */
        const int idx = s % t.n, idx2 = (s + 1) % t.n;
        double nl[2], nvel;
        plan_target(&t, p, lanes, vels, idx, tl[i], &vel[i]);
        plan_target(&t, p, lanes, vels, idx2, nl, &nvel);
        const int stopped = (float)k->v <= 5.0f;                                /* :808 */
        double tx = tl[i][0], tz = tl[i][1], tv = stopped ? 0.0 : vel[i];
        const float d_t = magnitude2((float)(tl[i][0] - k->x), (float)(tl[i][1] - k->z));
        const int near = d_t <= (is_straight(&t, k->section) ? 10.5f : 7.5f);   /* :823 */
        const float d_c = magnitude2((float)(t.trig[idx * 2] - k->x), (float)(t.trig[idx * 2 + 1] - k->z));
        const int follow = near && (d_c <= 4.0f);                               /* :877-890, centre-line distance stand-in */
        const double f1w = (double)wrap2pi_f(mathf_atan2((float)(tl[i][1] - k->z), (float)(tl[i][0] - k->x)));
        const double f2w = (double)wrap2pi_f(mathf_atan2((float)(nl[1] - tl[i][1]), (float)(nl[0] - tl[i][0])));
        const double f6w = (double)wrap2pi_f(mathf_atan2((float)(nl[1] - k->z), (float)(nl[0] - k->x)));
        const double h0 = k->h;
        double blend = f1w - angle_difference(f2w, f1w) * (double)0.4f;         /* :896 */
        if (blend < 0) blend += 2 * (double)PI_F;
        double th;
        if (follow) th = h0 - angle_difference(h0, f6w);                        /* :887 */
        else if (near) th = h0 - angle_difference(h0, blend);                   /* :898 */
        else th = h0 - angle_difference(h0, f1w);                               /* :921 */
        if (follow) { tx = nl[0]; tz = nl[1]; if (!stopped) tv = nvel; }
        target[i * 4 + 0] = tx; target[i * 4 + 1] = tz; target[i * 4 + 2] = tv; target[i * 4 + 3] = th;
        /* own target weights, 2-agent branch (:930-962) */
        const double vmax1 = k->v > 1.0 ? k->v : 1.0;
        const double w_xz = stopped ? 0.3 * 3.1 : 0.3 * 3.1 / vmax1;
        tw[i * 4 + 0] = w_xz; tw[i * 4 + 1] = w_xz;
        tw[i * 4 + 2] = stopped ? -2.0 : 5e-4;
        tw[i * 4 + 3] = p->highModeMcts ? 3.5 : 1.9;
        cw[i] = 0.115;                                                          /* :1192-1196 */
    }
    for (int i = 0; i < 2; ++i) {                                               /* avoid + opponent-target weights (:964-1190) */
        const int o = 1 - i;
        const hk_race_kart* ki = &karts[i == 0 ? e : 1 - e];
        const hk_race_kart* ko = &karts[o == 0 ? e : 1 - e];
        float mult;                                                              /* :996-1003 (2-agent environment) */
        if (i == 0) mult = p->highModeMcts ? 1.0f : 0.45f;                       /* k == this: HighMode == Fixed ? 0.45f : 1.0f */
        else mult = 1.3f;                                                        /* Fixed ? 1.3f : 1.3f */
        const float dist = magnitude2((float)(ko->x - ki->x), (float)(ko->z - ki->z));
        const int far = dist > 8 || !ko->active;                                 /* || !o.is_active, :1010 */
        const float w32 = 1.0f / ((float)pow((double)dist, (double)1.5f) * mult);   /* 1f/(Mathf.Pow(d,1.5f)*mult), :1019 */
        const double w = far ? 0.0 : (double)w32;
        aw[i * 2 + 0] = w; aw[i * 2 + 1] = w;
        otgt[i * 4 + 0] = tl[o][0]; otgt[i * 4 + 1] = tl[o][1]; otgt[i * 4 + 2] = vel[o]; otgt[i * 4 + 3] = 0.0;
        const double vmax1 = ki->v > 1.0 ? ki->v : 1.0;
        const double wxz = (p->highModeMcts ? 0.2 : 0.1) / vmax1;               /* :1089-1091 */
        otw[i * 3 + 0] = far ? 0.0 : wxz; otw[i * 3 + 1] = far ? 0.0 : wxz; otw[i * 3 + 2] = far ? 0.0 : 0.08;
    }
}

void hk_oracle_race_recipe(const hk_section* sections, const double* trig, const double* fwd, const double* lane, int n_sections,
                           const hk_race_params* p, int n_races, const hk_race_kart* karts, const hk_race_plan* plans,
                           double* x0, double* target, double* tw, double* cw, double* aw, double* otgt, double* otw)
{
    for (int r = 0; r < n_races; ++r)
        for (int e = 0; e < 2; ++e) {
            const size_t b = (size_t)2 * r + e;
            hk_oracle_race_recipe_one(sections, trig, fwd, lane, n_sections, p, karts + 2 * r, plans + 2 * r, e, x0 + b * 8,
                                      target + b * 8, tw + b * 8, cw + b * 2, aw + b * 4, otgt + b * 8, otw + b * 6);
        }
}

/* SolveLQR's problem for ANY number of agents (HierarchicalKartAgent.cs:699-1201), written from the C# text like the 2-agent function above;
 * the Python restatement oracle/np_recipe.py is compared with it in tests/.  One race: karts[K], plans[K] (each kart's own plan), beliefs[K]
 * = the EGO's beliefs about the other karts (opponentUpcomingLanes / Velocities), e = the ego.  Outputs for the n = *n_players real players
 * in joint order: players[n] (race-local kart), x0[n][4], target[n][4], tw[n][4], cw[n], and per player its n - 1 private slots (its
 * otherAgents first, then its teamAgents, each in environment order): aw[n][3][2], otgt[n][3][4], otw[n][3][3] (unused slots zero). */
void hk_oracle_raceN_recipe_one(const hk_section* sections, const double* trig, const double* fwd, const double* lane, int n_sections,
                                const hk_race_params* p, int K, const hk_race_kart* karts, const hk_race_plan* plans,
                                const hk_race_belief* beliefs, int e, int* n_players, int* players, double* x0, double* target, double* tw,
                                double* cw, double* aw, double* otgt, double* otw)
{
    const track_view t = {n_sections, sections, trig, fwd, lane};
    const int fixed = !p->highModeMcts;
    const hk_race_kart* me = &karts[e];
    /* allPlayers = { this } + teamAgents + otherAgents (:702) */
    int all[HK_MAX_KARTS], n_all = 0;
    all[n_all++] = e;
    for (int a = 0; a < K; ++a) if (a != e && karts[a].team == me->team) all[n_all++] = a;
    for (int a = 0; a < K; ++a) if (karts[a].team != me->team) all[n_all++] = a;
    /* with more than two agents in the environment only those within 8 m of the ego take part (:709-721) */
    int act[HK_MAX_KARTS], n = 0, nearbyAgents = -1;
    if (K > 2) {
        for (int i = 0; i < n_all; ++i) {
            const hk_race_kart* k = &karts[all[i]];
            if (magnitude2((float)(k->x - me->x), (float)(k->z - me->z)) < 8) { nearbyAgents += 1; act[n++] = all[i]; }
        }
    } else {
        for (int i = 0; i < n_all; ++i) act[n++] = all[i];
    }
    if (nearbyAgents < 1) nearbyAgents = 1;                                       /* Math.Max(nearbyAgents, 1) :726 */
    *n_players = n;
    memset(aw, 0, sizeof(double) * (size_t)n * 6);
    memset(otgt, 0, sizeof(double) * (size_t)n * 12);
    memset(otw, 0, sizeof(double) * (size_t)n * 9);
    for (int i = 0; i < n; ++i) {
        const int ki = act[i];
        const hk_race_kart* k = &karts[ki];
        players[i] = ki;
        /* the plan the ego holds for kart ki: its own m_UpcomingLanes, or its belief about ki */
        const int8_t* lanes = ki == e ? plans[e].lane : beliefs[ki].lane;
        const float* vels = ki == e ? plans[e].vel : beliefs[ki].vel;
        x0[i * 4 + 0] = k->x; x0[i * 4 + 1] = k->z; x0[i * 4 + 2] = k->v; x0[i * 4 + 3] = k->h;       /* :730-736 */
        const int s = k->section + 1;                                             /* :745 */
        const int idx = s % t.n, idx2 = (s + 1) % t.n;
        double tl[2], nl[2], vel, nvel;
        plan_target(&t, p, lanes, vels, idx, tl, &vel);
        plan_target(&t, p, lanes, vels, idx2, nl, &nvel);
        const int stopped = (float)k->v <= 5.0f;                                  /* :808 */
        double tx = tl[0], tz = tl[1], tv = stopped ? 0.0 : vel;
        const float d_t = magnitude2((float)(tl[0] - k->x), (float)(tl[1] - k->z));
        const int near = d_t <= (is_straight(&t, k->section) ? 10.5f : 7.5f);     /* :823 */
        const float d_c = magnitude2((float)(t.trig[idx * 2] - k->x), (float)(t.trig[idx * 2 + 1] - k->z));
        const int follow = near && (d_c <= 4.0f);                                 /* :877-890, centre-line distance stand-in */
        const double h0 = k->h;
        double th;
        if (follow) {
            const double f6w = (double)wrap2pi_f(mathf_atan2((float)(nl[1] - k->z), (float)(nl[0] - k->x)));
            th = h0 - angle_difference(h0, f6w);                                  /* :887 */
            tx = nl[0]; tz = nl[1];
            if (!stopped) tv = nvel;
        } else {
            const double f1w = (double)wrap2pi_f(mathf_atan2((float)(tl[1] - k->z), (float)(tl[0] - k->x)));
            if (near) {
                const double f2w = (double)wrap2pi_f(mathf_atan2((float)(nl[1] - tl[1]), (float)(nl[0] - tl[0])));
                double blend = f1w - angle_difference(f2w, f1w) * (double)0.4f;   /* :896 */
                if (blend < 0) blend += 2 * (double)PI_F;
                th = h0 - angle_difference(h0, blend);                            /* :898 */
            } else th = h0 - angle_difference(h0, f1w);                           /* :921 */
        }
        target[i * 4 + 0] = tx; target[i * 4 + 1] = tz; target[i * 4 + 2] = tv; target[i * 4 + 3] = th;
        /* own target weights (:928-962) */
        const double vmax1 = k->v > 1.0 ? k->v : 1.0;
        tw[i * 4 + 3] = n > 2 ? (fixed ? 2.5 : 3.5) * nearbyAgents : (fixed ? 1.9 : 3.5);
        if (stopped) { tw[i * 4 + 0] = tw[i * 4 + 1] = nearbyAgents * 0.3 * 3.1; tw[i * 4 + 2] = nearbyAgents * -2; }
        else { tw[i * 4 + 0] = tw[i * 4 + 1] = nearbyAgents * 0.3 * 3.1 / vmax1; tw[i * 4 + 2] = nearbyAgents * 5e-4; }
        cw[i] = n > 2 ? (fixed ? 0.135 : 0.25) : 0.115;                           /* :1192-1196 */
        /* avoid-weight multiplier (:977-1003) */
        float multiplier;
        if (K > 2 && n > 2) multiplier = ki == e ? (fixed ? 0.55f : 1.0f) / nearbyAgents : 1.7f / nearbyAgents;
        else multiplier = ki == e ? (fixed ? 0.45f : 1.0f) : 1.3f;
        int slot = 0, nearbyOpponents = 0;
        for (int pass = 0; pass < 2; ++pass) {                                    /* k.otherAgents (:1004-1096), then k.teamAgents (:1098-1190) */
            for (int o = 0; o < K; ++o) {
                if (o == ki) continue;
                const int mate = karts[o].team == k->team;
                if (pass == 0 ? mate : !mate) continue;
                int in_game = 0;
                for (int j = 0; j < n; ++j) in_game |= act[j] == o;
                if (!in_game) continue;                                           /* !actualAllPlayers.Contains(o) */
                const hk_race_kart* ko = &karts[o];
                const float dist = magnitude2((float)(ko->x - k->x), (float)(ko->z - k->z));
                const int off = dist > 8 || !ko->active;
                const float m = pass == 0 ? multiplier : multiplier / 2.0f;       /* multiplier2 :1113 */
                const double w = off ? 0.0 : (double)(1.0f / ((float)pow((double)dist, (double)1.5f) * m));    /* :1019, :1114 */
                aw[(i * 3 + slot) * 2 + 0] = w; aw[(i * 3 + slot) * 2 + 1] = w;
                if (pass == 0 && !off) nearbyOpponents += 1;
                /* the other kart's target: the ego's plan / belief for it at ITS next checkpoint (:1036-1066); a teammate's target speed is
                 * getMaxSpeedForState() (:1138-1160), the stand-in's top speed */
                const int8_t* ol = o == e ? plans[e].lane : beliefs[o].lane;
                const float* ov = o == e ? plans[e].vel : beliefs[o].vel;
                double oxz[2], ovel;
                plan_target(&t, p, ol, ov, (ko->section + 1) % t.n, oxz, &ovel);
                otgt[(i * 3 + slot) * 4 + 0] = oxz[0]; otgt[(i * 3 + slot) * 4 + 1] = oxz[1];
                otgt[(i * 3 + slot) * 4 + 2] = pass == 0 ? ovel : (double)p->topSpeed; otgt[(i * 3 + slot) * 4 + 3] = 0.0;
                double wxz = 0.0, wv = 0.0;
                if (pass == 0) {
                    if (!off) {
                        if (n > 2) { wxz = (fixed ? 0.1 : 0.2) / (vmax1 * nearbyAgents); wv = 0.08 / nearbyAgents; }   /* :1083-1085 */
                        else { wxz = (fixed ? 0.1 : 0.2) / vmax1; wv = 0.08; }                                          /* :1089-1091 */
                    }
                } else if (!off && nearbyOpponents >= 1) {
                    if (n > 2) wxz = -(fixed ? 0.0 : 3e-5) / (vmax1 * nearbyAgents);                                    /* :1178-1180 */
                    else wxz = -(fixed ? 1e-4 : 2e-4) / vmax1;                                                          /* :1184-1186 */
                }
                otw[(i * 3 + slot) * 3 + 0] = wxz; otw[(i * 3 + slot) * 3 + 1] = wxz; otw[(i * 3 + slot) * 3 + 2] = wv;
                ++slot;
            }
        }
    }
}

/* planFixed (:145-166) */
void hk_oracle_race_plan_fixed(const hk_section* sections, int n_sections, const hk_race_params* p, int n_karts,
                               const hk_race_kart* karts, hk_race_plan* plans)
{
    for (int k = 0; k < n_karts; ++k) {
        if (!karts[k].active) continue;
        const int s = karts[k].section;
        const int hi = s + p->treeSearchDepth < 1000 ? s + p->treeSearchDepth : 1000;
        for (int i = s + 1; i < hi + 1; ++i) {
            const int key = i % n_sections;
            if (plans[k].lane[key] == 0) {
                plans[k].lane[key] = (int8_t)sections[(i - 1) % n_sections].optimalLane;      /* getOptimalNextLane */
                plans[k].vel[key] = p->topSpeed;                                               /* GetMaxSpeed() */
            }
        }
    }
}

/* actuator map (:1206-1224) + kinematic plant (KartMPCDynamics.cs:55-70) + OnTriggerEnter bookkeeping (:611-662) */
void hk_oracle_race_step(const hk_section* sections, const double* trig, const double* fwd, const double* lane, int n_sections,
                         const hk_race_params* p, int n_karts, int episode_step, const double* u, hk_race_kart* karts,
                         hk_race_plan* plans)
{
    const track_view t = {n_sections, sections, trig, fwd, lane};
    const double TWO_PI = 6.283185307179586;
    for (int i = 0; i < n_karts; ++i) {
        hk_race_kart* k = &karts[i];
        if (!k->active) continue;
        const float max_ang = k->steer * 0.4f;                                  /* getMaxAngularVelocity, ArcadeKart.cs:505-510 */
        float ang = (float)u[i * 2 + 1];
        ang = ang < -max_ang ? -max_ang : (ang > max_ang ? max_ang : ang);      /* Mathf.Clamp :1206 */
        int accel = 0, brake = 0;
        if (u[i * 2] < 0) brake = 1;
        else if (u[i * 2] > 0) accel = 1;
        else ang = 0.0f;
        const float steering = ang / (0.4f * k->steer);                         /* m_Steering :1224 */
        const float turning_power = steering * k->steer * (fabsf((float)k->v) > 0.5f ? 1.0f : 0.0f);   /* ArcadeKart.cs:406 */
        const double omega = (double)(turning_power * 0.4f);
        const double x_old = k->x, z_old = k->z;
        k->x = x_old + p->dt * k->v * cos(k->h);
        k->z = z_old + p->dt * k->v * sin(k->h);
        /* Unity's yaw is left-handed: a positive TurnInput turns the kart clockwise seen from above, i.e. the solver's heading
         * atan2(forward.z, forward.x) DEcreases.  SolveLQR compensates by mirroring the target heading about the current one
         * (h0 - AngleDifference(h0, target) = 2 h0 - target, :887,:898,:921), so the stand-in must turn like Unity does. */
        double h = k->h - p->dt * omega;
        if (h < 0) h += TWO_PI;
        if (h >= TWO_PI) h -= TWO_PI;
        double v = k->v;
        if (accel) { v += p->dt * (double)p->accel; if (v > (double)p->topSpeed) v = (double)p->topSpeed; }
        else if (brake) { v -= p->dt * (double)p->braking; if (v < 0) v = 0; }
        else { v -= p->dt * (double)p->coastingDrag; if (v < 0) v = 0; }       /* MoveTowards(v, 0, dt CoastingDrag), ArcadeKart.cs:431 */
        k->h = h; k->v = v;
        /* did the kart enter the trigger of checkpoint section+1 during this step? */
        const int index = k->section + 1, c = index % t.n;
        const double fx = t.fwd[c * 2], fz = t.fwd[c * 2 + 1];
        const double s_old = (x_old - t.trig[c * 2]) * fx + (z_old - t.trig[c * 2 + 1]) * fz;
        const double s_new = (k->x - t.trig[c * 2]) * fx + (k->z - t.trig[c * 2 + 1]) * fz;
        const double lat = -(k->x - t.trig[c * 2]) * fz + (k->z - t.trig[c * 2 + 1]) * fx;
        if (s_old < 0 && s_new >= 0 && fabs(lat) <= (double)p->gateHalfWidth) {
            int lane_new = 1;                                                   /* CalculateLane: first minimum */
            float best = 0;
            for (int l = 0; l < 4; ++l) {
                const float d = magnitude2((float)(k->x - t.lane[(c * 4 + l) * 2]), (float)(k->z - t.lane[(c * 4 + l) * 2 + 1]));
                if (l == 0 || d < best) { best = d; lane_new = l + 1; }
            }
            hk_race_plan* pl = &plans[i];
            if (pl->lane[c] != 0) {                                             /* KartAgent.cs:226-239, InitCheckpointIndex = 0 */
                const float d = magnitude2((float)(k->x - t.lane[(c * 4 + pl->lane[c] - 1) * 2]),
                                           (float)(k->z - t.lane[(c * 4 + pl->lane[c] - 1) * 2 + 1])) - 1.3f;
                pl->avgLaneDiff = ((d > 0.0f ? d : 0.0f) + pl->avgLaneDiff * (float)(index - 1)) / (float)index;
                pl->avgVelDiff = (((float)k->v - pl->vel[c]) + pl->avgVelDiff * (float)(index - 1)) / (float)index;
            }
            pl->lane[c] = 0;                                                    /* m_UpcomingLanes.Remove (:631-632) */
            pl->vel[c] = 0.0f;
            pl->sectionTimes[c] = episode_step;                                 /* :650 */
            if (c == 0 && index / t.n >= 1 && index / t.n <= HK_MAX_LAPS) pl->lapStep[index / t.n - 1] = episode_step;
            const int dl = abs(k->lane - lane_new);
            if (k->laneChanges + dl > p->maxLaneChanges && is_straight(&t, k->section)) k->illegalLaneChanges += 1;   /* :638-642 */
            if (is_straight(&t, k->section) != is_straight(&t, index)) k->laneChanges = 0;                           /* :643-646 */
            else if (k->lane != lane_new) k->laneChanges += dl;                                                        /* :647-650 */
            k->section = index;
            k->lane = lane_new;
            k->sectionStep = episode_step;
            if (k->section == p->goalSection) k->active = 0;                    /* ReachGoalSection (:652-655) */
        }
    }
}

/* the loop of hk_race_run; returns the number of solves with non-zero status */
long long hk_oracle_race_run(const hk_section* sections, const double* trig, const double* fwd, const double* lane, int n_sections,
                             const hk_race_params* p, int n_races, int first_step, int n_steps, hk_race_kart* karts,
                             hk_race_plan* plans, double* u_last)
{
    long long bad = 0;
    for (int step = first_step; step < first_step + n_steps; ++step) {
        if (step > 0 && step % p->planEvery == 0 && !p->highModeMcts)       /* HierarchicalKartAgent.cs:331-353: planFixed only in Fixed mode */
            hk_oracle_race_plan_fixed(sections, n_sections, p, 2 * n_races, karts, plans);
#pragma omp parallel for reduction(+ : bad) schedule(static)
        for (int r = 0; r < n_races; ++r) {
            double u[4];
            for (int e = 0; e < 2; ++e) {
                double x0[8], target[8], tw[8], cw[2], aw[4], otgt[8], otw[6];
                hk_oracle_race_recipe_one(sections, trig, fwd, lane, n_sections, p, karts + 2 * r, plans + 2 * r, e, x0, target, tw,
                                          cw, aw, otgt, otw);
                double A[2 * 16], B[2 * 8], Q[2 * 64], q[2 * 8], R[2 * 4], u0[4];
                for (int i = 0; i < 2; ++i) {
                    hk_oracle_bicycle_A(p->dt, x0 + 4 * i, A + 16 * i);
                    hk_oracle_bicycle_B(p->dt, B + 8 * i);
                    hk_oracle_cost(1, target + 4 * i, tw + 4 * i, cw[i], aw + 2 * i, otgt + 4 * i, otw + 3 * i, Q + 64 * i, q + 8 * i,
                                   R + 4 * i);
                }
                bad += hk_oracle_lqng_solve(2, p->horizon, 0, A, B, Q, q, R, x0, u0, NULL, NULL, NULL) != 0;
                u[2 * e] = u0[0]; u[2 * e + 1] = u0[1];
            }
            if (u_last) memcpy(u_last + 4 * (size_t)r, u, sizeof(u));
            hk_oracle_race_step(sections, trig, fwd, lane, n_sections, p, 2, step, u, karts + 2 * r, plans + 2 * r);
        }
    }
    return bad;
}

/* ---- MCTS high level of the loop: planWithMCTS's root state, the waypoint hand-off, and the planner's schedule --------------------------
 * Restated from HierarchicalKartAgent.cs (not from the CUDA library or its host mirrors):
 *   root state            :180-245  (nearby agents within sectionWindow sections, all placed at the furthest one's section; velocity bucket
 *                                    (0, bucket) because the search loop breaks at i = 0, quirk B.6-1; player = 0, quirk B.6-2)
 *   schedule              :85-93 (plan at episode start, T = 1.5), :331-353 (every planEvery steps), :175 / :265 (new tree / continue while
 *                          CyclesRootProcessed < 3 / nothing), :250-253, :271-273 (what the background thread sets when it finishes),
 *                          :660-661 (a checkpoint crossing drops the tree)
 *   hand-off              :366-402
 * 2-kart races; each kart is its own team (getTeamID of the head-to-head scenes). */
int hk_oracle_race_mcts_root(const hk_race_params* p, const hk_game_params* gp, int n_sections, const hk_race_kart* karts,
                             const hk_race_plan* plans, int n_agents_in_race, int ego, hk_game_state* st, int* nearby)
{
    const int me_sec = karts[ego].section;
    int initial = me_sec, furthest = ego, n = 0;
    for (int a = 0; a < n_agents_in_race; ++a) {                                        /* :182-193 */
        if (abs(karts[a].section - me_sec) < gp->sectionWindow) {
            nearby[n++] = a;
            if (karts[a].section > initial) initial = karts[a].section;
            if (initial == karts[a].section) furthest = a;
        }
    }
    for (int i = n; i < HK_MAX_KARTS; ++i) nearby[i] = -1;
    memset(st, 0, sizeof(*st));
    st->n_karts = n;
    st->initialSection = initial; st->lastCompletedSection = initial; st->finalSection = initial + gp->treeSearchDepth;   /* :194, :235-245 */
    const int max_speed = (int)p->topSpeed;                                             /* (int)GetMaxSpeed() */
    for (int i = 0; i < n; ++i) {
        const hk_race_kart* k = &karts[nearby[i]];
        hk_kart_state* ks = &st->karts[i];
        ks->min_velocity = 0;                                                            /* the loop :200-208 breaks at i = 0: |v| >= 0 */
        ks->max_velocity = gp->velocityBucketSize < max_speed ? gp->velocityBucketSize : max_speed;
        int t_at = 0;
        if (k->section != initial) {                                                     /* :211-214: int * float * int, then (int) */
            const int d = plans[nearby[i]].sectionTimes[k->section % n_sections] - plans[furthest].sectionTimes[k->section % n_sections];
            t_at = (int)(((float)d * 0.02f) * (float)gp->timePrecision);
        }
        ks->player = 0;                                                                  /* count is never incremented (:195, :221) */
        ks->team = nearby[i];
        ks->section = initial;
        ks->timeAtSection = t_at;
        ks->lane = k->lane;
        ks->tireAge = (int)((4.0f - k->steer) / (4.0f - 1.0f) * 10000);                  /* (MaxSteer - Steer) / (MaxSteer - MinSteer) * 10000, :226 */
        ks->laneChanges = k->laneChanges;
        ks->infeasible = 0;
    }
    return n;
}

void hk_oracle_race_apply_best(int n_sections, const hk_race_kart* karts, hk_race_plan* plans, int ego, const int* nearby,
                               const hk_game_state* best, int n_best)
{
    const int sec = karts[ego].section;
    hk_race_plan* pl = &plans[ego];
    for (int b = 0; b < n_best; ++b)                                                     /* foreach gameState in bestStates :366 */
        for (int i = 0; i < best[b].n_karts; ++i) {                                      /* foreach kartState :368 */
            const hk_kart_state* ks = &best[b].karts[i];
            const int key = ks->section % n_sections;
            if (nearby[i] == ego && ks->section > sec + (sec == 0 ? 0 : 1)) {            /* :371 */
                pl->lane[key] = (int8_t)ks->lane;                                        /* :381-382 */
                pl->vel[key] = (float)ks->max_velocity;
            } else if (nearby[i] != ego && nearby[i] >= 0) {                             /* :395-400 */
                pl->oppLane[key] = (int8_t)ks->lane;
                pl->oppVel[key] = (float)ks->max_velocity;
            }
        }
}

typedef struct {
    hk_oracle_tree* tree;       /* currentRoot */
    int root_valid, cycles;     /* currentRoot != null, CyclesRootProcessed */
    int pending;                /* a search result has not landed yet */
    int pending_fresh;
    int nearby[HK_MAX_KARTS];
    hk_game_state best[HK_MCTS_MAX_SEQ];
    int n_best;
} oagent;

struct hk_oracle_planner {
    const hk_oracle_game* g;
    hk_game_params gp;
    hk_race_mcts_params mp;
    int n_agents, pending_step;
    oagent* a;
};

hk_oracle_planner* hk_oracle_planner_create(const hk_oracle_game* g, const hk_game_params* gp, const hk_race_mcts_params* mp, int n_races)
{
    hk_oracle_planner* pl = (hk_oracle_planner*)calloc(1, sizeof(*pl));
    pl->g = g; pl->gp = *gp; pl->mp = *mp; pl->n_agents = 2 * n_races; pl->pending_step = -1;
    pl->a = (oagent*)calloc((size_t)pl->n_agents, sizeof(oagent));
    return pl;
}

void hk_oracle_planner_destroy(hk_oracle_planner* pl)
{
    if (!pl) return;
    for (int i = 0; i < pl->n_agents; ++i) hk_oracle_tree_destroy(pl->a[i].tree);
    free(pl->a); free(pl);
}

void hk_oracle_planner_state(const hk_oracle_planner* pl, int32_t* root_valid, int32_t* cycles)
{
    for (int i = 0; i < pl->n_agents; ++i) { if (root_valid) root_valid[i] = pl->a[i].root_valid; if (cycles) cycles[i] = pl->a[i].cycles; }
}

/* The loop of hk_oracle_race_run with the MCTS high level (sequential search only).  Returns the number of LQNG solves with a zero
 * pivot, or -1 if a search failed. */
long long hk_oracle_race_run_planned(const hk_section* sections, const double* trig, const double* fwd, const double* lane, int n_sections,
                                     const hk_race_params* p, hk_oracle_planner* pl, int n_races, int first_step, int n_steps,
                                     hk_race_kart* karts, hk_race_plan* plans, double* u_last)
{
    long long bad = 0;
    int failed = 0;
    for (int step = first_step; step < first_step + n_steps; ++step) {
        const int replan = step > 0 && step % p->planEvery == 0 && step < pl->gp.maxEpisodeSteps;     /* :331 */
        const int begin = step == 0 && pl->mp.first_iterations > 0;                                  /* :85-93 */
        if (replan || begin) {
            const int budget = begin ? pl->mp.first_iterations : pl->mp.iterations;
            const uint64_t seed = pl->mp.seed + (uint64_t)(step / p->planEvery) * (uint64_t)pl->n_agents;
#pragma omp parallel for schedule(dynamic, 1)
            for (int id = 0; id < pl->n_agents; ++id) {
                oagent* ag = &pl->a[id];
                const int r2 = id & ~1, ego = id & 1;
                ag->pending = 0;
                if (!karts[id].active) continue;                                                      /* inactiveAgents.Contains(this) */
                uint64_t rng = 1;
                int fresh;
                if (!(pl->mp.reuse_cycles > 0 && ag->root_valid)) {                                   /* currentRoot == null (:175) */
                    hk_game_state root;
                    hk_oracle_race_mcts_root(p, &pl->gp, n_sections, karts + r2, plans + r2, 2, ego, &root, ag->nearby);
                    hk_oracle_tree_destroy(ag->tree);
                    ag->tree = hk_oracle_tree_create(pl->g, &root);
                    hk_oracle_tree_set_key(ag->tree, seed + (uint64_t)id);
                    fresh = 1;
                } else if (ag->cycles < pl->mp.reuse_cycles) {                                        /* :265 */
                    fresh = 0;
                } else continue;
                const uint64_t key = hk_oracle_tree_key(ag->tree);
                if (hk_oracle_tree_search(ag->tree, budget, 0, key, &rng) == -1) {
#pragma omp atomic write
                    failed = 1;
                }
                ag->n_best = hk_oracle_tree_best_states(ag->tree, 0, key, &rng, ag->best, HK_MCTS_MAX_SEQ);
                ag->pending = 1; ag->pending_fresh = fresh;
            }
            pl->pending_step = step + pl->mp.apply_delay;
        }
        if (pl->pending_step == step) {                                                               /* the thread finishes; FixedUpdate hands off */
            for (int id = 0; id < pl->n_agents; ++id) {
                oagent* ag = &pl->a[id];
                if (!ag->pending) continue;
                const int r2 = id & ~1, ego = id & 1;
                ag->root_valid = 1;                                                                   /* currentRoot = ... (:250, :271) */
                ag->cycles = ag->pending_fresh ? 1 : ag->cycles + 1;                                  /* :253, :273 */
                hk_oracle_race_apply_best(n_sections, karts + r2, plans + r2, ego, ag->nearby, ag->best, ag->n_best);
                ag->pending = 0;
            }
            pl->pending_step = -1;
        }
#pragma omp parallel for reduction(+ : bad) schedule(static)
        for (int r = 0; r < n_races; ++r) {
            double u[4];
            int sec_before[2] = {karts[2 * r].section, karts[2 * r + 1].section};
            for (int e = 0; e < 2; ++e) {
                double x0[8], target[8], tw[8], cw[2], aw[4], otgt[8], otw[6];
                hk_oracle_race_recipe_one(sections, trig, fwd, lane, n_sections, p, karts + 2 * r, plans + 2 * r, e, x0, target, tw,
                                          cw, aw, otgt, otw);
                double A[2 * 16], B[2 * 8], Q[2 * 64], q[2 * 8], R[2 * 4], u0[4];
                for (int i = 0; i < 2; ++i) {
                    hk_oracle_bicycle_A(p->dt, x0 + 4 * i, A + 16 * i);
                    hk_oracle_bicycle_B(p->dt, B + 8 * i);
                    hk_oracle_cost(1, target + 4 * i, tw + 4 * i, cw[i], aw + 2 * i, otgt + 4 * i, otw + 3 * i, Q + 64 * i, q + 8 * i,
                                   R + 4 * i);
                }
                bad += hk_oracle_lqng_solve(2, p->horizon, 0, A, B, Q, q, R, x0, u0, NULL, NULL, NULL) != 0;
                u[2 * e] = u0[0]; u[2 * e + 1] = u0[1];
            }
            if (u_last) memcpy(u_last + 4 * (size_t)r, u, sizeof(u));
            hk_oracle_race_step(sections, trig, fwd, lane, n_sections, p, 2, step, u, karts + 2 * r, plans + 2 * r);
            for (int e = 0; e < 2; ++e)
                if (karts[2 * r + e].section != sec_before[e]) { pl->a[2 * r + e].root_valid = 0; pl->a[2 * r + e].cycles = 0; }   /* :660-661 */
        }
    }
    return failed ? -1 : bad;
}
