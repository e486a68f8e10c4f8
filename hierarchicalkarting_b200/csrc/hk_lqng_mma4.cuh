// hk_lqng_mma4.cuh — warp-per-problem kernel for the 4-kart LQNG (n = 16, m = 8; BASELINE config 3), any operands the ABI
// accepts (time-varying, non-symmetric Q, all outputs).  Same algorithm as KartLQR.solveFeedbackLQR
// (reference: Assets/Karting/Scripts/AI/LQR/KartLQR.cs:17-128) with quirks Q1/Q2 of SURVEY.md A.3.
//
// Where the flops are: Z_i <- Q_i + P_i^T R_i P_i + F^T Z_i F for four players is 2 x 4 x 16^3 x 2 = 65 k of the 139 k flops
// of a backward step.  Those products run on the FP64 tensor pipe as DMMA m8n8k4 over 2 x 2 tiles of 8 x 8 with operand
// fragments read straight from shared memory.  The 16 x 16 arrays are row-major with leading dimension 16 and the four 4-column
// groups of row r XOR-swizzled by r & 3 (element (r, c) at r*16 + (c ^ 4 (r & 3))): fragment loads (lane (g, t) reads (g, 4 ks + t))
// are the minimal two wavefronts, row reads stay contiguous, and a warp needs 17.3 KB instead of the 20.3 KB of a padded layout
// (leading dimension 20) — 12 resident warps per SM instead of 10.  F and Y = Z_i F are kept TRANSPOSED (Ft, Yt) so that both their uses — as the k x n operand of one product and
// the m x k operand of the next — are row reads.  The coupled 8 x 8 system [LHS | RHSMat | RHSVec] (25 columns) is dealt one
// column per lane and reduced in registers by Gauss-Jordan with partial pivoting (the pivot MathNet's LU would take: first
// largest magnitude at or below the diagonal, KartLQR.cs:104-105), the pivot column broadcast by shuffles.  Everything else
// (W = B_i^T Z_i, F = A - B P, eta, the rollout) is lane-per-row / lane-per-column scalar code on the same shared arrays.
#pragma once

namespace hk {

struct Mma4Layout {
    static constexpr int LD = 16;
    static constexpr int oZ = 0;                     // Z[4][16][16], swizzled
    static constexpr int oFt = oZ + 4 * 16 * LD;     // Ft[c][r] = F[r][c], swizzled
    static constexpr int oYt = oFt + 16 * LD;        // Yt[c][k] = (Z_i F)[k][c], swizzled
    static constexpr int oW = oYt;                   // W[8][16] = stacked B_i^T Z_i: dead before the first Y of a step is stored
    static constexpr int oP = oYt + 16 * LD;         // P[8][16]
    static constexpr int oA = oP + 128;              // A_i[4][4][4]
    static constexpr int oB = oA + 64;               // B_i[4][4][2]
    static constexpr int oR = oB + 32;               // R_i[4][2][2]
    static constexpr int oEta = oR + 16;             // eta[4][16]
    static constexpr int oTmp = oEta + 64;           // q_i[4][16] of the current stage (the eta update reads it on its critical path)
    static constexpr int oBeta = oTmp + 64;
    static constexpr int oAlpha = oBeta + 16;
    static constexpr int oX = oAlpha + 8;
    static constexpr int oU = oX + 16;
    static constexpr int total = oU + 8;
    static constexpr int LDS = 26;                   // coupled-system scratch [8][26] inside the Ft buffer
    // element (r, c) of a swizzled 16 x 16 array
    __host__ __device__ static constexpr int at(int r, int c) { return r * 16 + (c ^ ((r & 3) << 2)); }
};

constexpr int MMA4_WARPS = 2;

// NP = players of the game (3 or 4).  A 3-kart game runs in the same 16 x 16 / 8 x 8 frame with a decoupled dummy fourth player
// (B_3 = 0, Q_3 = 0, q_3 = 0, A_3 = R_3 = I): its gains, its Z and its eta stay exactly zero, the coupled system is block diagonal
// with an identity block (pivot choices of the real 6 x 6 system unchanged), and its pass through the player loop is skipped.
template <int NP>
__global__ void __launch_bounds__(32 * MMA4_WARPS, 7) lqng_mma4_kernel(LqngParams p)
{
    using L = Mma4Layout;
    constexpr int N = 4, n = 16, m = 8, LD = L::LD;
    constexpr int rn = 4 * NP, rm = 2 * NP;                         // real joint state / control dimensions (record strides)
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3, lo = lane & 15, hf = lane >> 4;
    double* s = smem + (size_t)wib * L::total;
    double *Z = s + L::oZ, *Ft = s + L::oFt, *Yt = s + L::oYt, *Pm = s + L::oP, *W = s + L::oW, *Sy = s + L::oFt, *As = s + L::oA,
           *Bs = s + L::oB, *Rs = s + L::oR, *eta = s + L::oEta, *qs = s + L::oTmp, *beta = s + L::oBeta, *alpha = s + L::oAlpha,
           *xs = s + L::oX, *us = s + L::oU;
    const long long nwarps = (long long)gridDim.x * MMA4_WARPS;
    // fragment element (g, 4 ks + t) of a swizzled array sits at fb ^ (4 ks); rows 8 + g at + 128
    const int fb = 16 * g + 4 * (g & 3) + t;
    const int f0 = fb, f1 = fb ^ 4, f2 = fb ^ 8, f3 = fb ^ 12;
    const int bo = g < 4 ? 4 * t + g : 64 + 2 * t + (g - 4), bs = g < 4 ? 16 : 8;   // B operand [A_pc | B_pc | 0][t][g] relative to As (Bs follows As)
    const int zo = L::at(g, 2 * t);                                 // C fragment (g, 2t..2t+1); columns 8 + 2t at ^ 8; rows 8 + g at + 128
    const int yo0 = L::at(2 * t, g), yo1 = L::at(2 * t + 1, g);     // transposed store of a C fragment

    for (long long prob = (long long)blockIdx.x * MMA4_WARPS + wib; prob < p.batch; prob += nwarps) {
        if (p.gate && p.gate[prob] < p.gate_min) continue;         // solved by another kernel (warp-uniform)
        const int T = p.horizon + 1, Tm = p.time_varying ? T : 1;
        const double* gA = p.A + (size_t)prob * Tm * NP * 16;
        const double* gB = p.B + (size_t)prob * Tm * NP * 8;
        const double* gQ = p.Q + (size_t)prob * Tm * NP * rn * rn;
        const double* gq = p.q + (size_t)prob * Tm * NP * rn;
        const double* gR = p.R + (size_t)prob * Tm * NP * 4;
        const double* gx = p.x0 + (size_t)prob * rn;
        double* gP = p.P ? p.P + (size_t)prob * T * rm * rn : nullptr;
        double* ga = p.alpha ? p.alpha + (size_t)prob * T * rm : nullptr;
        int singular = 0;
        if (prob + nwarps < p.batch && !(p.gate && p.gate[prob + nwarps] < p.gate_min)) {   // this warp's next record into L2 while this one is solved
            const size_t nx = (size_t)(prob + nwarps) * Tm;
            const char* nq = reinterpret_cast<const char*>(p.Q + nx * NP * rn * rn);
            constexpr int q_lines = (NP * rn * rn * 8 + 127) / 128;                      // 64 (4 karts) / 27 (3 karts)
            if (lane < q_lines) asm volatile("prefetch.global.L2 [%0];" ::"l"(nq + 128 * lane));
            if (32 + lane < q_lines) asm volatile("prefetch.global.L2 [%0];" ::"l"(nq + 128 * (32 + lane)));
            const char* no = lane < 4 ? reinterpret_cast<const char*>(p.A + nx * NP * 16) + 128 * lane
                           : lane < 6 ? reinterpret_cast<const char*>(p.B + nx * NP * 8) + 128 * (lane - 4)
                           : lane < 10 ? reinterpret_cast<const char*>(p.q + nx * NP * rn) + 128 * (lane - 6)
                           : lane == 10 ? reinterpret_cast<const char*>(p.R + nx * NP * 4) : reinterpret_cast<const char*>(p.x0 + (size_t)(prob + nwarps) * rn);
            const bool in = lane < 4 ? 128 * lane < NP * 128 : lane < 6 ? 128 * (lane - 4) < NP * 64 : lane < 10 ? 128 * (lane - 6) < NP * rn * 8 : true;
            if (lane < 12 && in) asm volatile("prefetch.global.L2 [%0];" ::"l"(no));
        }
        __syncwarp();
        {   // Zs = Q, etas = q of the last stage (KartLQR.cs:62-63): 128-bit coalesced loads
            const double* q0 = gQ + (size_t)(Tm - 1) * NP * rn * rn;
            if (NP == 4) {
                for (int e = lane; e < N * n * n / 2; e += 32) {
                    const double2 v = *reinterpret_cast<const double2*>(q0 + 2 * e);
                    const int i = e >> 7, r = (e >> 3) & 15, c = (e & 7) * 2;
                    *reinterpret_cast<double2*>(Z + i * 256 + L::at(r, c)) = v;
                }
                for (int e = lane; e < N * n; e += 32) { const double v = gq[(size_t)(Tm - 1) * N * n + e]; eta[e] = v; qs[e] = v; }
            } else {
                for (int e = lane; e < N * n * n / 2; e += 32) *reinterpret_cast<double2*>(Z + 2 * e) = make_double2(0.0, 0.0);
                eta[lane] = 0.0; eta[32 + lane] = 0.0; qs[lane] = 0.0; qs[32 + lane] = 0.0;
                __syncwarp();
                for (int e = lane; e < NP * rn * rn / 2; e += 32) {
                    const double2 v = *reinterpret_cast<const double2*>(q0 + 2 * e);
                    const int i = e / (rn * rn / 2), rem = e % (rn * rn / 2), r = rem / (rn / 2), c = (rem % (rn / 2)) * 2;
                    *reinterpret_cast<double2*>(Z + i * 256 + L::at(r, c)) = v;
                }
                for (int e = lane; e < NP * rn; e += 32) {
                    const double v = gq[(size_t)(Tm - 1) * NP * rn + e];
                    eta[(e / rn) * n + e % rn] = v; qs[(e / rn) * n + e % rn] = v;
                }
                // the dummy player: A_3 = I, B_3 = 0, R_3 = I
                if (lane < 16) As[NP * 16 + lane] = (lane >> 2) == (lane & 3) ? 1.0 : 0.0;
                if (lane < 8) Bs[NP * 8 + lane] = 0.0;
                if (lane < 4) Rs[NP * 4 + lane] = (lane == 0 || lane == 3) ? 1.0 : 0.0;
            }
            if (lane < n) xs[lane] = lane < rn ? gx[lane] : 0.0;
        }
        for (int st = p.horizon; st >= 0; --st) {                   // KartLQR.cs:64
            const int tt = p.time_varying ? st : 0;
            const double* Qt = gQ + (size_t)tt * NP * rn * rn;
            if (p.time_varying || st == p.horizon) {
                if (st != p.horizon)
                    for (int e = lane; e < NP * rn; e += 32) qs[(e / rn) * n + e % rn] = gq[(size_t)tt * NP * rn + e];
                for (int e = lane; e < NP * 16; e += 32) As[e] = gA[(size_t)tt * NP * 16 + e];
                if (lane < NP * 8) Bs[lane] = gB[(size_t)tt * NP * 8 + lane];
                if (lane < NP * 4) Rs[lane] = gR[(size_t)tt * NP * 4 + lane];
            }
            __syncwarp();
            // W_i = B_i^T Z_i (rows of block i only: B_i is zero elsewhere, KartLQR.cs:41-52); lane owns column lo of two players
#pragma unroll
            for (int ii = 0; ii < 2; ++ii) {
                const int i = 2 * hf + ii;
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    double acc = 0.0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc = fma(Bs[i * 8 + k * 2 + a], Z[i * 256 + (4 * i + k) * 16 + (lo ^ (4 * k))], acc);
                    W[(2 * i + a) * 16 + (lo ^ (4 * (2 * ii + a)))] = acc;              // swizzled like Z: fragment reads below
                }
            }
            if (lane < 8) {                                         // RHSVec (:96): row r = (i, a) is B_i[:, a] . eta_i[4i..4i+3]
                const int i = lane >> 1, a = lane & 1;
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < 4; ++k) acc = fma(Bs[i * 8 + k * 2 + a], eta[i * n + 4 * i + k], acc);
                Sy[lane * L::LDS + 24] = acc;
            }
            __syncwarp();
            // Coupled system [LHS | RHSMat | RHSVec] (:68-96): one DMMA per player block pc gives W[:, 4pc..4pc+3] [A_pc | B_pc] —
            // columns 0..3 are RHSMat[:, 4pc..4pc+3], columns 4, 5 are M[(i, a)][(pc, b)] = LHS[(pc, a)][(i, b)] (quirk Q1 placement;
            // + R_pc on the diagonal block, :78).  The D fragments go through a scratch array (the F^T buffer, dead here) to reach the
            // column-per-lane distribution of the Gauss-Jordan sweep.
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) {
                const double wa = W[pc == 0 ? f0 : pc == 1 ? f1 : pc == 2 ? f2 : f3];
                const double ab = g < 6 ? As[bo + pc * bs] : 0.0;
                double d0 = 0.0, d1 = 0.0;
                dmma(d0, d1, wa, ab);
                if (t == 2 && (g >> 1) == pc) {
                    const double2 r2 = *reinterpret_cast<const double2*>(Rs + pc * 4 + (g & 1) * 2);
                    d0 = r2.x + d0; d1 = r2.y + d1;
                }
                if (t < 2) *reinterpret_cast<double2*>(Sy + g * L::LDS + 8 + 4 * pc + 2 * t) = make_double2(d0, d1);
                else if (t == 2) *reinterpret_cast<double2*>(Sy + (2 * pc + (g & 1)) * L::LDS + 2 * (g >> 1)) = make_double2(d0, d1);
            }
            __syncwarp();
            double col[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) col[r] = lane < 25 ? Sy[r * L::LDS + lane] : 0.0;
            // P = LHS.Solve(RHSMat), alpha = LHS.Solve(RHSVec): Gauss-Jordan with partial pivoting (KartLQR.cs:104-105)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                double pc[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) pc[r] = shfl_d(col[r], k);
                // common case (R dominates the LHS): the high word of |pc[k]| alone exceeds every candidate below it — no exchange
                unsigned below = 0;
#pragma unroll
                for (int r = k + 1; r < 8; ++r) below = max(below, (unsigned)__double2hiint(pc[r]) & 0x7fffffffu);
                if (((unsigned)__double2hiint(pc[k]) & 0x7fffffffu) <= below) {              // warp-uniform (pc[] is a broadcast)
                    int pr = k;
                    double best = pc[k];
#pragma unroll
                    for (int r = k + 1; r < 8; ++r)
                        if (abs_gt(pc[r], best)) { best = pc[r]; pr = r; }  // |.| compared on the integer pipes (exact)
#pragma unroll
                    for (int r = k + 1; r < 8; ++r)
                        if (pr == r) {                              // warp-uniform row exchange
                            double tv = col[k]; col[k] = col[r]; col[r] = tv;
                            tv = pc[k]; pc[k] = pc[r]; pc[r] = tv;
                        }
                }
                if (pc[k] == 0.0) singular = 1;
                // 1 / pivot: MUFU seed + one cubic step (2^-60) unless the pivot is zero, denormal or huge (warp-uniform)
                const double rinv = bad_pivot(pc[k]) ? 1.0 / pc[k] : rcp_fast(pc[k]);
                const double rk = col[k] * rinv;
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (r != k) col[r] = fma(-pc[r], rk, col[r]);
                col[k] = rk;
            }
            if (lane >= 8 && lane < 24) {
#pragma unroll
                for (int r = 0; r < 8; ++r) Pm[r * n + lane - 8] = col[r];
                if (gP && lane - 8 < rn)
#pragma unroll
                    for (int r = 0; r < rm; ++r) gP[(size_t)st * rm * rn + r * rn + lane - 8] = col[r];
            } else if (lane == 24) {
#pragma unroll
                for (int r = 0; r < 8; ++r) alpha[r] = col[r];
                if (ga)
#pragma unroll
                    for (int r = 0; r < rm; ++r) ga[(size_t)st * rm + r] = col[r];
            }
            __syncwarp();
            // F = A - sum_k B_k P_k (stored transposed), beta = -sum_k B_k alpha_k (:110-111)
            {
                const int r = lo, pr = r >> 2, rr = r & 3;
                const double b0 = Bs[pr * 8 + rr * 2 + 0], b1 = Bs[pr * 8 + rr * 2 + 1];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = hf * 8 + j;
                    const double a = (c >> 2) == pr ? As[pr * 16 + rr * 4 + (c & 3)] : 0.0;
                    Ft[c * 16 + (r ^ (4 * (j & 3)))] = a - fma(b1, Pm[(2 * pr + 1) * n + c], b0 * Pm[(2 * pr) * n + c]);
                }
                if (hf == 0) beta[r] = -fma(b1, alpha[2 * pr + 1], b0 * alpha[2 * pr]);
            }
            __syncwarp();
            // Z_i <- Q_i + P_i^T R_i P_i + F^T Z_i F ; eta_i <- q_i + P_i^T R_i alpha_i + F^T (eta_i + Z_i^{new} beta)  (:116-117)
            // F^T fragments (Ft[g][4 ks + t], Ft[8 + g][4 ks + t]) serve all four players: B operand of Y = Z_i F, A operand of F^T Y,
            // and the F^T (eta_i + Z_i beta) product of the eta update
            const double fa0 = Ft[f0], fa1 = Ft[f1], fa2 = Ft[f2], fa3 = Ft[f3];
            const double fb0 = Ft[128 + f0], fb1 = Ft[128 + f1], fb2 = Ft[128 + f2], fb3 = Ft[128 + f3];
            const double2 be0 = *reinterpret_cast<const double2*>(beta + 2 * t), be1 = *reinterpret_cast<const double2*>(beta + 8 + 2 * t);
            for (int i = 0; i < NP; ++i) {
                double* Zi = Z + i * 256;
                // C fragments of the new Z_i start from Q_i: issue the loads before the first product to hide their L2 latency
                const double* Qi = Qt + (size_t)i * rn * rn;
                const double2 zero2 = make_double2(0.0, 0.0);
                const bool cin = 8 + 2 * t < rn, rin = 8 + g < rn;   // rows / columns 12..15 of a 3-kart game are padding
                const double2 q00 = *reinterpret_cast<const double2*>(Qi + g * rn + 2 * t);
                const double2 q01 = cin ? *reinterpret_cast<const double2*>(Qi + g * rn + 8 + 2 * t) : zero2;
                const double2 q10 = rin ? *reinterpret_cast<const double2*>(Qi + (8 + g) * rn + 2 * t) : zero2;
                const double2 q11 = (cin && rin) ? *reinterpret_cast<const double2*>(Qi + (8 + g) * rn + 8 + 2 * t) : zero2;
                {   // Y = Z_i F, stored transposed
                    double c00a = 0, c00b = 0, c01a = 0, c01b = 0, c10a = 0, c10b = 0, c11a = 0, c11b = 0;
#define HK_M4_Y(fk, bl, bh)                                                                       \
                    {                                                                             \
                        const double a0 = Zi[fk], a1 = Zi[128 + fk];                              \
                        dmma(c00a, c00b, a0, bl); dmma(c01a, c01b, a0, bh);                       \
                        dmma(c10a, c10b, a1, bl); dmma(c11a, c11b, a1, bh);                       \
                    }
                    HK_M4_Y(f0, fa0, fb0) HK_M4_Y(f1, fa1, fb1) HK_M4_Y(f2, fa2, fb2) HK_M4_Y(f3, fa3, fb3)
#undef HK_M4_Y
                    __syncwarp();
                    Yt[yo0] = c00a;              Yt[yo1] = c00b;
                    Yt[128 + yo0] = c01a;        Yt[128 + yo1] = c01b;
                    Yt[yo0 ^ 8] = c10a;          Yt[yo1 ^ 8] = c10b;
                    Yt[128 + (yo0 ^ 8)] = c11a;  Yt[128 + (yo1 ^ 8)] = c11b;
                }
                __syncwarp();
                double zl, zh;                                      // partial sums of (Z_i^{new} beta)[g], [8 + g]; zl ends as [8 t + g] on lanes t < 2
                {
                    double c00a = q00.x, c00b = q00.y, c01a = q01.x, c01b = q01.y, c10a = q10.x, c10b = q10.y, c11a = q11.x, c11b = q11.y;
                    {   // + P_i^T (R_i P_i): k = 2; lanes t < 2 hold row t of P_i and form row t of R_i P_i
                        const double p0l = Pm[(2 * i) * n + g], p1l = Pm[(2 * i + 1) * n + g], p0h = Pm[(2 * i) * n + 8 + g], p1h = Pm[(2 * i + 1) * n + 8 + g];
                        const double2 rt = *reinterpret_cast<const double2*>(Rs + i * 4 + 2 * (t & 1));
                        const double a0 = t < 2 ? (t ? p1l : p0l) : 0.0, a1 = t < 2 ? (t ? p1h : p0h) : 0.0;
                        const double b0 = t < 2 ? fma(rt.y, p1l, rt.x * p0l) : 0.0, b1 = t < 2 ? fma(rt.y, p1h, rt.x * p0h) : 0.0;
                        dmma(c00a, c00b, a0, b0); dmma(c01a, c01b, a0, b1);
                        dmma(c10a, c10b, a1, b0); dmma(c11a, c11b, a1, b1);
                    }
#define HK_M4_Z(fk, al, ah)                                                                       \
                    {                                                                             \
                        const double b0 = Yt[fk], b1 = Yt[128 + fk];                              \
                        dmma(c00a, c00b, al, b0); dmma(c01a, c01b, al, b1);                       \
                        dmma(c10a, c10b, ah, b0); dmma(c11a, c11b, ah, b1);                       \
                    }
                    HK_M4_Z(f0, fa0, fb0) HK_M4_Z(f1, fa1, fb1) HK_M4_Z(f2, fa2, fb2) HK_M4_Z(f3, fa3, fb3)      // + F^T Y
#undef HK_M4_Z
                    *reinterpret_cast<double2*>(Zi + zo) = make_double2(c00a, c00b);
                    *reinterpret_cast<double2*>(Zi + (zo ^ 8)) = make_double2(c01a, c01b);
                    *reinterpret_cast<double2*>(Zi + 128 + zo) = make_double2(c10a, c10b);
                    *reinterpret_cast<double2*>(Zi + 128 + (zo ^ 8)) = make_double2(c11a, c11b);
                    // Z_i^{new} beta (quirk Q2) straight from the C fragments: lane (g, t) holds columns 2t, 2t+1, 8+2t, 8+2t+1 of rows g, 8+g
                    zl = fma(c01b, be1.y, fma(c01a, be1.x, fma(c00b, be0.y, c00a * be0.x)));
                    zh = fma(c11b, be1.y, fma(c11a, be1.x, fma(c10b, be0.y, c10a * be0.x)));
                    // quad reduction of both sums with two exchanges: odd lanes collect zh, even lanes zl; lane t = 0 ends with all of zl,
                    // lane t = 1 with all of zh
                    zl = (t & 1 ? zh : zl) + __shfl_xor_sync(0xffffffffu, t & 1 ? zl : zh, 1);
                    zl += __shfl_xor_sync(0xffffffffu, zl, 2);
                }
                if (t < 2) eta[i * n + 8 * t + g] += zl;                                         // eta_i + Z_i^{new} beta, in place
                __syncwarp();
                {   // F^T (eta_i + Z_i beta): rows g and 8 + g of F^T from the fragments, reduced over the quad; the new eta_i below
                    // overwrites what these loads read
                    const double v0 = eta[i * n + t], v1 = eta[i * n + 4 + t], v2 = eta[i * n + 8 + t], v3 = eta[i * n + 12 + t];
                    double al = fma(fa3, v3, fma(fa2, v2, fma(fa1, v1, fa0 * v0)));
                    double ah = fma(fb3, v3, fma(fb2, v2, fma(fb1, v1, fb0 * v0)));
                    al = (t & 1 ? ah : al) + __shfl_xor_sync(0xffffffffu, t & 1 ? al : ah, 1);
                    al += __shfl_xor_sync(0xffffffffu, al, 2);
                    __syncwarp();                                   // every lane has read eta_i + Z_i beta before it is overwritten
                    if (t < 2) {
                        const int c = 8 * t + g;
                        const double a0 = alpha[2 * i], a1 = alpha[2 * i + 1];
                        const double ra0 = fma(Rs[i * 4 + 1], a1, Rs[i * 4 + 0] * a0), ra1 = fma(Rs[i * 4 + 3], a1, Rs[i * 4 + 2] * a0);
                        const double pra = fma(Pm[(2 * i + 1) * n + c], ra1, Pm[(2 * i) * n + c] * ra0);
                        eta[i * n + c] = (qs[i * n + c] + pra) + al;
                    }
                }
            }
            __syncwarp();
        }
        // optimal_control = -P x0 - alpha with the t = 0 pair (:121-126), every player
        if (lane < rm) {
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < rn; ++c) acc = fma(-Pm[lane * n + c], xs[c], acc);
            p.u0[(size_t)prob * rm + lane] = acc - alpha[lane];
        }
        if (lane == 0 && p.status) p.status[prob] = singular;
        // closed-loop rollout; gains are re-read from the P/alpha output buffers written above
        if (p.traj) {
            double* gt = p.traj + (size_t)prob * (T + 1) * rn;
            if (lane < rn) gt[lane] = xs[lane];
            __syncwarp();
            for (int st = 0; st <= p.horizon; ++st) {
                const int tt = p.time_varying ? st : 0;
                if (lane < rm) {
                    double acc = 0.0;
                    for (int c = 0; c < rn; ++c) acc = fma(-gP[(size_t)st * rm * rn + lane * rn + c], xs[c], acc);
                    us[lane] = acc - ga[(size_t)st * rm + lane];
                }
                __syncwarp();
                double xn = 0.0;
                if (lane < rn) {
                    const int pr = lane >> 2, rr = lane & 3;
                    for (int c = 0; c < 4; ++c) xn = fma(gA[(size_t)tt * NP * 16 + pr * 16 + rr * 4 + c], xs[4 * pr + c], xn);
                    for (int c = 0; c < 2; ++c) xn = fma(gB[(size_t)tt * NP * 8 + pr * 8 + rr * 2 + c], us[2 * pr + c], xn);
                }
                __syncwarp();
                if (lane < rn) { xs[lane] = xn; gt[(size_t)(st + 1) * rn + lane] = xn; }
                __syncwarp();
            }
        }
    }
}

}  // namespace hk
