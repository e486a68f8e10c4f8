// hk_lqng.cu — batched feedback LQ Nash game solver for sm_100a.
//
// Replaces the body of KartLQR.solveFeedbackLQR (reference: Assets/Karting/Scripts/AI/LQR/KartLQR.cs:17-128):
// coupled-Riccati backward recursion, one coupled N-player m x m solve per step (KartLQR.cs:104-105), quirks Q1/Q2
// of SURVEY.md A.3 reproduced (block placement of the coupled LHS, eta update with the already updated Z_i).
//
// Kernels in this file
//   lqng_generic_kernel<N>   any N in 1..4, time-varying or not, general (also non-symmetric) Q, all outputs.
//                            G = 4/8/16 lanes per problem, matrices staged in shared memory, partial-pivot elimination.
//                            This is the robust path every other kernel falls back to.
//   (hk_lqng_mma.cuh)        warp-per-problem FP64 tensor-core (DMMA m8n8k4) kernels for the time-invariant 2- and
//                            4-kart games — the throughput path, see DESIGN.md §4.
//
// Record layout of the operands = include/hk_abi.h (problem-major: one warp reads one problem's record with fully
// coalesced 128-bit loads; there is no cross-problem reuse, so HBM traffic = algorithmic bytes).
#include "hk_common.cuh"
#include <cstdlib>
#include <atomic>
#include <mutex>

namespace hk {

struct LqngParams {
    int batch, horizon, time_varying;
    const double *A, *B, *Q, *q, *R, *x0;
    double *u0, *P, *alpha, *traj;
    int* status;
    const int* redo_list;      // non-null: solve only problems redo_list[0 .. *redo_count) (queued by a fast kernel)
    const int* redo_count;
    int* reset_count;          // counter of the NEXT launch, cleared here (stream order makes that safe)
    int* work;                 // persistent 2-kart kernel: {next problem - resident warps, warps done}; zero between launches
    // compact (fused assembly) mode of the 2-kart kernel: hk_lqng_assemble_solve_batch's arrays + (cos h, sin h) per player
    const double *c_x0, *c_target, *c_tw, *c_cw, *c_aw, *c_otgt, *c_otw, *c_cs;
    double dt;
    // 3-/4-kart kernel: solve only problems with gate[prob] >= gate_min (hk_raceN_run sends games of one or two players to the 2-kart kernel)
    const int* gate = nullptr;
    int gate_min = 0;
};

template <int N>
struct GenericLayout {
    static constexpr int n = 4 * N, m = 2 * N;
    static constexpr int G = n <= 4 ? 4 : (n <= 8 ? 8 : 16);     // lanes per problem
    static constexpr int LD = n + 1;                              // padded leading dimension: conflict-free column reads
    static constexpr int AW = m + n + 1;                          // [LHS | RHSMat | RHSVec]
    // offsets in doubles
    static constexpr int oZ = 0;
    static constexpr int oF = oZ + N * n * LD;
    static constexpr int oY = oF + n * LD;
    static constexpr int oM = oY + n * LD;
    static constexpr int oW = oM + m * AW;
    static constexpr int oA = oW + m * n;
    static constexpr int oB = oA + N * 16;
    static constexpr int oR = oB + N * 8;
    static constexpr int oEta = oR + N * 4;
    static constexpr int oBeta = oEta + N * n;
    static constexpr int oTmp = oBeta + n;
    static constexpr int oX = oTmp + n;
    static constexpr int oU = oX + n;
    static constexpr int total = oU + m + 1;
};

template <int N>
__device__ __noinline__ void lqng_generic_body(const LqngParams& p, long long prob, const bool live, double* s, const int r, const unsigned mask)
{
    using L = GenericLayout<N>;
    constexpr int n = L::n, m = L::m, G = L::G, LD = L::LD, AW = L::AW;
    double *Z = s + L::oZ, *F = s + L::oF, *Y = s + L::oY, *M = s + L::oM, *W = s + L::oW, *As = s + L::oA, *Bs = s + L::oB,
           *Rs = s + L::oR, *eta = s + L::oEta, *beta = s + L::oBeta, *tmp = s + L::oTmp, *xs = s + L::oX, *us = s + L::oU;

    const int T = p.horizon + 1, Tm = p.time_varying ? T : 1;
    const double* gA = p.A + (size_t)prob * Tm * N * 16;
    const double* gB = p.B + (size_t)prob * Tm * N * 8;
    const double* gQ = p.Q + (size_t)prob * Tm * N * n * n;
    const double* gq = p.q + (size_t)prob * Tm * N * n;
    const double* gR = p.R + (size_t)prob * Tm * N * 4;
    const double* gx = p.x0 + (size_t)prob * n;
    double* gP = p.P ? p.P + (size_t)prob * T * m * n : nullptr;
    double* ga = p.alpha ? p.alpha + (size_t)prob * T * m : nullptr;
    int singular = 0;

    {   // Zs = Q, etas = q of the last stage (KartLQR.cs:62-63)
        const int tl = Tm - 1;
        for (int e = r; e < N * n * n; e += G) {
            int i = e / (n * n), rc = e % (n * n);
            Z[i * n * LD + (rc / n) * LD + rc % n] = gQ[(size_t)tl * N * n * n + e];
        }
        for (int e = r; e < N * n; e += G) eta[e] = gq[(size_t)tl * N * n + e];
        for (int e = r; e < n; e += G) xs[e] = gx[e];
    }
    for (int t = p.horizon; t >= 0; --t) {                       // KartLQR.cs:64
        const int tt = p.time_varying ? t : 0;
        if (p.time_varying || t == p.horizon) {
            for (int e = r; e < N * 16; e += G) As[e] = gA[(size_t)tt * N * 16 + e];
            for (int e = r; e < N * 8; e += G) Bs[e] = gB[(size_t)tt * N * 8 + e];
            for (int e = r; e < N * 4; e += G) Rs[e] = gR[(size_t)tt * N * 4 + e];
        }
        __syncwarp(mask);
        // W_i = B_i^T Z_i (rows of block i only: B_i is zero elsewhere, KartLQR.cs:41-52); lane c owns column c
        if (r < n) {
            const int c = r;
#pragma unroll
            for (int i = 0; i < N; ++i)
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    double acc = 0.0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc = fma(Bs[i * 8 + k * 2 + a], Z[i * n * LD + (4 * i + k) * LD + c], acc);
                    W[(2 * i + a) * n + c] = acc;
                }
        }
        __syncwarp(mask);
        // coupled system [LHS | RHSMat | RHSVec]; block (i,j) -> row-block j, column-block i (quirk Q1, KartLQR.cs:68-87)
        for (int e = r; e < m * m; e += G) {
            const int ia = e / m, jb = e % m;                     // raw[2i+a][2j+b] = (B_i^T Z_i B_j)[a][b]
            const int i = ia / 2, a = ia % 2, j = jb / 2, b = jb % 2;
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc = fma(W[ia * n + 4 * j + k], Bs[j * 8 + k * 2 + b], acc);
            if (i == j) acc = Rs[i * 4 + a * 2 + b] + acc;        // getRMatrix() + ... (:78)
            M[(2 * j + a) * AW + (2 * i + b)] = acc;
        }
        if (r < n) {                                              // RHSMat = vstack_i B_i^T Z_i A (:89-95), A block diagonal
            const int c = r, pc = c / 4, cc = c % 4;
#pragma unroll
            for (int ia = 0; ia < m; ++ia) {
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < 4; ++k) acc = fma(W[ia * n + 4 * pc + k], As[pc * 16 + k * 4 + cc], acc);
                M[ia * AW + m + c] = acc;
            }
        }
        if (r < m) {                                              // RHSVec = concat_i B_i^T eta_i (:96)
            const int i = r / 2, a = r % 2;
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc = fma(Bs[i * 8 + k * 2 + a], eta[i * n + 4 * i + k], acc);
            M[r * AW + m + n] = acc;
        }
        __syncwarp(mask);
        // P = LHS.Solve(RHSMat), alpha = LHS.Solve(RHSVec): LU with partial pivoting (MathNet, KartLQR.cs:104-105)
        for (int k = 0; k < m; ++k) {
            int pr = k;
            double best = fabs(M[k * AW + k]);
            for (int i = k + 1; i < m; ++i) {
                double v = fabs(M[i * AW + k]);
                if (v > best) { best = v; pr = i; }
            }
            __syncwarp(mask);
            if (pr != k)
                for (int c = r; c < AW; c += G) { double t0 = M[k * AW + c]; M[k * AW + c] = M[pr * AW + c]; M[pr * AW + c] = t0; }
            __syncwarp(mask);
            const double piv = M[k * AW + k];
            if (piv == 0.0) singular = 1;
            for (int i = k + 1; i < m; ++i) {
                const double l = M[i * AW + k] / piv;
                for (int c = k + 1 + r; c < AW; c += G) M[i * AW + c] = fma(-l, M[k * AW + c], M[i * AW + c]);
            }
            __syncwarp(mask);
        }
        for (int c = m + r; c < AW; c += G)
            for (int k = m - 1; k >= 0; --k) {
                double x = M[k * AW + c];
                for (int j = k + 1; j < m; ++j) x = fma(-M[k * AW + j], M[j * AW + c], x);
                M[k * AW + c] = x / M[k * AW + k];
            }
        __syncwarp(mask);
        // from here P[k][c] = M[k][m+c], alpha[k] = M[k][m+n]
        if (live) {
            if (gP) for (int e = r; e < m * n; e += G) gP[(size_t)t * m * n + e] = M[(e / n) * AW + m + e % n];
            if (ga) for (int e = r; e < m; e += G) ga[(size_t)t * m + e] = M[e * AW + m + n];
        }
        // F = A - sum_k B_k P_k, beta = -sum_k B_k alpha_k (:110-111); lane r owns row r
        if (r < n) {
            const int pr = r / 4, rr = r % 4;
            const double b0 = Bs[pr * 8 + rr * 2 + 0], b1 = Bs[pr * 8 + rr * 2 + 1];
#pragma unroll
            for (int c = 0; c < n; ++c) {
                const double a = (c / 4 == pr) ? As[pr * 16 + rr * 4 + c % 4] : 0.0;
                const double bp = fma(b1, M[(2 * pr + 1) * AW + m + c], b0 * M[(2 * pr) * AW + m + c]);
                F[r * LD + c] = a - bp;
            }
            beta[r] = -fma(b1, M[(2 * pr + 1) * AW + m + n], b0 * M[(2 * pr) * AW + m + n]);
        }
        __syncwarp(mask);
        // Z_i <- Q_i + P_i^T R_i P_i + F^T Z_i F ; eta_i <- q_i + P_i^T R_i alpha_i + F^T (eta_i + Z_i^{new} beta)  (:116-117)
        for (int i = 0; i < N; ++i) {
            double* Zi = Z + i * n * LD;
            const double r00 = Rs[i * 4 + 0], r01 = Rs[i * 4 + 1], r10 = Rs[i * 4 + 2], r11 = Rs[i * 4 + 3];
            if (r < n) {                                          // Y = Z_i F, row r
                double zr[n];
#pragma unroll
                for (int k = 0; k < n; ++k) zr[k] = Zi[r * LD + k];
#pragma unroll
                for (int c = 0; c < n; ++c) {
                    double acc = 0.0;
#pragma unroll
                    for (int k = 0; k < n; ++k) acc = fma(zr[k], F[k * LD + c], acc);
                    Y[r * LD + c] = acc;
                }
            }
            __syncwarp(mask);
            if (r < n) {                                          // row r of the new Z_i
                double fc[n];
#pragma unroll
                for (int k = 0; k < n; ++k) fc[k] = F[k * LD + r];
                const double p0r = M[(2 * i) * AW + m + r], p1r = M[(2 * i + 1) * AW + m + r];
                double zb = 0.0;
#pragma unroll
                for (int c = 0; c < n; ++c) {
                    double acc = 0.0;
#pragma unroll
                    for (int k = 0; k < n; ++k) acc = fma(fc[k], Y[k * LD + c], acc);
                    const double p0c = M[(2 * i) * AW + m + c], p1c = M[(2 * i + 1) * AW + m + c];
                    const double rp0 = fma(r01, p1c, r00 * p0c), rp1 = fma(r11, p1c, r10 * p0c);   // (R_i P_i)[:, c]
                    const double prp = fma(p1r, rp1, p0r * rp0);                                    // (P_i^T R_i P_i)[r][c]
                    const double znew = (gQ[(size_t)tt * N * n * n + (size_t)i * n * n + r * n + c] + prp) + acc;
                    Zi[r * LD + c] = znew;                        // row r is only read by lane r from here on
                    zb = fma(znew, beta[c], zb);
                }
                tmp[r] = eta[i * n + r] + zb;                     // eta_i + Z_i^{new} beta (quirk Q2)
            }
            __syncwarp(mask);
            if (r < n) {
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < n; ++k) acc = fma(F[k * LD + r], tmp[k], acc);
                const double a0 = M[(2 * i) * AW + m + n], a1 = M[(2 * i + 1) * AW + m + n];
                const double ra0 = fma(r01, a1, r00 * a0), ra1 = fma(r11, a1, r10 * a0);
                const double pra = fma(M[(2 * i + 1) * AW + m + r], ra1, M[(2 * i) * AW + m + r] * ra0);
                eta[i * n + r] = (gq[(size_t)tt * N * n + i * n + r] + pra) + acc;
            }
            __syncwarp(mask);
        }
    }
    // optimal_control = -P x0 - alpha with the t = 0 pair (:121-126), every player
    if (r < m) {
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < n; ++c) acc = fma(-M[r * AW + m + c], xs[c], acc);
        const double u = acc - M[r * AW + m + n];
        if (live) p.u0[(size_t)prob * m + r] = u;
    }
    if (live && r == 0 && p.status) p.status[prob] = singular;
    // closed-loop rollout; gains are re-read from the P/alpha output buffers written above
    if (p.traj) {
        double* gt = p.traj + (size_t)prob * (T + 1) * n;
        if (live && r < n) gt[r] = xs[r];
        __syncwarp(mask);
        for (int t = 0; t <= p.horizon; ++t) {
            const int tt = p.time_varying ? t : 0;
            if (r < m) {                                            // groups past the end of the batch (live == false) only keep step
                double acc = 0.0;
                if (live) {
                    for (int c = 0; c < n; ++c) acc = fma(-gP[(size_t)t * m * n + r * n + c], xs[c], acc);
                    acc -= ga[(size_t)t * m + r];
                }
                us[r] = acc;
            }
            __syncwarp(mask);
            double xn = 0.0;
            if (r < n) {
                const int pr = r / 4, rr = r % 4;
                for (int c = 0; c < 4; ++c) xn = fma(gA[(size_t)tt * N * 16 + pr * 16 + rr * 4 + c], xs[4 * pr + c], xn);
                for (int c = 0; c < 2; ++c) xn = fma(gB[(size_t)tt * N * 8 + pr * 8 + rr * 2 + c], us[2 * pr + c], xn);
            }
            __syncwarp(mask);
            if (r < n) { xs[r] = xn; if (live) gt[(size_t)(t + 1) * n + r] = xn; }
            __syncwarp(mask);
        }
    }
}


template <int N>
__global__ void __launch_bounds__(128) lqng_generic_kernel(LqngParams p)
{
    using L = GenericLayout<N>;
    constexpr int G = L::G;
    extern __shared__ double smem[];
    const int gpb = blockDim.x / G;
    const int g = threadIdx.x / G, r = threadIdx.x % G;
    long long count = p.batch;
    if (p.redo_list) {
        count = *p.redo_count;
        if (blockIdx.x == 0 && threadIdx.x == 0 && p.reset_count) *p.reset_count = 0;
    }
    // direct mode: one pass; redo mode: a small grid strides over whatever the fast kernel queued (normally nothing)
    for (long long base = (long long)blockIdx.x * gpb; base < count; base += (long long)gridDim.x * gpb) {
        long long slot = base + g;
        const bool live = slot < count;
        if (!live) slot = count - 1;         // dead groups shadow the last problem so that __syncwarp stays converged
        const long long prob = p.redo_list ? p.redo_list[slot] : slot;
        lqng_generic_body<N>(p, prob, live, smem + (size_t)g * L::total, r, 0xffffffffu);
        __syncwarp();
    }
}

}  // namespace hk
#include "hk_lqng_mma.cuh"
#include "hk_lqng_mma2p.cuh"
#include "hk_lqng_mma4.cuh"
namespace hk {

template <int N>
static int launch_generic(const LqngParams& p, cudaStream_t stream)
{
    using L = GenericLayout<N>;
    const int threads = 128, gpb = threads / L::G;
    const size_t smem = (size_t)gpb * L::total * sizeof(double);
    static bool configured = false;
    if (!configured) {
        HK_CUDA(cudaFuncSetAttribute(lqng_generic_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int blocks = (p.batch + gpb - 1) / gpb;
    if (p.redo_list && blocks > 148) blocks = 148;     // redo mode: grid-stride over the (normally empty) queue
    count_launch(); lqng_generic_kernel<N><<<blocks, threads, smem, stream>>>(p);
    HK_CUDA(cudaGetLastError());
    return HK_OK;
}

// Device-side problem assembly: one thread per (problem, player) expands the compact description
// into the dense record: LinearizedBicycle.getA/getB (KartLQRDynamics.cs:40-62) and
// LQRCheckpointReachAvoidCost.getQMatrix/getQVec/getRMatrix (KartLQRCosts.cs:57-140, quirks Q4/Q5 of SURVEY.md A.3).
__global__ void lqng_assemble_kernel(int batch, int N, double dt, const double* __restrict__ x0, const double* __restrict__ target,
                                     const double* __restrict__ tw, const double* __restrict__ cw, const double* __restrict__ aw,
                                     const double* __restrict__ otgt, const double* __restrict__ otw,
                                     double* A, double* B, double* Q, double* q, double* R, double* xj,
                                     const int* __restrict__ n_players = nullptr, int min_players = 0)
{
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long long)batch * N) return;
    const int n = 4 * N, K = N - 1;
    if (n_players && n_players[id / N] < min_players) return;      // this game goes to another kernel: no record
    if (n_players && (int)(id % N) >= n_players[id / N]) {
        // decoupled dummy player of a game with fewer than N real players (hk_raceN_*: HierarchicalKartAgent.cs:709-725 keeps only the karts
        // within 8 m): A = I, B = 0, Q = 0, q = 0, R = I, x = 0 — its gains, Z and eta stay exactly zero, the coupled system gets an
        // identity block, the real players' results are those of the smaller game
        for (int e = 0; e < 16; ++e) A[id * 16 + e] = (e % 5 == 0) ? 1.0 : 0.0;
        for (int e = 0; e < 8; ++e) B[id * 8 + e] = 0.0;
        R[id * 4 + 0] = 1.0; R[id * 4 + 1] = 0.0; R[id * 4 + 2] = 0.0; R[id * 4 + 3] = 1.0;
        for (int e = 0; e < 4; ++e) xj[id * 4 + e] = 0.0;
        for (int e = 0; e < n * n; ++e) Q[id * n * n + e] = 0.0;
        for (int e = 0; e < n; ++e) q[id * n + e] = 0.0;
        return;
    }
    const double* x = x0 + id * 4;
    double* Ai = A + id * 16;
    for (int e = 0; e < 16; ++e) Ai[e] = (e % 5 == 0) ? 1.0 : 0.0;                 // SparseIdentity (:44)
    const double ch = cos(x[3]), sh = sin(x[3]);
    Ai[0 * 4 + 2] = ch * dt;                                                       // :45
    Ai[1 * 4 + 2] = sh * dt;                                                       // :46
    Ai[0 * 4 + 3] = -sh * dt * x[2];                                               // :47
    Ai[1 * 4 + 3] = ch * dt * x[2];                                                // :48
    double* Bi = B + id * 8;
    for (int e = 0; e < 8; ++e) Bi[e] = 0.0;
    Bi[2 * 2 + 0] = dt; Bi[3 * 2 + 1] = dt;                                        // :58-59
    double* Ri = R + id * 4;
    Ri[0] = cw[id]; Ri[1] = 0.0; Ri[2] = 0.0; Ri[3] = cw[id];                      // KartLQRCosts.cs:136
    for (int e = 0; e < 4; ++e) xj[id * 4 + e] = x[e];
    double* Qi = Q + id * n * n;
    double* qi = q + id * n;
    for (int e = 0; e < n * n; ++e) Qi[e] = 0.0;
    const double* awi = aw + id * K * 2;
    const double* oti = otgt + id * K * 4;
    const double* owi = otw + id * K * 3;
    for (int s = 0; s < 2; ++s) {                                                  // :64-80
        double total = 0.0;
        for (int k = 0; k < K; ++k) {
            const int t = 4 * (1 + k) + s;
            const double w = awi[k * 2 + s];
            Qi[s * n + t] = w; Qi[t * n + s] = w; Qi[t * n + t] = -w;
            total -= w;
        }
        Qi[s * n + s] = total;
    }
    for (int s = 0; s < 4; ++s) Qi[s * n + s] += tw[id * 4 + s];                   // :81-84
    for (int k = 0; k < K; ++k)
        for (int o = 0; o < 3; ++o) Qi[(4 * (1 + k) + o) * n + 4 * (1 + k) + o] = -owi[k * 3 + o];   // :86-94 (assignment)
    for (int s = 0; s < 4; ++s) qi[s] = (-target[id * 4 + s]) * tw[id * 4 + s];    // :109-113
    for (int k = 0; k < K; ++k) {                                                  // :115-124
        for (int s = 0; s < 4; ++s) qi[4 * (1 + k) + s] = oti[k * 4 + s];
        for (int o = 0; o < 3; ++o) qi[4 * (1 + k) + o] = qi[4 * (1 + k) + o] * -owi[k * 3 + o];
    }
}

// Persistent TMA-staged 2-kart kernel (hk_lqng_mma2p.cuh): dense records (p.A .. p.x0) or, with `compact`, the description of
// hk_lqng_assemble_solve_batch (p.c_*), assembled by the warp in shared memory.
static int launch_mma2p(LqngParams p, cudaStream_t stream, bool compact, bool full = false, bool tv = false)
{
    const int batch = p.batch;
    static const int variant_env = getenv("HK_MMA2_VARIANT") ? atoi(getenv("HK_MMA2_VARIANT")) : 2;
    const int variant = variant_env >= 1 ? variant_env : 2;
    // variant 1: 4 warps per CTA, MINB resident CTAs per SM; variant 2 (default): one warp per CTA, MINB resident warps
    // per SM.  16 warps x 128 registers is the measured optimum (profiles/lqng_mma2_tuning_r01.md).
    static const int minb = getenv("HK_MMA2_MINB") ? atoi(getenv("HK_MMA2_MINB")) : (variant == 2 ? 16 : 4);
    const int warps = (compact || full || tv || variant == 2) ? 1 : 4;
    void (*kern)(LqngParams) = nullptr;
    if (compact) {
        kern = lqng_mma2p_kernel<16, 1, true>;
    } else if (tv) {
        kern = full ? lqng_mma2p_kernel<12, 1, false, true, true> : lqng_mma2p_kernel<12, 1, false, false, true>;   // shared memory allows ~11 warps per SM at horizon 3
    } else if (full) {
        kern = lqng_mma2p_kernel<16, 1, false, true>;
    } else if (variant == 2) {
        kern = minb >= 32 ? lqng_mma2p_kernel<32, 1> : minb >= 24 ? lqng_mma2p_kernel<24, 1> : minb >= 20 ? lqng_mma2p_kernel<20, 1>
               : minb >= 18 ? lqng_mma2p_kernel<18, 1> : minb >= 17 ? lqng_mma2p_kernel<17, 1>
               : minb >= 16 ? lqng_mma2p_kernel<16, 1> : lqng_mma2p_kernel<12, 1>;
    } else {
        kern = minb >= 8 ? lqng_mma2p_kernel<8, 4> : minb >= 6 ? lqng_mma2p_kernel<6, 4> : minb == 5 ? lqng_mma2p_kernel<5, 4>
               : lqng_mma2p_kernel<4, 4>;
    }
    // persistent grid: SMs x resident CTAs (dense, compact, full; time-varying per number of stages and output form)
    static int resident_of[3 + 2 * 8] = {};
    int& resident = resident_of[tv ? 3 + 2 * p.horizon + (full ? 1 : 0) : compact ? 1 : full ? 2 : 0];
    static const int pad_env = getenv("HK_MMA2_PAD_SMEM") ? atoi(getenv("HK_MMA2_PAD_SMEM")) : 0;   // occupancy experiments only
    // time-varying: the whole horizon of a problem is staged, double-buffered: 2 x (horizon + 1) records of 1,728 B per warp
    const int pad = tv ? 2 * (p.horizon + 1) * P2_STRIDE * (int)sizeof(double) : pad_env;
    if (!resident) {
        int dev = 0, sms = 0, occ = 0;
        HK_CUDA(cudaGetDevice(&dev));
        HK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (pad) HK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pad > 48 * 1024 ? pad : 48 * 1024));
        HK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * warps, pad));
        resident = sms * (occ > 0 ? occ : 1);
    }
    const long long want = ((long long)batch + warps - 1) / warps;
    const unsigned grid = (unsigned)(want < resident ? want : resident);
    // dynamic work distribution: a pool of zero-initialised counter pairs, one per launch in flight (the kernel's
    // last warp clears its pair); launches on one stream serialise, 64 pairs cover concurrent host threads
    static const bool dyn = !(getenv("HK_MMA2_DYNAMIC") && atoi(getenv("HK_MMA2_DYNAMIC")) == 0);
    static int* pool = nullptr;
    static std::atomic<unsigned> next_slot{0};
    static std::mutex pool_mu;
    if (dyn && !pool) {
        std::lock_guard<std::mutex> lk(pool_mu);
        if (!pool) {
            int* d = nullptr;
            HK_CUDA(cudaMalloc(&d, 64 * 2 * sizeof(int)));
            HK_CUDA(cudaMemset(d, 0, 64 * 2 * sizeof(int)));
            pool = d;
        }
    }
    if (dyn && (long long)grid * warps < batch) p.work = pool + 2 * (next_slot.fetch_add(1) % 64);
    count_launch();
    // Launched with programmatic stream serialization: if the previous kernel on this stream is this same kernel (the
    // only one here that triggers early), the new grid's ramp-up overlaps its tail; after any other kernel or copy the
    // attribute changes nothing.  The kernel waits for the previous grid before its first global store.
    static const bool pdl = !(getenv("HK_MMA2_PDL") && atoi(getenv("HK_MMA2_PDL")) == 0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(32 * warps); cfg.dynamicSmemBytes = (size_t)pad; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    HK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    HK_CUDA(cudaGetLastError());
    return HK_OK;
}

// (cos h, sin h) of every player: the one transcendental of the assembly, evaluated one thread per player (a warp of the
// solve kernel would spend a whole sincos instruction sequence on two useful lanes)
// n_players (optional, per problem): players past it are DUMMY players, marked (cos h, sin h) = (0, 0) — the 2-kart kernel then assembles
// A = I, B = 0 for them (the rest of a dummy's description is zero weights, control weight 1, written by the caller)
__global__ void lqng_trig_kernel(long long n_players_total, const double* __restrict__ x0, double* __restrict__ cs, int N = 1, long long stride = 4,
                                 const int* __restrict__ n_players = nullptr)
{
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_players_total) return;
    if (n_players && (int)(id % N) >= n_players[id / N]) { cs[id * 2] = 0.0; cs[id * 2 + 1] = 0.0; return; }
    const double h = x0[(id / N) * stride + (id % N) * 4 + 3];      // stride: doubles between the x0 blocks of consecutive problems
    cs[id * 2] = cos(h);
    cs[id * 2 + 1] = sin(h);
}

int lqng_assemble_launch(int batch, int N, int horizon, double dt, const double* dx0, const double* dtarget,
                         const double* dtw, const double* dcw, const double* daw, const double* dotgt, const double* dotw,
                         double* du0, int* dstatus, cudaStream_t stream, int scratch_slot, const int* dn_players, int min_players)
{
    ThreadCtx* c = ctx();
    if (!c) return HK_ERR_NO_DEVICE;
    static const bool fused = !(getenv("HK_LQNG_FUSED_ASSEMBLY") && atoi(getenv("HK_LQNG_FUSED_ASSEMBLY")) == 0);
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(dx0) | reinterpret_cast<uintptr_t>(dtarget) | reinterpret_cast<uintptr_t>(dtw) |
                             reinterpret_cast<uintptr_t>(dcw) | reinterpret_cast<uintptr_t>(daw) | reinterpret_cast<uintptr_t>(dotgt) |
                             reinterpret_cast<uintptr_t>(dotw)) & 15) == 0;
    if (N == 2 && fused && aligned16 && batch > 0 && !dn_players) {
        // 2-kart game: no dense records in HBM at all — the solve kernel stages the 352-byte description by TMA and assembles
        // A, B, Q, q, R in shared memory (hk_lqng_mma2p.cuh, COMPACT)
        double* dcs = (double*)dscratch(c, scratch_slot, sizeof(double) * 4 * (size_t)batch);
        if (!dcs) return HK_ERR_OUT_OF_MEMORY;
        const long long np = (long long)batch * 2;
        count_launch(); lqng_trig_kernel<<<(unsigned)((np + 255) / 256), 256, 0, stream>>>(np, dx0, dcs);
        HK_CUDA(cudaGetLastError());
        LqngParams p{batch, horizon, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, du0, nullptr, nullptr, nullptr, dstatus,
                     nullptr, nullptr, nullptr, nullptr, dx0, dtarget, dtw, dcw, daw, dotgt, dotw, dcs, dt};
        return launch_mma2p(p, stream, true);
    }
    const int n = 4 * N;
    const size_t per = (size_t)N * 16 + N * 8 + (size_t)N * n * n + (size_t)N * n + N * 4 + n;
    double* d = (double*)dscratch(c, scratch_slot, per * sizeof(double) * (size_t)batch);
    if (!d) return HK_ERR_OUT_OF_MEMORY;
    double *dA = d, *dB = dA + (size_t)batch * N * 16, *dQ = dB + (size_t)batch * N * 8, *dq = dQ + (size_t)batch * N * n * n,
           *dR = dq + (size_t)batch * N * n, *dx = dR + (size_t)batch * N * 4;
    const long long threads = (long long)batch * N;
    // games below min_players are left out (no record, no solve) when the kernel that honours the gate takes the launch
    const bool gated = dn_players && min_players > 0 && (N == 3 || N == 4) && !(getenv("HK_LQNG_MMA4") && atoi(getenv("HK_LQNG_MMA4")) == 0) &&
                       !getenv("HK_LQNG_FORCE_GENERIC");
    count_launch(); lqng_assemble_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, stream>>>(batch, N, dt, dx0, dtarget, dtw, dcw, daw, dotgt, dotw,
                                                                              dA, dB, dQ, dq, dR, dx, dn_players, gated ? min_players : 0);
    HK_CUDA(cudaGetLastError());
    return lqng_launch(batch, N, horizon, 0, dA, dB, dQ, dq, dR, dx, du0, nullptr, nullptr, nullptr, dstatus, stream, gated ? dn_players : nullptr,
                       gated ? min_players : 0);
}

// packed record [x0 N x 4 | target N x 4 | tw N x 4 | cw N | aw N x K x 2 | otgt N x K x 4 | otw N x K x 3] -> the seven arrays
__global__ void lqng_unpack_kernel(int batch, int N, const double* __restrict__ rec, double* x0, double* target, double* tw, double* cw, double* aw,
                                   double* otgt, double* otw)
{
    const int K = N - 1, P = 13 * N + 9 * N * K;
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long long)batch * P) return;
    const long long b = id / P;
    int e = (int)(id % P);
    const double v = rec[id];
    if (e < 4 * N) { x0[b * 4 * N + e] = v; return; } e -= 4 * N;
    if (e < 4 * N) { target[b * 4 * N + e] = v; return; } e -= 4 * N;
    if (e < 4 * N) { tw[b * 4 * N + e] = v; return; } e -= 4 * N;
    if (e < N) { cw[b * N + e] = v; return; } e -= N;
    if (e < 2 * N * K) { aw[b * 2 * N * K + e] = v; return; } e -= 2 * N * K;
    if (e < 4 * N * K) { otgt[b * 4 * N * K + e] = v; return; } e -= 4 * N * K;
    otw[b * 3 * N * K + e] = v;
}

int lqng_assemble_launch_packed(int batch, int N, int horizon, double dt, const double* drec, double* du0, int* dstatus, cudaStream_t stream,
                                int scratch_slot, const int* dn_players, const double* dcs_ready)
{
    ThreadCtx* c = ctx();
    if (!c) return HK_ERR_NO_DEVICE;
    if (batch == 0) return HK_OK;
    const int K = N - 1, P = 13 * N + 9 * N * K;
    if (N == 2 && (reinterpret_cast<uintptr_t>(drec) & 15) == 0) {
        // 2-kart game: the solve kernel stages each 352-byte record with ONE bulk copy (+ the (cos h, sin h) pairs) and assembles in shared memory
        // dcs_ready: the caller has the (cos h, sin h) pairs already (the race loop's recipe kernels write them beside the records)
        const double* dcs = dcs_ready;
        if (!dcs) {
            double* w = (double*)dscratch(c, scratch_slot, sizeof(double) * 4 * (size_t)batch);
            if (!w) return HK_ERR_OUT_OF_MEMORY;
            const long long np = (long long)batch * 2;
            count_launch(); lqng_trig_kernel<<<(unsigned)((np + 255) / 256), 256, 0, stream>>>(np, drec, w, 2, P, dn_players);
            HK_CUDA(cudaGetLastError());
            dcs = w;
        }
        LqngParams p{batch, horizon, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, du0, nullptr, nullptr, nullptr, dstatus,
                     nullptr, nullptr, nullptr, nullptr, drec, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, dcs, dt};
        return launch_mma2p(p, stream, true);
    }
    const int n = 4 * N;
    const size_t dense = (size_t)N * 16 + N * 8 + (size_t)N * n * n + (size_t)N * n + N * 4 + n;
    double* d = (double*)dscratch(c, scratch_slot, (dense + P) * sizeof(double) * (size_t)batch);
    if (!d) return HK_ERR_OUT_OF_MEMORY;
    double* u = d + dense * (size_t)batch;                          // the seven arrays behind the dense records lqng_assemble_launch builds in the same slot
    double *ux0 = u, *utg = ux0 + (size_t)batch * 4 * N, *utw = utg + (size_t)batch * 4 * N, *ucw = utw + (size_t)batch * 4 * N,
           *uaw = ucw + (size_t)batch * N, *uot = uaw + (size_t)batch * 2 * N * K, *uow = uot + (size_t)batch * 4 * N * K;
    const long long threads = (long long)batch * P;
    count_launch(); lqng_unpack_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(batch, N, drec, ux0, utg, utw, ucw, uaw, uot, uow);
    HK_CUDA(cudaGetLastError());
    return lqng_assemble_launch(batch, N, horizon, dt, ux0, utg, utw, ucw, uaw, uot, uow, du0, dstatus, stream, scratch_slot, dn_players, 0);
}

// 2-kart packed records solved straight from where they lie (e.g. pinned host memory, device-accessible under UVA): no (cos h, sin h) pre-pass —
// the records are read exactly once, by the solve kernel's bulk copies
int lqng_solve_packed_in_place(int batch, int horizon, double dt, const double* drec, double* du0, int* dstatus, cudaStream_t stream)
{
    if (batch == 0) return HK_OK;
    LqngParams p{batch, horizon, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, du0, nullptr, nullptr, nullptr, dstatus,
                 nullptr, nullptr, nullptr, nullptr, drec, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, dt};
    return launch_mma2p(p, stream, true);
}

int lqng_launch(int batch, int N, int horizon, int time_varying, const double* dA, const double* dB, const double* dQ,
                const double* dq, const double* dR, const double* dx0, double* du0, double* dP, double* dalpha,
                double* dtraj, int* dstatus, cudaStream_t stream, const int* gate, int gate_min)
{
    if (batch == 0) return HK_OK;
    LqngParams p{batch, horizon, time_varying, dA, dB, dQ, dq, dR, dx0, du0, dP, dalpha, dtraj, dstatus, nullptr, nullptr, nullptr, nullptr,
                 nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0};
    p.gate = gate; p.gate_min = gate_min;                           // honoured by lqng_mma4_kernel only (lqng_assemble_launch checks that it takes the launch)
    static const bool force_generic = getenv("HK_LQNG_FORCE_GENERIC") != nullptr;
    static const bool tv2 = !(getenv("HK_MMA2_TV") && atoi(getenv("HK_MMA2_TV")) == 0);
    if (N == 2 && time_varying && horizon <= 7 && !force_generic && tv2 &&
        ((reinterpret_cast<uintptr_t>(dA) | reinterpret_cast<uintptr_t>(dB) | reinterpret_cast<uintptr_t>(dQ) |
          reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dR) | reinterpret_cast<uintptr_t>(dx0)) & 15) == 0) {
        // time-varying operands: the same persistent DMMA kernel with the whole horizon of a problem staged by TMA
        const bool full = dP || dalpha || dtraj;
        return launch_mma2p(p, stream, false, full, true);
    }
    static const bool full2 = !(getenv("HK_MMA2_FULL") && atoi(getenv("HK_MMA2_FULL")) == 0);
    if (N == 2 && !time_varying && (dP || dalpha || dtraj) && !force_generic && full2 &&
        ((reinterpret_cast<uintptr_t>(dA) | reinterpret_cast<uintptr_t>(dB) | reinterpret_cast<uintptr_t>(dQ) |
          reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dR) | reinterpret_cast<uintptr_t>(dx0)) & 15) == 0) {
        // every output (gains, offsets, rollout) from the same persistent DMMA kernel; it re-reads the gains it has written, so it
        // gets scratch for whichever of P / alpha the caller does not keep
        if (!dP || !dalpha) {
            ThreadCtx* c = ctx();
            if (!c) return HK_ERR_NO_DEVICE;
            const size_t T = (size_t)horizon + 1;
            double* scratch = (double*)dscratch(c, 1, sizeof(double) * (size_t)batch * T * 36);
            if (!scratch) return HK_ERR_OUT_OF_MEMORY;
            if (!dP) p.P = scratch;
            if (!dalpha) p.alpha = scratch + (size_t)batch * T * 32;
        }
        return launch_mma2p(p, stream, false, true);
    }
    if (N == 2 && !time_varying && !dP && !dalpha && !dtraj && !force_generic) {
        // throughput path: one launch of a DMMA kernel (problems it cannot take fall back inside the kernel)
        static const int variant = getenv("HK_MMA2_VARIANT") ? atoi(getenv("HK_MMA2_VARIANT")) : 2;   // 0: one CTA per 4 problems; 1, 2: persistent + TMA
        const bool aligned = ((reinterpret_cast<uintptr_t>(dA) | reinterpret_cast<uintptr_t>(dB) | reinterpret_cast<uintptr_t>(dQ) |
                               reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dR) | reinterpret_cast<uintptr_t>(dx0)) & 15) == 0;
        if (variant >= 1 && aligned) return launch_mma2p(p, stream, false);   // cp.async.bulk needs 16-byte aligned sources
        static const int minb = getenv("HK_MMA2_MINB") ? atoi(getenv("HK_MMA2_MINB")) : 5;   // tuning knob, see DESIGN.md §4
        const int wpb = MMA2_THREADS / 32;
        const unsigned grid = (unsigned)((batch + wpb - 1) / wpb);
        count_launch();
        if (minb >= 8) lqng_mma2_kernel<8><<<grid, MMA2_THREADS, 0, stream>>>(p);
        else if (minb >= 6) lqng_mma2_kernel<6><<<grid, MMA2_THREADS, 0, stream>>>(p);
        else if (minb == 5) lqng_mma2_kernel<5><<<grid, MMA2_THREADS, 0, stream>>>(p);
        else lqng_mma2_kernel<4><<<grid, MMA2_THREADS, 0, stream>>>(p);
        HK_CUDA(cudaGetLastError());
        return HK_OK;
    }
    static const bool mma4 = !(getenv("HK_LQNG_MMA4") && atoi(getenv("HK_LQNG_MMA4")) == 0);
    if ((N == 4 || N == 3) && mma4 && !force_generic && (reinterpret_cast<uintptr_t>(dQ) & 15) == 0) {
        // 3- and 4-kart games: warp per problem, DMMA for the 16 x 16 products (hk_lqng_mma4.cuh); takes every operand form
        // (Q_i is read with 128-bit loads: an 8-byte aligned Q goes to the general kernel)
        const size_t smem = (size_t)MMA4_WARPS * Mma4Layout::total * sizeof(double);
        const long long want = ((long long)batch + MMA4_WARPS - 1) / MMA4_WARPS;
        const unsigned grid = (unsigned)(want < 148 * 40 ? want : 148 * 40);
        count_launch();
        if (N == 4) lqng_mma4_kernel<4><<<grid, 32 * MMA4_WARPS, smem, stream>>>(p);
        else lqng_mma4_kernel<3><<<grid, 32 * MMA4_WARPS, smem, stream>>>(p);
        HK_CUDA(cudaGetLastError());
        return HK_OK;
    }
    switch (N) {
        case 1: return launch_generic<1>(p, stream);
        case 2: return launch_generic<2>(p, stream);
        case 3: return launch_generic<3>(p, stream);
        case 4: return launch_generic<4>(p, stream);
    }
    set_error("n_players must be 1..4");
    return HK_ERR_INVALID_ARGUMENT;
}

}  // namespace hk
