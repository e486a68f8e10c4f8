// hk_lqng_mma.cuh — warp-per-problem FP64 tensor-core (DMMA m8n8k4) kernel for the time-invariant 2-kart LQNG
// (n = 8, m = 4), the BASELINE headline configuration.  Same algorithm as KartLQR.solveFeedbackLQR
// (reference: Assets/Karting/Scripts/AI/LQR/KartLQR.cs:64-127), re-associated so that every 8x8 product is a pair of
// mma.sync.m8n8k4.f64 and no matrix ever leaves the register file.
//
// Why DMMA (measured on B200, profiles/fp64_microbench_r01.md): DFMA and DMMA share the FP64 pipe and reach the same
// flop rate (33.5 vs 37.0 TFLOP/s sustained), but one DMMA replaces 8 DFMA warp-instructions AND does the operand
// broadcast that a lane-per-row DFMA kernel pays for with shuffles (1.28 SHFL/clk/SM) or shared-memory loads.
//
// Register forms of an 8x8 matrix M, lane L = 4g + t (g = L>>2 in 0..7, t = L&3):
//   R-form (m0,m1) = (M[g][2t],  M[g][2t+1])   = the DMMA C/D fragment; directly usable as an A operand whose four k-slots
//                                                  are k = 2t (first DMMA) and k = 2t+1 (second DMMA);
//   T-form (m0,m1) = (M[2t][g],  M[2t+1][g])   = R-form of M^T; directly usable as a B operand with the same k permutation.
// so  D(R-form) = X(R-form) * Y(T-form)  costs two DMMAs and zero data movement, and the R-form registers of M double as
// the T-form of M^T.  The recursion is arranged so that every product finds its operands already in the right form:
//   Y_i^T = F^T Z_i^T            X = F T-form (= F^T R-form),   Y = Z_i R-form (= Z_i^T T-form)   -> Y_i in T-form
//   Z_i  <- Q_i + P_i^T(R_i P_i) + F^T Y_i                       (accumulated in the C fragment)
//   W     = sum_i [B_i^T ; beta^T] Z_i      rows 0..3 = stacked B_i^T Z_i, row 4+i = (Z_i beta)^T  (needs Z_i symmetric)
//   L     = W B,  RM^T = A^T W^T,  P = LHS^-1 RM (as RM^T Lambda^T),  alpha = LHS^-1 rv,  u^T = (eta + Z beta)^T F
// Vectors (eta_i, q_i, beta, Z_i beta) ride in rows 4 and 5 of these products, lane (4+i, t) holding entries (2t, 2t+1).
// The 4x4 coupled system is inverted in place by Gauss-Jordan over 8 lanes; the kernel verifies at every pivot that
// partial pivoting (MathNet LU, KartLQR.cs:104-105) would not have exchanged rows — otherwise, or if Q_i / R_i are not
// symmetric, the warp solves its problem with the generic pivoting algorithm instead (lqng_generic_body).
#pragma once

namespace hk {

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// D += X(R-form) * Y(T-form)
__device__ __forceinline__ void mm(double& c0, double& c1, double x0, double x1, double y0, double y1)
{
    dmma(c0, c1, x0, y0);
    dmma(c0, c1, x1, y1);
}
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
// 1/x to ~2^-60 relative: MUFU.RCP64H seed (20 bits) + one cubic step, 3 DFMA instead of the 5 DFMA + slow-path call of an
// IEEE division.  Callers route denormal / huge pivots to the pivoting kernel, so no special cases are needed here.
__device__ __forceinline__ double rcp_fast(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
}
// |a| > |b| and |a| outside the normal range, on the integer pipes (DSETP/DADD would occupy the FP64 pipe)
__device__ __forceinline__ bool abs_gt(double a, double b)
{
    const unsigned ha = (unsigned)__double2hiint(a) & 0x7fffffffu, hb = (unsigned)__double2hiint(b) & 0x7fffffffu;
    return (ha > hb) | ((ha == hb) & ((unsigned)__double2loint(a) > (unsigned)__double2loint(b)));   // no short-circuit: branch-free
}
__device__ __forceinline__ bool bad_pivot(double a)                 // zero, denormal, tiny, huge, inf or NaN
{
    const unsigned ex = ((unsigned)__double2hiint(a) >> 20) & 0x7ffu;
    return (ex - 0x100u) > (0x6ffu - 0x100u);                      // one unsigned compare
}

constexpr int MMA2_THREADS = 128;
constexpr int MMA2_PREFETCH_DISTANCE = 4096;                       // problems; > resident warps per GPU (148 SMs x 20)

template <int MINB>                                             // resident CTAs per SM the register allocation is sized for
__global__ void __launch_bounds__(MMA2_THREADS, MINB) lqng_mma2_kernel(LqngParams p)
{
    __shared__ double fallback_smem[MMA2_THREADS / 32][GenericLayout<2>::total];   // only touched by warps that fall back
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const long long prob = (long long)blockIdx.x * (MMA2_THREADS / 32) + (threadIdx.x >> 5);
    if (prob >= p.batch) return;                                   // whole warps only: no divergence inside a warp
    const double* gA = p.A + (size_t)prob * 32;
    const double* gB = p.B + (size_t)prob * 16;
    const double* gQ = p.Q + (size_t)prob * 128;
    const double* gq = p.q + (size_t)prob * 16;
    const double* gR = p.R + (size_t)prob * 8;
    const double* gx = p.x0 + (size_t)prob * 8;
    {   // L2 prefetch of the record a later warp will need (14 lines of 128 B): turns its HBM latency into L2 latency
        const long long pf = prob + MMA2_PREFETCH_DISTANCE;
        if (pf < p.batch && lane < 14) {
            const double* a = lane < 8 ? p.Q + (size_t)pf * 128 + lane * 16
                            : lane < 10 ? p.A + (size_t)pf * 32 + (lane - 8) * 16
                            : lane == 10 ? p.B + (size_t)pf * 16
                            : lane == 11 ? p.q + (size_t)pf * 16
                            : lane == 12 ? p.R + (size_t)pf * 8 : p.x0 + (size_t)pf * 8;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
        }
    }

    // ---- per-lane constants --------------------------------------------------------------------------------------
    const int pl = t >> 1;                                          // player owning joint rows 2t, 2t+1 (and control row t)
    const int lr = (2 * t) & 3;                                     // local row of joint row 2t inside A_pl / B_pl
    // joint A in T-form: (A[2t][g], A[2t+1][g]); block diagonal (KartLQR.cs:33-37)
    const bool a_on = pl == (g >> 2);
    const double aT0 = a_on ? gA[pl * 16 + lr * 4 + (g & 3)] : 0.0;
    const double aT1 = a_on ? gA[pl * 16 + (lr + 1) * 4 + (g & 3)] : 0.0;
    // joint B (8x4, column c belongs to player c>>1; KartLQR.cs:41-52) in T-form, padded to 8 columns
    const bool b_on = g < 4 && pl == (g >> 1);
    const double bT0 = b_on ? gB[pl * 8 + lr * 2 + (g & 1)] : 0.0;
    const double bT1 = b_on ? gB[pl * 8 + (lr + 1) * 2 + (g & 1)] : 0.0;
    // B rows 2t, 2t+1 against their own player's two controls (F = A - B P, beta = -B alpha; KartLQR.cs:110-111)
    const double bF00 = gB[pl * 8 + lr * 2 + 0], bF01 = gB[pl * 8 + lr * 2 + 1];
    const double bF10 = gB[pl * 8 + (lr + 1) * 2 + 0], bF11 = gB[pl * 8 + (lr + 1) * 2 + 1];
    // R of player t>>1, row t&1 (for R P and R alpha), and the diagonal-block entries of the coupled LHS
    const double rr0 = gR[pl * 4 + (t & 1) * 2 + 0], rr1 = gR[pl * 4 + (t & 1) * 2 + 1];
    const bool lhs_lane = g < 4 && t < 2;                           // natural LHS distribution, see below
    const bool l_diag = lhs_lane && t == (g >> 1);
    const double rl0 = l_diag ? gR[(g >> 1) * 4 + (g & 1) * 2 + 0] : 0.0;
    const double rl1 = l_diag ? gR[(g >> 1) * 4 + (g & 1) * 2 + 1] : 0.0;
    // (A x0)[2t], (A x0)[2t+1]: rides in column 4 of the L product so that RHSMat x0 lands in the augmented lanes at t = 0
    double ax0 = 0.0, ax1 = 0.0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const double xc = gx[4 * pl + c];
        ax0 = fma(gA[pl * 16 + lr * 4 + c], xc, ax0);
        ax1 = fma(gA[pl * 16 + (lr + 1) * 4 + c], xc, ax1);
    }
    const double yL0 = g == 4 ? ax0 : bT0, yL1 = g == 4 ? ax1 : bT1;     // [B | A x0 | 0 0 0] in T-form
    const bool vec_lane = (g >> 1) == 2;                            // quads 4 and 5 carry the vectors of player g-4
    const int vp = g & 1;
    double2 qv = make_double2(0.0, 0.0);
    if (vec_lane) qv = *reinterpret_cast<const double2*>(gq + vp * 8 + 2 * t);
    // Q_i in R-form (coalesced 128-bit loads: lane L reads doubles 2L, 2L+1 of the 8x8 block)
    const double2 q0 = *reinterpret_cast<const double2*>(gQ + 2 * lane);
    const double2 q1 = *reinterpret_cast<const double2*>(gQ + 64 + 2 * lane);
    // symmetry of Q_i and R_i is what lets Z_i's R-form stand in for its T-form in the W product
    bool redo = (gQ[(2 * t) * 8 + g] != q0.x) | (gQ[(2 * t + 1) * 8 + g] != q0.y) | (gQ[64 + (2 * t) * 8 + g] != q1.x) |
                (gQ[64 + (2 * t + 1) * 8 + g] != q1.y) | (gR[1] != gR[2]) | (gR[5] != gR[6]);

    // shuffle sources of the Gauss-Jordan inversion. Natural LHS distribution: lane (g<4, t<2) holds
    // LHS[rho][2*kap], LHS[rho][2*kap+1] with rho = 2t + (g&1), kap = g>>1 — exactly where the D fragment of L = W B leaves
    // them once the reference's block placement is applied (block (i,j) of B_i^T Z_i B_j goes to row-block j,
    // column-block i: quirk Q1, KartLQR.cs:68-87).
    const bool aug_lane = g < 4 && t == 2;
    const int rho = aug_lane ? g : 2 * t + (g & 1), kap = g >> 1;
    const int rowhi = aug_lane ? (g >> 1) : t;                      // rho >> 1
    const bool below[4] = {lhs_lane && rho > 0, lhs_lane && rho > 1, lhs_lane && rho > 2, false};   // candidate rows of pivot k
    // Lambda rows re-dealt as the B operand of the P product: column map s(col) = (0,1,0,1,2,3,2,3)
    const int s_of_g = ((g >> 2) << 1) | (g & 1);
    const int srcLam = 4 * ((s_of_g & 1) + 2 * (t & 1)) + (s_of_g >> 1);
    const int srcRv = 4 * (4 + ((g >> 1) & 1)) + ((g >> 1) & 1);     // rv of player g>>1 sits in lane (4 + (g>>1), t = g>>1)

    // ---- state ------------------------------------------------------------------------------------------------------
    double z00 = q0.x, z01 = q0.y, z10 = q1.x, z11 = q1.y;         // Z_i = Q_i (KartLQR.cs:62)
    double e0 = qv.x, e1 = qv.y;                                    // eta_i = q_i (:63), lanes (4+i, t)
    double w0 = 0.0, w1 = 0.0;                                      // W (rows 0..3) and Z_i beta (rows 4, 5)
    // A operands of the W product: rows 2i, 2i+1 = B_i^T (loop invariant), row 4+i = beta^T (filled in per step)
    const double xb00 = g < 2 ? bT0 : 0.0, xb01 = g < 2 ? bT1 : 0.0;
    const double xb10 = (g >> 1) == 1 ? bT0 : 0.0, xb11 = (g >> 1) == 1 ? bT1 : 0.0;
    mm(w0, w1, xb00, xb01, z00, z01);
    mm(w0, w1, xb10, xb11, z10, z11);
    double pe = 0.0, po = 0.0, ae = 0.0, ao = 0.0, u_out = 0.0;

    for (int step = p.horizon; step >= 0; --step) {                 // KartLQR.cs:64
        // L = W B with eta^T B in rows 4, 5 (RHSVec, :96)
        double l0 = rl0, l1 = rl1;                                  // + R_i on the diagonal blocks of the LHS (:78)
        {
            const double x0 = g < 4 ? w0 : (vec_lane ? e0 : 0.0), x1 = g < 4 ? w1 : (vec_lane ? e1 : 0.0);
            mm(l0, l1, x0, x1, yL0, yL1);
        }
        // RM^T = A^T W^T (RHSMat, :89-95); not needed at t = 0, where only u0 = -(P x0 + alpha) = -LHS^-1 (RM x0 + rv) is
        double m0 = 0.0, m1 = 0.0;
        if (step != 0) mm(m0, m1, aT0, aT1, g < 4 ? w0 : 0.0, g < 4 ? w1 : 0.0);
        // coupled LHS in the natural distribution, inverted in place by Gauss-Jordan;
        // lanes (g<4, t=2) carry RHSVec as an augmented column (row g), so alpha = LHS^-1 rv falls out of the same sweep
        double M0 = l0, M1 = l1;
        {
            const double v0 = shfl_d(l0, srcRv), v1 = shfl_d(l1, srcRv);
            if (aug_lane) { M0 = ((g & 1) ? v1 : v0) + (step == 0 ? l0 : 0.0); M1 = 0.0; }    // l0 here = (RM x0)[g]
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int kk = k >> 1, bk = k & 1;
            const double mine = bk ? M1 : M0;
            const double c = shfl_d(mine, 4 * ((g & 1) + 2 * kk) + rowhi);                // LHS[rho][k] of this lane's row
            const int srcR = aug_lane ? 4 * k + 2 : 4 * ((k & 1) + 2 * (g >> 1)) + (k >> 1);
            const double r0 = shfl_d(M0, srcR), r1 = shfl_d(M1, srcR);                    // pivot row k, this lane's column(s)
            const double pv = shfl_d(mine, 4 * ((k & 1) + 2 * kk) + (k >> 1));            // LHS[k][k], straight from its owner
            redo |= (below[k] & abs_gt(c, pv)) | bad_pivot(pv);                           // partial pivoting would swap rows
            const double pinv = rcp_fast(pv);
            const bool prow = rho == k;
            const double coef = prow ? pinv : -(c * pinv);                                // also column k of the in-place inverse
            const double n0 = fma(coef, r0, prow ? 0.0 : M0);
            const double n1 = fma(coef, r1, prow ? 0.0 : M1);
            M0 = (lhs_lane && kap == kk && bk == 0) ? coef : n0;
            M1 = (lhs_lane && kap == kk && bk == 1) ? coef : n1;
        }
        if (step == 0) {
            // optimal_control = -P x0 - alpha with the t = 0 gains (:121-126); the last update of Z, eta is never used
            u_out = -M0;
            break;
        }
        // alpha of player t>>1 from the augmented lanes (row r lives in lane 4r + 2)
        ae = shfl_d(M0, 8 * (t >> 1) + 2);
        ao = shfl_d(M0, 8 * (t >> 1) + 6);
        // Lambda = LHS^-1 as B operand: Y[k][col] = Lambda[s(col)][k]
        const bool ylane = t < 2;
        double y0 = shfl_d(M0, srcLam), y1 = shfl_d(M1, srcLam);
        y0 = ylane ? y0 : 0.0; y1 = ylane ? y1 : 0.0;
        // P (both rows of player t>>1 for column g): (pe, po) = (P[2(t>>1)][g], P[2(t>>1)+1][g])   (:104-105)
        pe = 0.0; po = 0.0;
        mm(pe, po, m0, m1, y0, y1);
        const double pc = (t & 1) ? po : pe;                        // compact P: P[t][g]

        // F = A - B P (T-form), beta = -B alpha (rows 2t, 2t+1 at lanes (4+i, t))   (:110-111)
        const double f0 = fma(-bF01, po, fma(-bF00, pe, aT0));
        const double f1 = fma(-bF11, po, fma(-bF10, pe, aT1));
        const double be0 = fma(-bF01, ao, -bF00 * ae);             // only read in quads 4, 5
        const double be1 = fma(-bF11, ao, -bF10 * ae);
        const double rpc = fma(rr1, po, rr0 * pe);                  // (R_p P_p)[t&1][g]
        // Z_i <- Q_i + P_i^T R_i P_i + F^T (Z_i F)   (:116)
        {
            double yt0 = 0.0, yt1 = 0.0;
            mm(yt0, yt1, f0, f1, z00, z01);
            z00 = q0.x; z01 = q0.y;
            dmma(z00, z01, pl == 0 ? pc : 0.0, rpc);
            mm(z00, z01, f0, f1, yt0, yt1);
        }
        {
            double yt0 = 0.0, yt1 = 0.0;
            mm(yt0, yt1, f0, f1, z10, z11);
            z10 = q1.x; z11 = q1.y;
            dmma(z10, z11, pl == 1 ? pc : 0.0, rpc);
            mm(z10, z11, f0, f1, yt0, yt1);
        }
        // W for the next step, with beta^T in row 4+i so that row 4+i of the result is (Z_i^{new} beta)^T (quirk Q2);
        // rows 4, 5 start from eta_i, so they come out as eta_i + Z_i^{new} beta, the vector F^T is applied to below
        w0 = vec_lane ? e0 : 0.0; w1 = vec_lane ? e1 : 0.0;
        mm(w0, w1, g == 4 ? be0 : xb00, g == 4 ? be1 : xb01, z00, z01);
        mm(w0, w1, g == 5 ? be0 : xb10, g == 5 ? be1 : xb11, z10, z11);
        // eta_i <- q_i + P_i^T R_i alpha_i + F^T (eta_i + Z_i^{new} beta)   (:117)
        {
            const double ra = fma(rr1, ao, rr0 * ae);               // (R_i alpha_i)[t&1] at lanes (4+i, 2i + (t&1))
            double n0 = qv.x, n1 = qv.y;
            dmma(n0, n1, (vec_lane && pl == vp) ? ra : 0.0, pc);
            mm(n0, n1, vec_lane ? w0 : 0.0, vec_lane ? w1 : 0.0, f0, f1);
            e0 = n0; e1 = n1;
        }
    }
    const unsigned any_redo = __ballot_sync(0xffffffffu, redo);
    if (any_redo) {
        // Warp-uniform and rare: this problem needs row exchanges (or has non-symmetric Q/R, or a degenerate pivot).
        // Solve it right here with the pivoting shared-memory algorithm on 8 lanes of this warp.
        if (lane < 8) lqng_generic_body<2>(p, prob, true, fallback_smem[threadIdx.x >> 5], lane, 0xffu);
        return;
    }
    if (aug_lane) p.u0[(size_t)prob * 4 + g] = u_out;
    if (lane == 0 && p.status) p.status[prob] = 0;
}

}  // namespace hk
