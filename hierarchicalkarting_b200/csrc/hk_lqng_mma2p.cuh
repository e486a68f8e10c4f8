// hk_lqng_mma2p.cuh — persistent, TMA-staged version of the warp-per-problem FP64 tensor-core kernel for the
// time-invariant 2-kart LQNG (n = 8, m = 4): the BASELINE headline configuration.
// Same algorithm as KartLQR.solveFeedbackLQR (reference: Assets/Karting/Scripts/AI/LQR/KartLQR.cs:64-127) and the same
// DMMA fragment algebra as hk_lqng_mma.cuh (read its header first).  What changes:
//   * persistent warps: grid = SMs x resident CTAs, every warp strides over the batch, so the ~25 per-lane layout
//     constants (fragment ownership, shuffle sources, operand offsets) are computed once per warp, not once per problem;
//   * TMA staging: one elected lane issues six cp.async.bulk copies (Q 1024 B, A 256 B, B 128 B, q 128 B, R 64 B,
//     x0 64 B = the problem's whole 1,664 B record) into a per-warp, double-buffered shared-memory record that completes
//     on an mbarrier, one problem ahead of the one being solved: HBM latency is off the critical path and every operand
//     read is a conflict-free LDS.  Zero-padding of fragment operands is done by pointing a lane at a constant tail
//     ([0 0 0 0 0 1 1 0]) instead of select instructions; Q_i / q_i are re-read from the record each step instead of
//     being pinned in registers;
//   * the coupled 4x4 system [LHS | I | RHSVec] is reduced by fraction-free Gauss-Jordan (rows are rescaled by the pivot
//     instead of divided by it): the identity block lives in the otherwise unused rows 6,7 of the L fragment, there is
//     no placement select and no reciprocal on the pivot chain — one reciprocal per row, all rows in parallel, at the
//     end.  All rows carry the same cumulative scale, so the test "partial pivoting (MathNet LU, KartLQR.cs:104-105)
//     would not exchange rows" stays a plain magnitude comparison, done on the high words in the integer pipe.
// DRAM writes (ncu: 18.7 MB per 65,536-problem launch against 2.4 MB of outputs): 15.7 MB of it is the kernel's own parameter block —
// lqng_generic_body takes it by reference, so the compiler keeps the 208-byte LqngParams on every thread's local stack (26 STL.64 at entry,
// 2,368 warps x 32 lanes x 208 B) and reads p.* from there.  Measured alternatives (round 2): `const __grid_constant__` parameter, or a copy
// made only in the rare path — both bring the writes down to 4.9 MB and both are SLOWER (4.34e8 vs 4.52e8 solves/s): the parameter fields
// then live in registers and the 128-register budget spills inside the recursion.  The stack copy is a once-per-thread cost and stays.
// A problem whose coupled system needs row exchanges is solved a second time by the same warp with `pivot` set: the 4x4
// system [LHS | I | RHSVec] is then dealt one column per lane and reduced by Gauss-Jordan with partial pivoting (same
// pivot choice as MathNet's LU: first largest magnitude at or below the diagonal), everything else (all DMMA products)
// is the same code.  That second pass costs one more solve (~7 us) instead of the ~40 us of the shared-memory
// algorithm, which matters because a persistent launch ends with its slowest warp.  Only problems with non-symmetric
// Q_i/R_i or a zero / out-of-range pivot go to lqng_generic_body (same launch).
#pragma once

namespace hk {

constexpr int P2_oQ = 0, P2_oA = 128, P2_oB = 160, P2_oq = 176, P2_oR = 192, P2_ox = 200, P2_oC = 208;   // doubles
constexpr int P2_STRIDE = 216;                  // record + constant tail, 1,728 B
constexpr unsigned P2_TX_BYTES = 1664;

// compact (fused assembly) mode: the problem arrives as the constructor arguments of the reference's providers
// (include/hk_abi.h, hk_lqng_assemble_solve_batch) plus (cos h, sin h) per player; staging offsets in doubles
constexpr int C2_ox = 0, C2_otg = 8, C2_otw = 16, C2_ocw = 24, C2_oaw = 26, C2_oot = 30, C2_oow = 38, C2_ocs = 44, C2_STRIDE = 48;
constexpr unsigned C2_TX_BYTES = 384;

constexpr int P2_ROLL_STEPS = 8;                // FULL mode keeps the gains of up to 8 steps in shared memory for its rollout

template <int WARPS, int ROLL, int RECS>
struct P2Smem {
    double rec[WARPS][2][RECS];                  // TV mode keeps its records (one per stage) in dynamic shared memory instead
    double stage[WARPS][2][C2_STRIDE];           // compact mode only
    double roll[WARPS][ROLL];                    // FULL mode: state (8), controls (4) of the rollout; P_t (32) + alpha_t (4) per step
    double fallback[GenericLayout<2>::total];    // one pivoting scratch per CTA, serialised by `lock` (rare path)
    unsigned long long bar[WARPS][2];
    int lock;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tHK_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra HK_DONE;\n\tbra HK_WAIT;\n\tHK_DONE:\n\t}"
                 ::"r"(bar), "r"(parity) : "memory");
}
// out of line on purpose: the slow paths of sin / cos must not cost the solve loop registers
__device__ __noinline__ double p2_trig(double h, int want_sin) { return want_sin ? sin(h) : cos(h); }
__device__ __forceinline__ unsigned abs_hi(double a) { return (unsigned)__double2hiint(a) & 0x7fffffffu; }
__device__ __forceinline__ bool bits_differ(double a, double b)
{
    return ((__double2hiint(a) ^ __double2hiint(b)) | (__double2loint(a) ^ __double2loint(b))) != 0;
}

// WARPS = warps per CTA.  WARPS == 1 makes every address of the TMA issue path a function of blockIdx.x only, i.e.
// provably warp-uniform: the copies are issued from uniform registers without the elect/broadcast loops the compiler
// otherwise wraps around each cp.async.bulk, and the loop state lives in uniform registers.
// COMPACT: the warp assembles its dense record in shared memory from the staged compact description — LinearizedBicycle.getA/getB
// (KartLQRDynamics.cs:40-62) and LQRCheckpointReachAvoidCost.getQMatrix/getQVec/getRMatrix (KartLQRCosts.cs:57-140, quirks Q3-Q5 of
// SURVEY.md A.3), same arithmetic as lqng_assemble_kernel — instead of reading a dense record another kernel wrote to HBM.
// FULL: every output of the ABI — gains P_t, offsets alpha_t of every step (the reference computes them and keeps only t = 0,
// KartLQR.cs:104-105, 121-126), u0 = -P_0 x0 - alpha_0 of every player, and the closed-loop rollout.  The last
// backward step is then a full step (no u0 shortcut), gains go to global memory as they are produced and the rollout re-reads them.
// TV: time-varying operands A_t, B_t, Q_t, q_t, R_t (the reference's providers are constant over the horizon).  The
// whole horizon of a problem is staged by TMA — stage t as one record of the time-invariant layout at rec + t * P2_STRIDE, x0 with
// stage 0 — double-buffered across problems in dynamic shared memory; every per-lane offset then applies to the stage's record.
template <int MINB, int WARPS, bool COMPACT = false, bool FULL = false, bool TV = false>
__global__ void __launch_bounds__(32 * WARPS, MINB) lqng_mma2p_kernel(LqngParams p)
{
    static_assert(!TV || (WARPS == 1 && !COMPACT), "TV: one warp per CTA, dense records");
    __shared__ __align__(128) P2Smem<WARPS, FULL ? 16 + 36 * P2_ROLL_STEPS : 1, TV ? 2 : P2_STRIDE> sm;
    extern __shared__ __align__(128) double tv_rec[];               // TV: [2][T][P2_STRIDE]
    const int TS = TV ? p.horizon + 1 : 1;                          // stages per record
    const int lane = threadIdx.x & 31, wib = WARPS == 1 ? 0 : threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const long long nwarps = (long long)gridDim.x * WARPS;
    const long long first = (long long)wib * gridDim.x + blockIdx.x;      // leftovers of the last round spread over all SMs
    const unsigned bar_u32 = smem_u32(&sm.bar[wib][0]);
    double* const rec_base = TV ? tv_rec : &sm.rec[wib][0][0];
    const unsigned rec_u32 = smem_u32(rec_base);
    if (threadIdx.x == 0) sm.lock = 0;
    for (int i = lane; i < 16 * TS; i += 32) rec_base[(i >> 3) * P2_STRIDE + P2_oC + (i & 7)] = ((i & 7) == 5 || (i & 7) == 6) ? 1.0 : 0.0;
    if (lane == 0) {
        mbar_init(bar_u32, 1);
        mbar_init(bar_u32 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // Work distribution: the first problem of a warp is static, the following ones come from a device-wide counter when the
    // launcher provides one (p.work) — SMs differ by a few per cent in how fast they get through their problems, and a
    // persistent launch ends with its slowest SM.  p.work[0] = next index - gridDim.x * WARPS, p.work[1] = warps that are done;
    // the last warp to finish clears both for the next launch (every fetch of a warp precedes its own done-increment).
    const bool dynamic = p.work != nullptr;
    auto fetch = [&]() -> long long {                                // all lanes; returns the next problem of this warp
        int v = 0;
        if (lane == 0) v = atomicAdd(p.work, 1);
        return nwarps + (long long)__shfl_sync(0xffffffffu, v, 0);
    };
    auto finish = [&]() {
        if (dynamic && lane == 0 && atomicAdd(p.work + 1, 1) == (int)nwarps - 1) { p.work[0] = 0; p.work[1] = 0; }
    };
    // Programmatic dependent launch: a following launch of this kernel on the same stream may start filling SMs as this one's
    // warps exit (its ramp-up overlaps this launch's tail).  Such a dependent reads only its own inputs before
    // griddepcontrol.wait, which it executes before its first global store (see below) — by then this grid has completed.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    bool waited = false;                                            // FULL: gains are stored from inside the recursion — wait before the first one
    if (first >= p.batch) { asm volatile("griddepcontrol.wait;" ::: "memory"); finish(); return; }   // whole warps only; no block-wide sync below

    const unsigned stage_u32 = smem_u32(&sm.stage[wib][0][0]);
    auto issue = [&](long long prob, int b) {                         // lane 0 only
        const unsigned bar = bar_u32 + 8u * b, dst = rec_u32 + (unsigned)(b * TS * P2_STRIDE * 8);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic-proxy reads of this buffer come first
        if (COMPACT) {
            const unsigned sd = stage_u32 + (unsigned)(b * C2_STRIDE * 8);
            if (!p.c_target) {                                       // packed records (hk_lqng_assemble_solve_packed): the staged layout IS the record
                if (!p.c_cs) {                                       // no (cos h, sin h) array: the warp evaluates them itself (records read once, e.g.
                    mbar_expect_tx(bar, C2_ocs * 8);                 // straight from pinned host memory)
                    bulk_g2s(sd, p.c_x0 + (size_t)prob * C2_ocs, C2_ocs * 8, bar);
                    return;
                }
                mbar_expect_tx(bar, C2_TX_BYTES);
                bulk_g2s(sd, p.c_x0 + (size_t)prob * C2_ocs, C2_ocs * 8, bar);           // 352 B: x0 | target | tw | cw | aw | otgt | otw
                bulk_g2s(sd + C2_ocs * 8, p.c_cs + (size_t)prob * 4, 32, bar);
                return;
            }
            mbar_expect_tx(bar, C2_TX_BYTES);
            bulk_g2s(sd + C2_ox * 8, p.c_x0 + (size_t)prob * 8, 64, bar);
            bulk_g2s(sd + C2_otg * 8, p.c_target + (size_t)prob * 8, 64, bar);
            bulk_g2s(sd + C2_otw * 8, p.c_tw + (size_t)prob * 8, 64, bar);
            bulk_g2s(sd + C2_ocw * 8, p.c_cw + (size_t)prob * 2, 16, bar);
            bulk_g2s(sd + C2_oaw * 8, p.c_aw + (size_t)prob * 4, 32, bar);
            bulk_g2s(sd + C2_oot * 8, p.c_otgt + (size_t)prob * 8, 64, bar);
            bulk_g2s(sd + C2_oow * 8, p.c_otw + (size_t)prob * 6, 48, bar);
            bulk_g2s(sd + C2_ocs * 8, p.c_cs + (size_t)prob * 4, 32, bar);
            return;
        }
        if (TV) {
            mbar_expect_tx(bar, (unsigned)TS * 1600u + 64u);
            for (int st = 0; st < TS; ++st) {
                const unsigned d = dst + (unsigned)(st * P2_STRIDE * 8);
                const size_t ps = (size_t)prob * TS + st;
                bulk_g2s(d + P2_oQ * 8, p.Q + ps * 128, 1024, bar);
                bulk_g2s(d + P2_oA * 8, p.A + ps * 32, 256, bar);
                bulk_g2s(d + P2_oB * 8, p.B + ps * 16, 128, bar);
                bulk_g2s(d + P2_oq * 8, p.q + ps * 16, 128, bar);
                bulk_g2s(d + P2_oR * 8, p.R + ps * 8, 64, bar);
            }
            bulk_g2s(dst + P2_ox * 8, p.x0 + (size_t)prob * 8, 64, bar);
            return;
        }
        mbar_expect_tx(bar, P2_TX_BYTES);
        bulk_g2s(dst + P2_oQ * 8, p.Q + (size_t)prob * 128, 1024, bar);
        bulk_g2s(dst + P2_oA * 8, p.A + (size_t)prob * 32, 256, bar);
        bulk_g2s(dst + P2_oB * 8, p.B + (size_t)prob * 16, 128, bar);
        bulk_g2s(dst + P2_oq * 8, p.q + (size_t)prob * 16, 128, bar);
        bulk_g2s(dst + P2_oR * 8, p.R + (size_t)prob * 8, 64, bar);
        bulk_g2s(dst + P2_ox * 8, p.x0 + (size_t)prob * 8, 64, bar);
    };
    if (lane == 0) issue(first, 0);

    // ---- per-lane layout constants (once per warp) ------------------------------------------------------------------------
    const int pl = t >> 1;                                          // player owning joint rows 2t, 2t+1 (and control row t)
    const int lr = (2 * t) & 3;                                     // local row of joint row 2t inside A_pl / B_pl
    const bool a_on = pl == (g >> 2);
    const int offA = a_on ? P2_oA + pl * 16 + lr * 4 + (g & 3) : P2_oC;          // joint A, T-form: A[2t][g]; A[2t+1][g] at +4
    const bool b_on = g < 4 && pl == (g >> 1);
    const int offB = b_on ? P2_oB + pl * 8 + lr * 2 + (g & 1) : P2_oC;           // joint B (8x4), T-form; next row at +2
    const int offXb0 = (b_on && g < 2) ? offB : P2_oC;                           // rows 0,1 of the W operand: B_0^T
    const int offXb1 = (b_on && g >= 2) ? offB : P2_oC;                          // rows 2,3: B_1^T
    const int offBF = P2_oB + pl * 8 + lr * 2;                                   // B rows 2t, 2t+1 x own player's controls
    const int offAx = P2_oA + pl * 16 + lr * 4;
    const int offRr = P2_oR + pl * 4 + (t & 1) * 2;                              // R_pl row t&1
    const bool isL = g < 4 && t < 2, isI = g >= 6, isAug = g < 4 && t == 2;
    // coupled system, lane ownership: L lanes hold LHS[rho][2kap..2kap+1] exactly where the D fragment of L = W B leaves
    // them under the reference's block placement (quirk Q1, KartLQR.cs:68-87); I lanes (rows 6,7 of the same fragment)
    // hold the identity block; Aug lanes hold RHSVec[rho].
    const int rho = isL ? 2 * t + (g & 1) : isI ? 2 * (t & 1) + (g & 1) : isAug ? g : 0;
    const int kap = isL ? (g >> 1) : isI ? (t >> 1) : 0;
    const int offRl = (isL && t == (g >> 1)) ? P2_oR + (g >> 1) * 4 + (g & 1) * 2
                      : isI ? (rho == 2 * kap ? P2_oC + 6 : rho == 2 * kap + 1 ? P2_oC + 4 : P2_oC) : P2_oC;
    const bool vec = (g >> 1) == 2;                                 // quads 4 and 5 carry the vectors of player g-4
    const int vp = g & 1;
    const int offQv = vec ? P2_oq + vp * 8 + 2 * t : P2_oC;
    const int rho_chk = isL ? rho : -1;
    const int csrc = 4 * (rho & 1) + (rho >> 1);                    // + 8 (k>>1): L lane holding LHS[rho][k]
    const int rb = isL ? 8 * kap : isI ? 24 + 2 * kap : 2;          // row k, this lane's columns: rb + 4 (k&1) + (k>>1) [+ 7 (k>>1) on Aug lanes]
    const int rb2 = rb + 1 + (isAug ? 7 : 0);
    const int dsrc = csrc + 8 * (rho >> 1);                         // L lane holding the diagonal of row rho
    const int srcRv = 4 * (4 + ((g >> 1) & 1)) + ((g >> 1) & 1);    // RHSVec of player g>>1 sits in lane (4 + (g>>1), t = g>>1)
    const int s_of_g = ((g >> 2) << 1) | (g & 1);
    const int srcLam = 4 * (6 + (s_of_g & 1)) + (s_of_g >> 1) + 2 * (t & 1);   // I lane holding Lambda[s(g)][2t..2t+1]
    const int srcAe = 8 * pl + 2, srcAo = 8 * pl + 6;
    const bool ylane = t < 2, podd = t & 1, godd = g & 1, rodd = rho & 1;

    int it = 0;
    long long nxt = dynamic ? fetch() : first + nwarps;
    for (long long prob = first; prob < p.batch; ++it) {
        const int buf = it & 1;
        const double* rec = TV ? rec_base + (size_t)buf * TS * P2_STRIDE : &sm.rec[wib][COMPACT ? 0 : buf][0];   // TV: stage t at + t * P2_STRIDE
        __syncwarp();                                               // every lane is done with the other buffer
        if (lane == 0 && nxt < p.batch) issue(nxt, buf ^ 1);
        // The problem after the next: fetched now, needed at the end.  Inline PTX on purpose: nvcc turns a predicated atomicAdd() into
        // a warp-aggregated one (leader election + SHFL of the result), and that SHFL waits for the atomic's round trip right here
        // (4.6 % of the kernel's stall samples, ncu r01e) instead of ~5 us later where the value is consumed.
        int after = 0;
        if (dynamic && nxt < p.batch)                               // warp-uniform; the elected lane is lane 0
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\t@p atom.global.add.u32 %0, [%1], 1;\n\t}" : "+r"(after) : "l"(p.work) : "memory");
        mbar_wait(bar_u32 + 8u * buf, (unsigned)(it >> 1) & 1u);
        if (COMPACT) {
            // assemble the dense record (layout P2_o*) from the staged description; every lane writes a few entries
            double* r = &sm.rec[wib][0][0];
            double* c = &sm.stage[wib][buf][0];
            if (!p.c_cs) {                                           // warp-uniform: cos / sin of the two headings by lanes 0..3 (the same
                if (lane < 4) c[C2_ocs + lane] = p2_trig(c[C2_ox + 4 * (lane >> 1) + 3], lane & 1);   // double-precision functions lqng_trig_kernel calls)
                __syncwarp();
            }
            *reinterpret_cast<double2*>(r + P2_oQ + 4 * lane) = make_double2(0.0, 0.0);           // Q_0, Q_1 = 0 (128 doubles)
            *reinterpret_cast<double2*>(r + P2_oQ + 4 * lane + 2) = make_double2(0.0, 0.0);
            {   // A_i = I + dt [[0,0,cos h,-v sin h],[0,0,sin h,v cos h],0,0]  (KartLQRDynamics.cs:44-48), entry `lane` of A[2][4][4]
                const int i = lane >> 4, rr_ = (lane >> 2) & 3, cc = lane & 3;
                const double ch = c[C2_ocs + 2 * i], sh = c[C2_ocs + 2 * i + 1], v = c[C2_ox + 4 * i + 2];
                double a = rr_ == cc ? 1.0 : 0.0;
                if (rr_ == 0 && cc == 2) a = ch * p.dt;
                if (rr_ == 1 && cc == 2) a = sh * p.dt;
                if (rr_ == 0 && cc == 3) a = -sh * p.dt * v;
                if (rr_ == 1 && cc == 3) a = ch * p.dt * v;
                r[P2_oA + lane] = a;
            }
            if (lane < 16) {                                        // B_i (:57-59), q_i (KartLQRCosts.cs:109-124)
                const int e = lane & 7, rr_ = e >> 1, cc = e & 1;
                const int i = lane >> 3, s_ = lane & 7;
                // (cos h, sin h) = (0, 0) marks a decoupled DUMMY player (lqng_trig_kernel): B = 0, and A = I falls out of the formulas above
                const bool dummy = c[C2_ocs + 2 * i] == 0.0 && c[C2_ocs + 2 * i + 1] == 0.0;
                r[P2_oB + lane] = (!dummy && ((rr_ == 2 && cc == 0) || (rr_ == 3 && cc == 1))) ? p.dt : 0.0;
                double qv;
                if (s_ < 4) qv = (-c[C2_otg + 4 * i + s_]) * c[C2_otw + 4 * i + s_];
                else { qv = c[C2_oot + 4 * i + s_ - 4]; if (s_ < 7) qv = qv * -c[C2_oow + 3 * i + s_ - 4]; }
                r[P2_oq + lane] = qv;
            } else if (lane < 24) {                                 // R_i = controlWeight I (:136), x0
                const int e = lane - 16, i = e >> 2;
                r[P2_oR + e] = ((e & 3) == 0 || (e & 3) == 3) ? c[C2_ocw + i] : 0.0;
                r[P2_ox + e] = c[C2_ox + e];
            }
            __syncwarp();
            if (lane < 24) {                                        // the 12 non-zeros of each Q_i (:64-94; diag 4,5 overwritten: quirk Q4)
                const int i = lane / 12, j = lane % 12;
                double* Qi = r + P2_oQ + 64 * i;
                const double* aw = c + C2_oaw + 2 * i;
                if (j < 4) Qi[j * 9] = (j < 2 ? (0.0 - aw[j]) : 0.0) + c[C2_otw + 4 * i + j];
                else if (j < 7) Qi[j * 9] = -c[C2_oow + 3 * i + j - 4];
                else if (j == 7) {}
                else if (j < 10) Qi[(j - 8) * 8 + 4 + (j - 8)] = aw[j - 8];
                else Qi[(4 + j - 10) * 8 + (j - 10)] = aw[j - 10];
            }
            __syncwarp();
        }

        bool pivot = false, hard = false;
        double u_out = 0.0;
        for (;;) {                                                  // pass 0: no row exchanges; pass 1 (rare): partial pivoting
        // ---- operands of this problem (TV: of the current stage, reloaded every step) --------------------------------------------
        double aT0, aT1, xb00, xb01, xb10, xb11, yL0, yL1;          // yL: [B | A x0 | 0 0 0] in T-form
        double2 bFa, bFb, rr, rl;
        auto load_xb = [&](const double* r) { xb00 = r[offXb0]; xb01 = r[offXb0 + 2]; xb10 = r[offXb1]; xb11 = r[offXb1 + 2]; };
        auto load_ops = [&](const double* r, bool with_ax) {        // with_ax: A x0 rides in column 4 of the L product (needed at t = 0)
            aT0 = r[offA]; aT1 = r[offA + 4];
            load_xb(r);
            bFa = *reinterpret_cast<const double2*>(r + offBF); bFb = *reinterpret_cast<const double2*>(r + offBF + 2);
            rr = *reinterpret_cast<const double2*>(r + offRr);
            rl = *reinterpret_cast<const double2*>(r + offRl);
            double ax0 = 0.0, ax1 = 0.0;
            if (with_ax) {
                const double2 a0 = *reinterpret_cast<const double2*>(r + offAx), a1 = *reinterpret_cast<const double2*>(r + offAx + 2);
                const double2 a2 = *reinterpret_cast<const double2*>(r + offAx + 4), a3 = *reinterpret_cast<const double2*>(r + offAx + 6);
                const double2 xa = *reinterpret_cast<const double2*>(r + P2_ox + 4 * pl), xc = *reinterpret_cast<const double2*>(r + P2_ox + 4 * pl + 2);
                ax0 = fma(a1.y, xc.y, fma(a1.x, xc.x, fma(a0.y, xa.y, a0.x * xa.x)));
                ax1 = fma(a3.y, xc.y, fma(a3.x, xc.x, fma(a2.y, xa.y, a2.x * xa.x)));
            }
            yL0 = g == 4 ? ax0 : r[offB];
            yL1 = g == 4 ? ax1 : r[offB + 2];
        };
        // symmetry of Q_i and R_i is what lets Z_i's R-form stand in for its T-form
        auto asym = [&](const double* r, const double2& q0, const double2& q1) -> bool {
            return bits_differ(r[(2 * t) * 8 + g], q0.x) | bits_differ(r[(2 * t + 1) * 8 + g], q0.y) |
                   bits_differ(r[64 + (2 * t) * 8 + g], q1.x) | bits_differ(r[64 + (2 * t + 1) * 8 + g], q1.y) |
                   bits_differ(r[P2_oR + 1], r[P2_oR + 2]) | bits_differ(r[P2_oR + 5], r[P2_oR + 6]);
        };
        const double* rs = rec + (TV ? (size_t)p.horizon * P2_STRIDE : 0);      // record of the stage being processed
        load_ops(rs, !TV || p.horizon == 0);
        double z00, z01, z10, z11, e0, e1;
        bool redo, redo_piv = false;                                // redo: the shared-memory algorithm must take this problem
        {
            const double2 q0 = *reinterpret_cast<const double2*>(rs + P2_oQ + 2 * lane);           // Z_i = Q_i (KartLQR.cs:62), R-form
            const double2 q1 = *reinterpret_cast<const double2*>(rs + P2_oQ + 64 + 2 * lane);
            const double2 qv = *reinterpret_cast<const double2*>(rs + offQv);                       // eta_i = q_i (:63), lanes (4+i, t)
            z00 = q0.x; z01 = q0.y; z10 = q1.x; z11 = q1.y; e0 = qv.x; e1 = qv.y;
            redo = false;                                           // assembled Q_i, R_i are symmetric by construction
            if (!COMPACT) redo = asym(rs, q0, q1);
        }
        // W: rows 0..3 = stacked B_i^T Z_i, rows 4+i = eta_i (+ Z_i beta from the second step on)
        double w0 = e0, w1 = e1;
        mm(w0, w1, xb00, xb01, z00, z01);
        mm(w0, w1, xb10, xb11, z10, z11);

        for (int step = p.horizon; step >= 0; --step) {             // KartLQR.cs:64
            const bool last = step == 0;
            if (TV && step != p.horizon) { rs = rec + (size_t)step * P2_STRIDE; load_ops(rs, last); }
            // L = W [B | A x0] with eta^T B in rows 4, 5 (RHSVec, :96), + R_i on the diagonal blocks (:78), identity in rows 6, 7
            double l0 = rl.x, l1 = rl.y;
            mm(l0, l1, g < 4 ? w0 : e0, g < 4 ? w1 : e1, yL0, yL1);
            // RM^T = A^T W^T (RHSMat, :89-95); not needed at t = 0, where only u0 = -LHS^-1 (RM x0 + rv) is
            double m0 = 0.0, m1 = 0.0;
            if (FULL || !last) mm(m0, m1, aT0, aT1, w0, w1);
            double M0 = l0, M1 = l1;
            {
                const double v0 = shfl_d(l0, srcRv), v1 = shfl_d(l1, srcRv);
                double rv = godd ? v1 : v0;
                if (last && !FULL) rv += l0;                        // l0 of an Aug lane = (RM x0)[rho]
                if (isAug) M0 = rv;
            }
            if (!pivot) {
            // fraction-free Gauss-Jordan: row_i <- pv row_i - LHS[i][k] row_k (the pivot row is only rescaled)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double mine = (k & 1) ? M1 : M0;
                const double c = shfl_d(mine, csrc + 8 * (k >> 1));
                const int rs = (k >> 1 ? rb2 : rb) + 4 * (k & 1);
                const double r0 = shfl_d(M0, rs), r1 = shfl_d(M1, rs);
                const double pv = shfl_d(mine, 8 * (k >> 1) + 4 * (k & 1) + (k >> 1));
                redo_piv |= (rho_chk > k) & (abs_hi(c) > abs_hi(pv));                       // MathNet would exchange rows
                const double cz = rho == k ? 0.0 : c;
                M0 = fma(pv, M0, -(cz * r0));
                M1 = fma(pv, M1, -(cz * r1));
            }
            {
                const double d = shfl_d(rodd ? M1 : M0, dsrc);
                // One range test per sweep instead of one per pivot: the rows carry the product of the pivots, so a zero pivot
                // leaves d = 0 and an overflow leaves inf / NaN; |d| in 2^-900 .. 2^900 keeps rcp_fast exact to 2^-60.
                redo |= (abs_hi(d) - 0x07b00000u) > 0x70800000u;
                const double rinv = rcp_fast(d);
                M0 *= rinv; M1 *= rinv;
            }
            } else {
                // Gauss-Jordan with partial pivoting, lane j < 9 owning column j of [LHS | I | RHSVec] (KartLQR.cs:104-105)
                double c0, c1, c2, c3;
                {
                    const int j = lane < 9 ? lane : 0, jq = j >> 1;            // LHS[i][j] sits in L lane (g = 2 jq + (i&1), t = i>>1)
                    const double a0 = shfl_d(M0, 8 * jq), b0 = shfl_d(M1, 8 * jq);
                    const double a1 = shfl_d(M0, 8 * jq + 4), b1 = shfl_d(M1, 8 * jq + 4);
                    const double a2 = shfl_d(M0, 8 * jq + 1), b2 = shfl_d(M1, 8 * jq + 1);
                    const double a3 = shfl_d(M0, 8 * jq + 5), b3 = shfl_d(M1, 8 * jq + 5);
                    const double v0 = shfl_d(M0, 2), v1 = shfl_d(M0, 6), v2 = shfl_d(M0, 10), v3 = shfl_d(M0, 14);
                    const bool od = j & 1;
                    c0 = od ? b0 : a0; c1 = od ? b1 : a1; c2 = od ? b2 : a2; c3 = od ? b3 : a3;
                    if (lane >= 4) { c0 = lane == 4 ? 1.0 : 0.0; c1 = lane == 5 ? 1.0 : 0.0; c2 = lane == 6 ? 1.0 : 0.0; c3 = lane == 7 ? 1.0 : 0.0; }
                    if (lane == 8) { c0 = v0; c1 = v1; c2 = v2; c3 = v3; }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    double p0 = shfl_d(c0, k), p1 = shfl_d(c1, k), p2 = shfl_d(c2, k), p3 = shfl_d(c3, k);   // column k, every lane
                    // pivot row: first largest |.| among rows k..3, brought to row k by an exchange
                    int pr = k;
                    double best = fabs(k == 0 ? p0 : k == 1 ? p1 : k == 2 ? p2 : p3);
                    if (k < 1 && fabs(p1) > best) { best = fabs(p1); pr = 1; }
                    if (k < 2 && fabs(p2) > best) { best = fabs(p2); pr = 2; }
                    if (k < 3 && fabs(p3) > best) { best = fabs(p3); pr = 3; }
                    double& ck = k == 0 ? c0 : k == 1 ? c1 : k == 2 ? c2 : c3;
                    double& pk = k == 0 ? p0 : k == 1 ? p1 : k == 2 ? p2 : p3;
                    if (k < 1 && pr == 1) { double tmp = ck; ck = c1; c1 = tmp; tmp = pk; pk = p1; p1 = tmp; }   // warp-uniform
                    if (k < 2 && pr == 2) { double tmp = ck; ck = c2; c2 = tmp; tmp = pk; pk = p2; p2 = tmp; }
                    if (k < 3 && pr == 3) { double tmp = ck; ck = c3; c3 = tmp; tmp = pk; pk = p3; p3 = tmp; }
                    redo |= bad_pivot(pk);                          // zero, denormal, tiny, huge, inf or NaN: shared-memory algorithm
                    const double rk = ck / pk;                      // row k of the reduced system, this lane's column
                    if (k != 0) c0 = fma(-p0, rk, c0);
                    if (k != 1) c1 = fma(-p1, rk, c1);
                    if (k != 2) c2 = fma(-p2, rk, c2);
                    if (k != 3) c3 = fma(-p3, rk, c3);
                    ck = rk;
                }
                // back to the fragment distribution: I lanes take Lambda[rho][2kap..2kap+1], Aug lanes take (LHS^-1 rhs)[rho]
                const int sa = isAug ? 8 : 4 + 2 * kap, sb = isAug ? 8 : 5 + 2 * kap;
                const double a0 = shfl_d(c0, sa), a1 = shfl_d(c1, sa), a2 = shfl_d(c2, sa), a3 = shfl_d(c3, sa);
                const double b0 = shfl_d(c0, sb), b1 = shfl_d(c1, sb), b2 = shfl_d(c2, sb), b3 = shfl_d(c3, sb);
                M0 = rho == 0 ? a0 : rho == 1 ? a1 : rho == 2 ? a2 : a3;
                M1 = rho == 0 ? b0 : rho == 1 ? b1 : rho == 2 ? b2 : b3;
            }
            if (!FULL && last) {                                    // optimal_control = -P x0 - alpha with the t = 0 gains (:121-126)
                u_out = -M0;
                break;
            }
            const double ae = shfl_d(M0, srcAe), ao = shfl_d(M0, srcAo);           // alpha of player t>>1
            double y0 = shfl_d(M0, srcLam), y1 = shfl_d(M1, srcLam);              // Lambda as B operand: Y[k][col] = Lambda[s(col)][k]
            y0 = ylane ? y0 : 0.0; y1 = ylane ? y1 : 0.0;
            double pe = 0.0, po = 0.0;                              // (P[2(t>>1)][g], P[2(t>>1)+1][g])   (:104-105)
            mm(pe, po, m0, m1, y0, y1);
            const double pc = podd ? po : pe;                       // compact P: P[t][g]
            if (FULL) {                                             // P_step [4][8] (one coalesced 256-byte store), alpha_step [4]
                if (!waited) { asm volatile("griddepcontrol.wait;" ::: "memory"); waited = true; }   // warp-uniform
                if (p.horizon < P2_ROLL_STEPS) {                    // the rollout below reads the gains from here
                    sm.roll[wib][16 + step * 36 + t * 8 + g] = pc;
                    if (g == 0 && !(t & 1)) { sm.roll[wib][16 + step * 36 + 32 + t] = ae; sm.roll[wib][16 + step * 36 + 33 + t] = ao; }
                }
                if (p.P) p.P[((size_t)prob * (p.horizon + 1) + step) * 32 + t * 8 + g] = pc;
                if (p.alpha && g == 0 && !(t & 1)) {
                    double* a = p.alpha + ((size_t)prob * (p.horizon + 1) + step) * 4 + t;
                    a[0] = ae; a[1] = ao;
                }
                if (last) break;
            }
            // F = A - B P (T-form), beta = -B alpha   (:110-111)
            const double f0 = fma(-bFa.y, po, fma(-bFa.x, pe, aT0));
            const double f1 = fma(-bFb.y, po, fma(-bFb.x, pe, aT1));
            const double be0 = fma(-bFa.y, ao, -bFa.x * ae);
            const double be1 = fma(-bFb.y, ao, -bFb.x * ae);
            const double rpc = fma(rr.y, po, rr.x * pe);            // (R_p P_p)[t&1][g]
            // Z_i <- Q_i + P_i^T R_i P_i + F^T (Z_i F)   (:116)
            {
                double ya0 = 0.0, ya1 = 0.0, yb0 = 0.0, yb1 = 0.0;
                mm(ya0, ya1, f0, f1, z00, z01);
                mm(yb0, yb1, f0, f1, z10, z11);
                const double2 q0 = *reinterpret_cast<const double2*>(rs + P2_oQ + 2 * lane);
                const double2 q1 = *reinterpret_cast<const double2*>(rs + P2_oQ + 64 + 2 * lane);
                if (TV && step != p.horizon) redo |= asym(rs, q0, q1);
                z00 = q0.x; z01 = q0.y; z10 = q1.x; z11 = q1.y;
                dmma(z00, z01, pl == 0 ? pc : 0.0, rpc);
                dmma(z10, z11, pl == 1 ? pc : 0.0, rpc);
                mm(z00, z01, f0, f1, ya0, ya1);
                mm(z10, z11, f0, f1, yb0, yb1);
            }
            // W for the next step, beta^T in row 4+i: row 4+i comes out as eta_i + Z_i^{new} beta (quirk Q2)
            const double ra = fma(rr.y, ao, rr.x * ae);             // (R_i alpha_i)[t&1], this step's R_i
            if (TV) load_xb(rec + (size_t)(step - 1) * P2_STRIDE);  // rows 0..3 of W are B_i^T Z_i with the NEXT step's B_i
            w0 = e0; w1 = e1;
            mm(w0, w1, g == 4 ? be0 : xb00, g == 4 ? be1 : xb01, z00, z01);
            mm(w0, w1, g == 5 ? be0 : xb10, g == 5 ? be1 : xb11, z10, z11);
            // eta_i <- q_i + P_i^T R_i alpha_i + F^T (eta_i + Z_i^{new} beta)   (:117)
            {
                const double2 qv = *reinterpret_cast<const double2*>(rs + offQv);
                double n0 = qv.x, n1 = qv.y;
                dmma(n0, n1, (vec && pl == vp) ? ra : 0.0, pc);
                mm(n0, n1, vec ? w0 : 0.0, vec ? w1 : 0.0, f0, f1);
                e0 = n0; e1 = n1;
            }
        }
        hard = __ballot_sync(0xffffffffu, redo) != 0;
        if (hard || pivot || __ballot_sync(0xffffffffu, redo_piv) == 0) break;
        pivot = true;
        }
        if (it == 0) asm volatile("griddepcontrol.wait;" ::: "memory");   // outputs of the previous launch are complete from here on
        if (hard) {
            // Warp-uniform and rare: non-symmetric Q/R or a zero / out-of-range pivot.
            if (lane == 0) while (atomicCAS(&sm.lock, 0, 1) != 0) {}
            __syncwarp();
            if (COMPACT) {
                LqngParams q = p;                                   // the assembled record is the only dense copy: solve it in place
                double* r = &sm.rec[wib][0][0];
                q.A = r + P2_oA; q.B = r + P2_oB; q.Q = r + P2_oQ; q.q = r + P2_oq; q.R = r + P2_oR; q.x0 = r + P2_ox;
                q.u0 = p.u0 + (size_t)prob * 4; q.status = p.status ? p.status + prob : nullptr; q.time_varying = 0;
                if (lane < 8) lqng_generic_body<2>(q, 0, true, sm.fallback, lane, 0xffu);
            } else if (lane < 8) lqng_generic_body<2>(p, prob, true, sm.fallback, lane, 0xffu);
            __syncwarp();
            if (lane == 0) { __threadfence_block(); atomicExch(&sm.lock, 0); }
        } else if (FULL) {
            // u_t = -P_t x_t - alpha_t, x_{t+1} = A x_t + B u_t in forward time order (t = 0 is the pair computed last); gains from the
            // warp's shared-memory copy (up to 8 steps), else re-read from the output buffers this warp has just written (the launcher
            // lends scratch when the caller keeps none)
            const int T = p.horizon + 1;
            const bool sg = T <= P2_ROLL_STEPS;
            const double* gs = &sm.roll[wib][16];
            const double* gP = p.P + (size_t)prob * T * 32;
            const double* ga = p.alpha + (size_t)prob * T * 4;
            double* gt = p.traj ? p.traj + (size_t)prob * (T + 1) * 8 : nullptr;
            double* xs = &sm.roll[wib][0];
            double* us = xs + 8;
            if (lane < 8) { const double v = rec[P2_ox + lane]; xs[lane] = v; if (gt) gt[lane] = v; }
            __syncwarp();
            for (int st = 0; st < (gt ? T : 1); ++st) {
                double sacc = (sg ? gs[st * 36 + t * 8 + g] : gP[st * 32 + t * 8 + g]) * xs[g];
                sacc += __shfl_xor_sync(0xffffffffu, sacc, 4);
                sacc += __shfl_xor_sync(0xffffffffu, sacc, 8);
                sacc += __shfl_xor_sync(0xffffffffu, sacc, 16);
                const double u = -sacc - (sg ? gs[st * 36 + 32 + t] : ga[st * 4 + t]);
                if (g == 0) { us[t] = u; if (st == 0) p.u0[(size_t)prob * 4 + t] = u; }
                if (!gt) break;
                __syncwarp();
                double xn = 0.0;
                if (lane < 8) {
                    const int pr = lane >> 2, rr_ = lane & 3;
#pragma unroll
                    const double* rt = rec + (TV ? (size_t)st * P2_STRIDE : 0);
#pragma unroll
                    for (int c = 0; c < 4; ++c) xn = fma(rt[P2_oA + pr * 16 + rr_ * 4 + c], xs[4 * pr + c], xn);
#pragma unroll
                    for (int c = 0; c < 2; ++c) xn = fma(rt[P2_oB + pr * 8 + rr_ * 2 + c], us[2 * pr + c], xn);
                }
                __syncwarp();
                if (lane < 8) { xs[lane] = xn; gt[(size_t)(st + 1) * 8 + lane] = xn; }
                __syncwarp();
            }
            if (lane == 0 && p.status) p.status[prob] = 0;
        } else {
            if (isAug) p.u0[(size_t)prob * 4 + g] = u_out;
            if (lane == 0 && p.status) p.status[prob] = 0;
        }
        prob = nxt;
        if (dynamic) nxt = nxt < p.batch ? nwarps + (long long)__shfl_sync(0xffffffffu, after, 0) : nxt;
        else nxt += nwarps;
    }
    finish();
}

}  // namespace hk
