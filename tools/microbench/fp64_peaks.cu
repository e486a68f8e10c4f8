// Microbenchmarks that decide the LQNG kernel mapping on B200 (sm_100a):
//   dfma      : FP64 vector FMA peak (the roofline denominator SURVEY.md §6 asks the builder to measure)
//   dmma      : mma.sync m8n8k4 f64 (DMMA) peak
//   dfma+dmma : do the two pipes add up?
//   shfl      : warp shuffle issue rate
//   dfma+shfl : overlap of shuffles under a DFMA stream
//   lds128    : shared-memory 128-bit loads, broadcast patterns (1, 4, 32 distinct addresses per warp)
//   dfma+lds  : DFMA fed by broadcast LDS.128 (1 load per 2 DFMA)
// Every kernel reports (a) wall time via CUDA events -> TFLOP/s or Ginstr/s and
// (b) per-SM cycles via clock64 -> lane-ops per clock per SM, which is independent of the clock the box ran at.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ITERS = 4096;

__global__ void k_dfma(double* out, long long* cyc, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < ITERS; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void k_dmma(double* out, long long* cyc, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j][0] = threadIdx.x + j; c[j][1] = j; }
    a += threadIdx.x * 1e-9; b += threadIdx.x * 1e-9;
    long long t0 = clock64();
#pragma unroll 2
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma(c[j][0], c[j][1], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// 8 DFMA + 1 DMMA (each DMMA = 8 DFMA-warp-instructions worth of flops)
__global__ void k_dfma_dmma(double* out, long long* cyc, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    double c[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) { c[j][0] = threadIdx.x + j; c[j][1] = j; }
    long long t0 = clock64();
#pragma unroll 2
    for (int i = 0; i < ITERS; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        dmma(c[i & 3][0], c[i & 3][1], a, b);
    }
    long long t1 = clock64();
    double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_shfl(double* out, long long* cyc, int lane_xor) {
    unsigned v0 = threadIdx.x, v1 = v0 * 3, v2 = v0 * 5, v3 = v0 * 7, v4 = v0 * 11, v5 = v0 * 13, v6 = v0 * 17, v7 = v0 * 19;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < ITERS; ++i) {
        v0 = __shfl_xor_sync(0xffffffffu, v0, lane_xor); v1 = __shfl_xor_sync(0xffffffffu, v1, lane_xor);
        v2 = __shfl_xor_sync(0xffffffffu, v2, lane_xor); v3 = __shfl_xor_sync(0xffffffffu, v3, lane_xor);
        v4 = __shfl_xor_sync(0xffffffffu, v4, lane_xor); v5 = __shfl_xor_sync(0xffffffffu, v5, lane_xor);
        v6 = __shfl_xor_sync(0xffffffffu, v6, lane_xor); v7 = __shfl_xor_sync(0xffffffffu, v7, lane_xor);
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = (double)(v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// NS shuffles (32-bit) per 8 DFMA
template <int NS>
__global__ void k_dfma_shfl(double* out, long long* cyc, double a, double b, int lane_xor) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    unsigned v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = threadIdx.x * (2 * j + 3);
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < ITERS; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
#pragma unroll
        for (int j = 0; j < NS; ++j) v[j & 7] = __shfl_xor_sync(0xffffffffu, v[j & 7], lane_xor);
    }
    long long t1 = clock64();
    double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += (double)v[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ double2 lds128(const double2* p) {
    double2 v; unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}

// LDS.128: GROUP lanes share one 16-byte address (GROUP = 32: full broadcast, 8: four addresses, 1: all distinct)
template <int GROUP>
__global__ void k_lds128(double* out, long long* cyc) {
    __shared__ double2 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_double2(i, -i);
    __syncthreads();
    int base = ((threadIdx.x / GROUP) * 8) & 1023;   // stride of 8 x 16 B = one full bank sweep per group -> no bank conflicts between groups only if GROUP>=8
    if (GROUP == 1) base = threadIdx.x & 1023;
    double s0 = 0, s1 = 0;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < ITERS; ++i) {
        double2 v = lds128(&sm[(base + (i & 7)) & 1023]);
        s0 += v.x; s1 += v.y;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// Same but groups read consecutive 16-byte slots (4 groups -> 64 contiguous bytes): the layout the LQNG kernel would use
template <int GROUP>
__global__ void k_lds128_packed(double* out, long long* cyc) {
    __shared__ double2 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_double2(i, -i);
    __syncthreads();
    int g = threadIdx.x / GROUP;
    double x0 = 0, x1 = 0, x2 = 0, x3 = 0;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < ITERS; ++i) {
        // 2 loads, 4 DFMA  (ratio of the row-owner matmul: one broadcast double per DFMA)
        double2 v = lds128(&sm[(g + 32 * (i & 31)) & 2047]);
        double2 w = lds128(&sm[(g + 32 * (i & 31) + 1024) & 2047]);
        x0 = fma(v.x, x0, v.y); x1 = fma(v.y, x1, v.x); x2 = fma(w.x, x2, w.y); x3 = fma(w.y, x3, w.x);
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

struct Res { double ms; double cyc; };

template <typename F>
Res run(F launch, int blocks, long long* d_cyc) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
    }
    std::vector<long long> h(blocks);
    CK(cudaMemcpy(h.data(), d_cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end());
    return { (double)best, (double)h[blocks / 2] };
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz_prop\": %d}\n", p.name, sms, p.clockRate);
    const int TPB = 256, BPS = 4;           // 1024 threads per SM = 8 warps per SMSP
    int blocks = sms * BPS;
    double* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, sizeof(double) * blocks * TPB)); CK(cudaMalloc(&d_cyc, sizeof(long long) * blocks));
    double warps_per_sm = BPS * TPB / 32.0;
    auto report = [&](const char* name, Res r, double lane_ops_per_thread, double flops_per_thread, const char* note) {
        double total_threads = (double)blocks * TPB;
        double per_sm_per_clk = lane_ops_per_thread * (BPS * TPB) / r.cyc;
        printf("{\"bench\": \"%s\", \"ms\": %.4f, \"median_block_cycles\": %.0f, \"eff_mhz\": %.0f, \"lane_ops_per_clk_per_sm\": %.2f, \"warp_instr_per_clk_per_sm\": %.3f, \"tflops\": %.3f, \"note\": \"%s\"}\n",
               name, r.ms, r.cyc, r.cyc / (r.ms * 1e3), per_sm_per_clk, per_sm_per_clk / 32.0,
               flops_per_thread * total_threads / (r.ms * 1e-3) / 1e12, note);
        (void)warps_per_sm;
    };
    {
        Res r = run([&] { k_dfma<<<blocks, TPB>>>(d_out, d_cyc, 1.0000001, 1e-9); }, blocks, d_cyc);
        report("dfma", r, 8.0 * ITERS, 16.0 * ITERS, "8 independent DFMA chains/thread, 32 warps/SM");
    }
    {
        Res r = run([&] { k_dmma<<<blocks, TPB>>>(d_out, d_cyc, 1.0000001, 1e-9); }, blocks, d_cyc);
        // one DMMA m8n8k4 = 256 FMA per warp = 8 FMA per lane
        report("dmma_m8n8k4", r, 8.0 * ITERS, 8.0 * ITERS * 16.0, "lane_ops counts DMMA instr per lane; tflops counts 512 flop per warp-instr");
    }
    {
        Res r = run([&] { k_dfma_dmma<<<blocks, TPB>>>(d_out, d_cyc, 1.0000001, 1e-9); }, blocks, d_cyc);
        report("dfma8_dmma1", r, 9.0 * ITERS, (16.0 + 16.0) * ITERS, "8 DFMA + 1 DMMA per iteration (equal flops from each)");
    }
    for (int x : {1, 16}) {
        Res r = run([&] { k_shfl<<<blocks, TPB>>>(d_out, d_cyc, x); }, blocks, d_cyc);
        report(x == 1 ? "shfl_xor1" : "shfl_xor16", r, 8.0 * ITERS, 0, "32-bit shuffles, 8 chains");
    }
    {
        Res r = run([&] { k_dfma_shfl<2><<<blocks, TPB>>>(d_out, d_cyc, 1.0000001, 1e-9, 1); }, blocks, d_cyc);
        report("dfma8_shfl2", r, 10.0 * ITERS, 16.0 * ITERS, "8 DFMA + 2 SHFL");
        r = run([&] { k_dfma_shfl<4><<<blocks, TPB>>>(d_out, d_cyc, 1.0000001, 1e-9, 1); }, blocks, d_cyc);
        report("dfma8_shfl4", r, 12.0 * ITERS, 16.0 * ITERS, "8 DFMA + 4 SHFL");
        r = run([&] { k_dfma_shfl<8><<<blocks, TPB>>>(d_out, d_cyc, 1.0000001, 1e-9, 1); }, blocks, d_cyc);
        report("dfma8_shfl8", r, 16.0 * ITERS, 16.0 * ITERS, "8 DFMA + 8 SHFL");
    }
    {
        Res r = run([&] { k_lds128<32><<<blocks, TPB>>>(d_out, d_cyc); }, blocks, d_cyc);
        report("lds128_bcast32", r, 1.0 * ITERS, 2.0 * ITERS, "all lanes one address; flops=DADD");
        r = run([&] { k_lds128<8><<<blocks, TPB>>>(d_out, d_cyc); }, blocks, d_cyc);
        report("lds128_bcast8", r, 1.0 * ITERS, 2.0 * ITERS, "4 addresses per warp (8-lane groups)");
        r = run([&] { k_lds128<1><<<blocks, TPB>>>(d_out, d_cyc); }, blocks, d_cyc);
        report("lds128_distinct", r, 1.0 * ITERS, 2.0 * ITERS, "32 distinct 16 B addresses");
    }
    {
        Res r = run([&] { k_lds128_packed<8><<<blocks, TPB>>>(d_out, d_cyc); }, blocks, d_cyc);
        report("dfma4_lds2_g8", r, 4.0 * ITERS, 8.0 * ITERS, "lane_ops counts DFMA; 2 LDS.128 (8-lane broadcast, packed) per 4 DFMA");
        r = run([&] { k_lds128_packed<4><<<blocks, TPB>>>(d_out, d_cyc); }, blocks, d_cyc);
        report("dfma4_lds2_g4", r, 4.0 * ITERS, 8.0 * ITERS, "same, 4-lane groups (8 addresses per warp)");
        r = run([&] { k_lds128_packed<16><<<blocks, TPB>>>(d_out, d_cyc); }, blocks, d_cyc);
        report("dfma4_lds2_g16", r, 4.0 * ITERS, 8.0 * ITERS, "same, 16-lane groups (2 addresses per warp)");
    }
    return 0;
}
