// Second round of FP64 microbenchmarks for the LQNG kernel design (B200, sm_100a):
//   long-running DFMA / DMMA peaks (>= 50 ms so that clocks settle), DFMA||DMMA mixes at several granularities,
//   double reciprocal / division cost, DMMA + SHFL overlap, and the larger sm_90+ f64 MMA shapes.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double* c, const double* a, double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// mode 0: DFMA only; 1: DMMA only; 2: even warps DFMA / odd warps DMMA; 3: per iteration 8 DMMA then 64 DFMA (equal flops);
// 4: per iteration 1 DMMA + 4 SHFL; 5: 1 DMMA + 8 SHFL; 6: 8 DFMA + 1 DMMA fine interleave
template <int MODE>
__global__ void k_mix(double* out, long long* cyc, double a, double b, int iters) {
    double x[8], c[8][2];
    unsigned v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { x[j] = threadIdx.x + j; c[j][0] = threadIdx.x * 0.5 + j; c[j][1] = j; v[j] = threadIdx.x * (2 * j + 3); }
    const int warp = threadIdx.x >> 5;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0 || (MODE == 2 && (warp & 1) == 0)) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = fma(x[j], a, b);
        } else if (MODE == 1 || (MODE == 2 && (warp & 1) == 1)) {
#pragma unroll
            for (int j = 0; j < 8; ++j) dmma884(c[j][0], c[j][1], a, b);
        } else if (MODE == 3) {
#pragma unroll
            for (int j = 0; j < 8; ++j) dmma884(c[j][0], c[j][1], a, b);
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = fma(x[j], a, b);
        } else if (MODE == 4 || MODE == 5) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                dmma884(c[j][0], c[j][1], a, b);
#pragma unroll
                for (int s = 0; s < (MODE == 4 ? 4 : 8); ++s) v[s] = __shfl_xor_sync(0xffffffffu, v[s], 1);
            }
        } else if (MODE == 6) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                dmma884(c[j][0], c[j][1], a, b);
#pragma unroll
                for (int s = 0; s < 8; ++s) x[s] = fma(x[s], a, b);
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j] + c[j][0] + c[j][1] + (double)v[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// reciprocal / division chains: MODE 0: x = 1.0/x + c (IEEE div), 1: __drcp_rn, 2: fp32 seed + 2 Newton steps
template <int MODE>
__global__ void k_rcp(double* out, long long* cyc, double cst, int iters) {
    double x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = 1.5 + threadIdx.x * 1e-3 + j;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (MODE == 0) x[j] = 1.0 / x[j] + cst;
            else if (MODE == 1) x[j] = __drcp_rn(x[j]) + cst;
            else { double y = (double)__frcp_rn((float)x[j]); y = y * fma(-x[j], y, 2.0); y = y * fma(-x[j], y, 2.0); x[j] = y + cst; }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x[0] + x[1] + x[2] + x[3];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int SHAPE>   // 0: m16n8k4, 1: m16n8k8, 2: m16n8k16
__global__ void k_bigmma(double* out, long long* cyc, double a, double b, int iters) {
    double c[4][4], av[8], bv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) for (int k = 0; k < 4; ++k) c[j][k] = threadIdx.x + j + k;
#pragma unroll
    for (int j = 0; j < 8; ++j) av[j] = a + j * 1e-9;
#pragma unroll
    for (int j = 0; j < 4; ++j) bv[j] = b + j * 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (SHAPE == 0) dmma1684(c[j], av, bv[0]);
            else if (SHAPE == 1) dmma1688(c[j], av, bv);
            else dmma16816(c[j], av, bv);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) for (int k = 0; k < 4; ++k) s += c[j][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

struct Res { double ms, cyc; };
template <typename F> Res run(F launch, int blocks, long long* d_cyc) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
    }
    std::vector<long long> h(blocks);
    CK(cudaMemcpy(h.data(), d_cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end());
    return {(double)best, (double)h[blocks / 2]};
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount, TPB = 256, BPS = 4, blocks = sms * BPS;
    double* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, sizeof(double) * blocks * TPB)); CK(cudaMalloc(&d_cyc, sizeof(long long) * blocks));
    // warm the clocks: ~0.5 s of DFMA
    for (int i = 0; i < 20; ++i) k_mix<0><<<blocks, TPB>>>(d_out, d_cyc, 1.0000001, 1e-9, 40000);
    CK(cudaDeviceSynchronize());
    auto rep = [&](const char* name, Res r, double flops_per_thread, double fp64_instr_per_thread, const char* note) {
        printf("{\"bench\": \"%s\", \"ms\": %.3f, \"mhz\": %.0f, \"tflops\": %.3f, \"fp64_warp_instr_per_clk_per_sm\": %.4f, \"note\": \"%s\"}\n", name, r.ms,
               r.cyc / (r.ms * 1e3), flops_per_thread * blocks * TPB / (r.ms * 1e-3) / 1e12, fp64_instr_per_thread * (BPS * TPB / 32.0) / r.cyc, note);
        fflush(stdout);
    };
    const int IT = 60000;
    { Res r = run([&] { k_mix<0><<<blocks, TPB>>>(d_out, d_cyc, 1.0000001, 1e-9, IT); }, blocks, d_cyc); rep("dfma_long", r, 128.0 * IT, 64.0 * IT, "64 DFMA/iter"); }
    { Res r = run([&] { k_mix<1><<<blocks, TPB>>>(d_out, d_cyc, 1.0000001, 1e-9, IT * 4); }, blocks, d_cyc); rep("dmma_long", r, 8 * 16.0 * IT * 4, 8.0 * IT * 4, "8 DMMA m8n8k4/iter (512 flop each per warp)"); }
    { Res r = run([&] { k_mix<2><<<blocks, TPB>>>(d_out, d_cyc, 1.0000001, 1e-9, IT); }, blocks, d_cyc); rep("mix_warp_split", r, 0.5 * (128.0 + 128.0) * IT, 0.5 * (64.0 + 8.0) * IT, "even warps 64 DFMA, odd warps 8 DMMA per iter (equal flops)"); }
    { Res r = run([&] { k_mix<3><<<blocks, TPB>>>(d_out, d_cyc, 1.0000001, 1e-9, IT); }, blocks, d_cyc); rep("mix_block_8dmma_64dfma", r, 256.0 * IT, 72.0 * IT, "same warp: 8 DMMA then 64 DFMA"); }
    { Res r = run([&] { k_mix<6><<<blocks, TPB>>>(d_out, d_cyc, 1.0000001, 1e-9, IT); }, blocks, d_cyc); rep("mix_fine_1dmma_8dfma", r, 256.0 * IT, 72.0 * IT, "same warp: (1 DMMA, 8 DFMA) x 8"); }
    { Res r = run([&] { k_mix<4><<<blocks, TPB>>>(d_out, d_cyc, 1.0000001, 1e-9, IT * 2); }, blocks, d_cyc); rep("dmma_shfl4", r, 128.0 * IT * 2, 8.0 * IT * 2, "1 DMMA + 4 SHFL"); }
    { Res r = run([&] { k_mix<5><<<blocks, TPB>>>(d_out, d_cyc, 1.0000001, 1e-9, IT * 2); }, blocks, d_cyc); rep("dmma_shfl8", r, 128.0 * IT * 2, 8.0 * IT * 2, "1 DMMA + 8 SHFL"); }
    { Res r = run([&] { k_rcp<0><<<blocks, TPB>>>(d_out, d_cyc, 0.25, IT / 2); }, blocks, d_cyc); rep("ddiv", r, 0, 4.0 * IT / 2, "instr = divisions (1.0/x + c); clk/SM per warp-division = 1/this"); }
    { Res r = run([&] { k_rcp<1><<<blocks, TPB>>>(d_out, d_cyc, 0.25, IT / 2); }, blocks, d_cyc); rep("drcp_rn", r, 0, 4.0 * IT / 2, "instr = __drcp_rn"); }
    { Res r = run([&] { k_rcp<2><<<blocks, TPB>>>(d_out, d_cyc, 0.25, IT / 2); }, blocks, d_cyc); rep("rcp_f32seed_newton2", r, 0, 4.0 * IT / 2, "instr = fp32 rcp seed + 2 Newton (approx 1e-14)"); }
    { Res r = run([&] { k_bigmma<0><<<blocks, TPB>>>(d_out, d_cyc, 1.0000001, 1e-9, IT * 2); }, blocks, d_cyc); rep("dmma_m16n8k4", r, 4 * 32.0 * IT * 2, 4.0 * IT * 2, "1024 flop per warp-instr"); }
    { Res r = run([&] { k_bigmma<1><<<blocks, TPB>>>(d_out, d_cyc, 1.0000001, 1e-9, IT); }, blocks, d_cyc); rep("dmma_m16n8k8", r, 4 * 64.0 * IT, 4.0 * IT, "2048 flop per warp-instr"); }
    { Res r = run([&] { k_bigmma<2><<<blocks, TPB>>>(d_out, d_cyc, 1.0000001, 1e-9, IT / 2); }, blocks, d_cyc); rep("dmma_m16n8k16", r, 4 * 128.0 * IT / 2, 4.0 * IT / 2, "4096 flop per warp-instr"); }
    return 0;
}
