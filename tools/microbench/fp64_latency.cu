// Latency / concurrency microbenchmarks for the LQNG DMMA kernel (B200): throughput of DEPENDENT DMMA / DFMA chains as a
// function of resident warps per SM sub-partition, and of a DMMA chain with a dependent DFMA chain interleaved.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// MODE 0: one dependent DMMA chain (D feeds C); 1: D feeds the A operand of the next DMMA (operand dependency);
// 2: one dependent DFMA chain; 3: dependent DMMA chain + dependent DFMA chain (independent of each other);
// 4: 2 independent DMMA chains; 5: dependent chain alternating DMMA -> DFMA -> DMMA (each feeds the next)
template <int MODE>
__global__ void k(double* out, long long* cyc, double a, double b, int iters) {
    double c0 = threadIdx.x, c1 = 1.0, d0 = 2.0, d1 = 3.0, x = threadIdx.x * 0.5;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (MODE == 0) dmma884(c0, c1, a, b);
            else if (MODE == 1) { double e0 = 0, e1 = 0; dmma884(e0, e1, c0, b); c0 = e0; c1 = e1; }
            else if (MODE == 2) x = fma(x, a, b);
            else if (MODE == 3) { dmma884(c0, c1, a, b); x = fma(x, a, b); x = fma(x, a, b); }
            else if (MODE == 4) { dmma884(c0, c1, a, b); dmma884(d0, d1, a, b); }
            else if (MODE == 5) { dmma884(c0, c1, a, b); c0 = fma(c0, a, b); }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + d0 + d1 + x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int sms, double* d_out, long long* d_cyc, double dmma_per_iter, double dfma_per_iter) {
    for (int warps_per_smsp : {1, 2, 3, 4, 5, 6, 8}) {
        const int tpb = 128 * warps_per_smsp;       // one CTA per SM, warps spread over the 4 sub-partitions
        const int iters = 20000;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        k<MODE><<<sms, tpb>>>(d_out, d_cyc, 1.0000001, 1e-9, iters); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); k<MODE><<<sms, tpb>>>(d_out, d_cyc, 1.0000001, 1e-9, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double ns_per_iter_group = ms * 1e6 / (iters * 8.0);   // time for one unrolled slot of one warp
        // per SM sub-partition instruction rates, in ns (clock-independent) and in 1.965 GHz cycles
        const double dmma_rate = dmma_per_iter * warps_per_smsp / ns_per_iter_group;   // DMMA per ns per SMSP
        printf("{\"bench\": \"%s\", \"warps_per_smsp\": %d, \"ns_per_slot\": %.2f, \"cycles_per_slot_at_1965\": %.1f, \"dmma_per_smsp_per_16clk\": %.3f, \"dfma_per_smsp_per_2.2clk\": %.3f}\n",
               name, warps_per_smsp, ns_per_iter_group, ns_per_iter_group * 1.965, dmma_rate * 16 / 1.965,
               dfma_per_iter * warps_per_smsp / ns_per_iter_group * 2.2 / 1.965);
        fflush(stdout);
    }
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, sizeof(double) * sms * 1024)); CK(cudaMalloc(&d_cyc, sizeof(long long) * sms));
    for (int i = 0; i < 10; ++i) k<2><<<sms, 1024>>>(d_out, d_cyc, 1.0000001, 1e-9, 100000);
    CK(cudaDeviceSynchronize());
    run<0>("dmma_chain_C", sms, d_out, d_cyc, 1, 0);
    run<1>("dmma_chain_A", sms, d_out, d_cyc, 1, 0);
    run<2>("dfma_chain", sms, d_out, d_cyc, 0, 1);
    run<3>("dmma_chain+2dfma_chain", sms, d_out, d_cyc, 1, 2);
    run<4>("dmma_2chains", sms, d_out, d_cyc, 2, 0);
    run<5>("dmma_dfma_alternating_chain", sms, d_out, d_cyc, 1, 1);
    return 0;
}
