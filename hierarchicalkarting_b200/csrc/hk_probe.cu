// hk_probe.cu — the FP64 roofline denominator, measured on the box the benchmark runs on (include/hk_abi.h: hk_probe_fp64_peak).
// A DMMA m8n8k4 f64 stream (the instruction the LQNG kernels' flops are issued with; DFMA shares the same pipe and peaks ~10 % lower,
// profiles/fp64_microbench_r01.md): 32 warps per SM, 8 independent accumulator pairs per warp, run back to back for the requested time.
// Reports the flop rate over the whole run (CUDA events) and the SM clock the run actually had (clock64 cycles of the last launch / its
// event time), so that a burst figure (clocks at boost) and a sustained one (clocks settled under FP64 power) can be told apart.
#include "hk_common.cuh"
#include <vector>

namespace hk {

constexpr int PROBE_ITERS = 2048;

__device__ __forceinline__ void probe_dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) fp64_probe_kernel(double* out, long long* cyc, double a, double b, int rounds)
{
    double c[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j][0] = threadIdx.x + j; c[j][1] = j; }
    a += threadIdx.x * 1e-9; b += threadIdx.x * 1e-9;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
#pragma unroll 2
        for (int i = 0; i < PROBE_ITERS; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) probe_dmma(c[j][0], c[j][1], a, b);
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

}  // namespace hk

using namespace hk;

extern "C" int hk_probe_fp64_peak(double seconds, double* tflops, double* sm_mhz_effective, double* seconds_run)
{
    if (!(seconds > 0) || !tflops) { set_error("hk_probe_fp64_peak: seconds must be > 0"); return HK_ERR_INVALID_ARGUMENT; }
    ThreadCtx* c = ctx();
    if (!c) return HK_ERR_NO_DEVICE;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 4, threads = 256;                        // 32 warps per SM
    double* out = (double*)dscratch(c, 14, sizeof(double) * (size_t)blocks * threads + sizeof(long long) * blocks);
    if (!out) return HK_ERR_OUT_OF_MEMORY;
    long long* cyc = (long long*)(out + (size_t)blocks * threads);
    cudaStream_t s = c->stream;
    // one launch = rounds x 2048 x 8 DMMA per warp; size it to ~2 ms, then repeat launches until `seconds` have passed on the device
    const double flop_per_round = (double)blocks * (threads / 32) * PROBE_ITERS * 8.0 * 512.0;     // a DMMA m8n8k4 is 512 flops
    const int rounds = 4;                                             // ~2.2 ms at 37 TFLOP/s
    count_launch(); fp64_probe_kernel<<<blocks, threads, 0, s>>>(out, cyc, 1.0000001, 1e-9, 1);   // page in
    HK_CUDA(cudaStreamSynchronize(s));
    cudaEvent_t e0, e1, l0, l1;
    HK_CUDA(cudaEventCreate(&e0)); HK_CUDA(cudaEventCreate(&e1)); HK_CUDA(cudaEventCreate(&l0)); HK_CUDA(cudaEventCreate(&l1));
    HK_CUDA(cudaEventRecord(e0, s));
    long long launches = 0;
    float ms = 0.0f;
    for (;;) {
        const int burst = 8;
        for (int k = 0; k < burst; ++k) {
            if (k == burst - 1) HK_CUDA(cudaEventRecord(l0, s));
            count_launch(); fp64_probe_kernel<<<blocks, threads, 0, s>>>(out, cyc, 1.0000001, 1e-9, rounds);
        }
        HK_CUDA(cudaEventRecord(l1, s));
        HK_CUDA(cudaEventRecord(e1, s));
        HK_CUDA(cudaEventSynchronize(e1));
        launches += burst;
        HK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms * 1e-3 >= seconds) break;
    }
    float last_ms = 0.0f;
    HK_CUDA(cudaEventElapsedTime(&last_ms, l0, l1));
    std::vector<long long> hc((size_t)blocks);
    HK_CUDA(cudaMemcpy(hc.data(), cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (long long v : hc) mx = v > mx ? v : mx;
    *tflops = flop_per_round * rounds * (double)launches / (ms * 1e-3) / 1e12;
    if (sm_mhz_effective) *sm_mhz_effective = last_ms > 0 ? (double)mx / (last_ms * 1e-3) / 1e6 : 0.0;
    if (seconds_run) *seconds_run = ms * 1e-3;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(l0); cudaEventDestroy(l1);
    return HK_OK;
}
