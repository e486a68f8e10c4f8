// hk_common.cuh — shared host-side plumbing of libhk_b200 (error state, per-thread CUDA context).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include "../../include/hk_abi.h"

namespace hk {

void set_error(const char* fmt, ...);
void count_launch();                       // every kernel launch of this library (hk_kernel_launch_count)

// Per host-thread context: the reference calls the planners from several C# threads (HierarchicalKartAgent.cs:246-283),
// so every thread owns a stream and grow-only scratch buffers; nothing is shared but the immutable hk_game objects.
struct ThreadCtx {
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t cstream[4] = {nullptr, nullptr, nullptr, nullptr};   // host-to-device copy streams of the chunk pipelines
    cudaEvent_t pev[16] = {};                                 // ring events of the chunk pipelines: [slot] copied in, [8 + slot] drained
    void* dbuf[16] = {};
    size_t dcap[16] = {};
    void* hbuf[4] = {nullptr, nullptr, nullptr, nullptr};      // pinned
    size_t hcap[4] = {0, 0, 0, 0};
    bool ready = false;
    int device = -1;                                          // the device the streams / buffers belong to
    void release();
    unsigned launch_id = 0;                                   // alternates the redo counters of the LQNG fast path
    ~ThreadCtx();
};

int  ensure_device();                       // HK_OK or HK_ERR_NO_DEVICE / HK_ERR_CUDA
ThreadCtx* ctx();                           // nullptr on failure (error set)
void* dscratch(ThreadCtx* c, int slot, size_t bytes);     // nullptr on OOM (error set)
void* hscratch(ThreadCtx* c, int slot, size_t bytes);
void drain_ctx(ThreadCtx* c);                // synchronises every stream the context owns (error paths of the pipelines)

#define HK_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            hk::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);  \
            return e_ == cudaErrorMemoryAllocation ? HK_ERR_OUT_OF_MEMORY : HK_ERR_CUDA;                \
        }                                                                                               \
    } while (0)

// LQNG device entry (hk_lqng.cu)
int lqng_launch(int batch, int N, int horizon, int time_varying, const double* dA, const double* dB, const double* dQ,
                const double* dq, const double* dR, const double* dx0, double* du0, double* dP, double* dalpha,
                double* dtraj, int* dstatus, cudaStream_t stream, const int* gate = nullptr, int gate_min = 0);
int lqng_assemble_launch(int batch, int N, int horizon, double dt, const double* dx0, const double* dtarget,
                         const double* dtw, const double* dcw, const double* daw, const double* dotgt, const double* dotw,
                         double* du0, int* dstatus, cudaStream_t stream, int scratch_slot, const int* dn_players = nullptr, int min_players = 0);

int lqng_solve_packed_in_place(int batch, int horizon, double dt, const double* drec, double* du0, int* dstatus, cudaStream_t stream);
int lqng_assemble_launch_packed(int batch, int N, int horizon, double dt, const double* drec, double* du0, int* dstatus, cudaStream_t stream,
                                int scratch_slot, const int* dn_players = nullptr, const double* dcs_ready = nullptr);

}  // namespace hk
