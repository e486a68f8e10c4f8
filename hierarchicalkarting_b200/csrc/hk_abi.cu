// hk_abi.cu — host side of the C-ABI (include/hk_abi.h): library state, per-thread CUDA context, and the LQNG entry
// points.  The discrete-game entry points live next to their kernels in hk_game.cu.
#include "hk_common.cuh"
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <atomic>

namespace hk {

static thread_local char g_err[512] = "";
static std::mutex g_mu;
static std::atomic<int> g_device{-1};          // -1 = not initialised
static std::atomic<int> g_device_state{0};     // 0 unknown, 1 ok, -1 no usable device
static std::atomic<long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int ensure_device()
{
    int st = g_device_state.load();
    if (st == 1) {
        cudaError_t e = cudaSetDevice(g_device.load());
        if (e != cudaSuccess) { set_error("cudaSetDevice: %s", cudaGetErrorString(e)); return HK_ERR_CUDA; }
        return HK_OK;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_device_state.load() == 1) {                  // lost the initialisation race: still bind THIS thread to the device
        cudaError_t e2 = cudaSetDevice(g_device.load());
        if (e2 != cudaSuccess) { set_error("cudaSetDevice: %s", cudaGetErrorString(e2)); return HK_ERR_CUDA; }
        return HK_OK;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n < 1) {
        set_error("no CUDA device (%s); libhk_b200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        cudaGetLastError();
        return HK_ERR_NO_DEVICE;
    }
    int dev = g_device.load();
    if (dev < 0) dev = 0;
    if (dev >= n) { set_error("device %d requested, %d present", dev, n); return HK_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) { set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return HK_ERR_CUDA; }
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libhk_b200 is built for sm_100a (B200) only", dev, prop.major, prop.minor);
        return HK_ERR_NO_DEVICE;
    }
    e = cudaSetDevice(dev);
    if (e != cudaSuccess) { set_error("cudaSetDevice: %s", cudaGetErrorString(e)); return HK_ERR_CUDA; }
    g_device.store(dev);
    g_device_state.store(1);
    return HK_OK;
}

void ThreadCtx::release()
{
    // Streams/buffers of this host thread; ignore errors (the context may already be gone at process exit).
    if (!ready) return;
    for (auto& b : dbuf) { if (b) cudaFree(b); b = nullptr; }
    for (auto& cp : dcap) cp = 0;
    for (auto& b : hbuf) { if (b) cudaFreeHost(b); b = nullptr; }
    for (auto& cp : hcap) cp = 0;
    for (auto& e : ev) { if (e) cudaEventDestroy(e); e = nullptr; }
    for (auto& e : pev) { if (e) cudaEventDestroy(e); e = nullptr; }
    for (auto& st : cstream) { if (st) cudaStreamDestroy(st); st = nullptr; }
    if (stream) cudaStreamDestroy(stream);
    if (stream2) cudaStreamDestroy(stream2);
    stream = stream2 = nullptr;
    ready = false;
}

ThreadCtx::~ThreadCtx() { release(); }

void drain_ctx(ThreadCtx* c)
{
    if (!c || !c->ready) return;
    cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->stream2);
    for (auto& st : c->cstream) if (st) cudaStreamSynchronize(st);
}

static ThreadCtx& thread_ctx() { static thread_local ThreadCtx c; return c; }

ThreadCtx* ctx()
{
    ThreadCtx& c = thread_ctx();
    if (ensure_device() != HK_OK) return nullptr;
    if (c.ready && c.device != g_device.load()) c.release();      // hk_shutdown() + hk_init(other device): streams / buffers of the old device
    if (!c.ready) {
        c.device = g_device.load();
        if (cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c.stream2, cudaStreamNonBlocking) != cudaSuccess) {
            set_error("cudaStreamCreate failed");
            return nullptr;
        }
        for (auto& e : c.ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        for (auto& e : c.pev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        for (auto& st : c.cstream)
            if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); return nullptr; }
        c.ready = true;
    }
    return &c;
}

void* dscratch(ThreadCtx* c, int slot, size_t bytes)
{
    if (bytes <= c->dcap[slot]) return c->dbuf[slot];
    if (c->dbuf[slot]) { cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->stream2); cudaFree(c->dbuf[slot]); c->dbuf[slot] = nullptr; c->dcap[slot] = 0; }
    size_t cap = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&c->dbuf[slot], cap);
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu): %s", cap, cudaGetErrorString(e)); cudaGetLastError(); return nullptr; }
    c->dcap[slot] = cap;
    return c->dbuf[slot];
}

void* hscratch(ThreadCtx* c, int slot, size_t bytes)
{
    if (bytes <= c->hcap[slot]) return c->hbuf[slot];
    if (c->hbuf[slot]) { cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->stream2); cudaFreeHost(c->hbuf[slot]); c->hbuf[slot] = nullptr; c->hcap[slot] = 0; }
    size_t cap = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMallocHost(&c->hbuf[slot], cap);
    if (e != cudaSuccess) { set_error("cudaMallocHost(%zu): %s", cap, cudaGetErrorString(e)); cudaGetLastError(); return nullptr; }
    c->hcap[slot] = cap;
    return c->hbuf[slot];
}

}  // namespace hk

using namespace hk;

extern "C" int hk_abi_version(void) { return HK_ABI_VERSION; }
extern "C" const char* hk_last_error(void) { return g_err; }

extern "C" long long hk_kernel_launch_count(void) { return g_launches.load(); }

extern "C" int hk_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int hk_init(int device)
{
    if (device < 0) { set_error("hk_init: device must be >= 0"); return HK_ERR_INVALID_ARGUMENT; }
    if (g_device_state.load() == 1 && g_device.load() != device) { set_error("hk_init: already bound to device %d (one process per GPU)", g_device.load()); return HK_ERR_INVALID_ARGUMENT; }
    g_device.store(device);
    return ensure_device();
}

extern "C" void hk_shutdown(void)
{
    // Releases the CALLING thread's streams and scratch buffers and forgets the device binding; other threads' contexts are
    // released by their owners (when they end, or at their next call after a re-initialisation on another device).
    if (g_device_state.load() == 1) { cudaDeviceSynchronize(); thread_ctx().release(); }
    g_device_state.store(0);
    g_device.store(-1);
}

static int check_lqng_args(const char* who, int batch, int N, int horizon, const void* A, const void* B, const void* Q, const void* q,
                           const void* R, const void* x0, const void* u0)
{
    if (batch < 0) { set_error("%s: batch must be >= 0", who); return HK_ERR_INVALID_ARGUMENT; }
    if (N < 1 || N > HK_MAX_PLAYERS) { set_error("%s: n_players must be 1..%d (got %d)", who, HK_MAX_PLAYERS, N); return HK_ERR_INVALID_ARGUMENT; }
    if (horizon < 0 || horizon > HK_MAX_HORIZON) { set_error("%s: horizon must be 0..%d (got %d)", who, HK_MAX_HORIZON, horizon); return HK_ERR_INVALID_ARGUMENT; }
    if (batch > 0 && (!A || !B || !Q || !q || !R || !x0 || !u0)) { set_error("%s: null operand", who); return HK_ERR_INVALID_ARGUMENT; }
    return HK_OK;
}

extern "C" int hk_lqng_solve_batch_device(int batch, int n_players, int horizon, int time_varying,
                                          const double* dA, const double* dB, const double* dQ, const double* dq, const double* dR,
                                          const double* dx0, double* du0, double* dP, double* dalpha, double* dtraj, int* dstatus,
                                          void* cuda_stream)
{
    int rc = check_lqng_args("hk_lqng_solve_batch_device", batch, n_players, horizon, dA, dB, dQ, dq, dR, dx0, du0);
    if (rc) return rc;
    if (batch == 0) return HK_OK;
    ThreadCtx* c = ctx();
    if (!c) return HK_ERR_NO_DEVICE;
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
    const int T = horizon + 1, n = 4 * n_players, m = 2 * n_players;
    if (dtraj && (!dP || !dalpha)) {        // the rollout re-reads every step's gains: give it scratch when the caller keeps none
        double* scratch = (double*)dscratch(c, 1, sizeof(double) * (size_t)batch * T * (m * n + m));
        if (!scratch) return HK_ERR_OUT_OF_MEMORY;
        if (!dP) dP = scratch;
        if (!dalpha) dalpha = scratch + (size_t)batch * T * m * n;
    }
    return lqng_launch(batch, n_players, horizon, time_varying ? 1 : 0, dA, dB, dQ, dq, dR, dx0, du0, dP, dalpha, dtraj, dstatus, s);
}

// Host-pointer entry: chunked two-stream pipeline so that the H2D copy of chunk k+1, the solve of chunk k and the D2H
// copy of chunk k-1 overlap (PCIe is full duplex).  Pinned caller buffers get full PCIe rate; pageable ones are staged
// by the driver.
extern "C" int hk_lqng_solve_batch(int batch, int n_players, int horizon, int time_varying,
                                   const double* A, const double* B, const double* Q, const double* q, const double* R, const double* x0,
                                   double* u0, double* P, double* alpha, double* traj, int* status)
{
    int rc = check_lqng_args("hk_lqng_solve_batch", batch, n_players, horizon, A, B, Q, q, R, x0, u0);
    if (rc) return rc;
    if (batch == 0) return HK_OK;
    ThreadCtx* c = ctx();
    if (!c) return HK_ERR_NO_DEVICE;
    struct DrainOnError { ThreadCtx* c; bool ok = false; ~DrainOnError() { if (!ok) drain_ctx(c); } } drain_guard{c};   // an error return must not leave copies in flight on the caller's buffers
    const int N = n_players, T = horizon + 1, Tm = time_varying ? T : 1, n = 4 * N, m = 2 * N;
    const size_t eA = (size_t)Tm * N * 16, eB = (size_t)Tm * N * 8, eQ = (size_t)Tm * N * n * n, eq = (size_t)Tm * N * n, eR = (size_t)Tm * N * 4;
    const size_t in_elems = eA + eB + eQ + eq + eR + n;
    const bool needPa = P || alpha || traj;
    const size_t out_elems = (size_t)m + (needPa ? (size_t)T * (m * n + m) : 0) + (traj ? (size_t)(T + 1) * n : 0);
    if ((in_elems + out_elems) * sizeof(double) * (size_t)batch <= (256u << 10)) {
        // Small calls — the reference's own pattern is ONE solve per agent and physics step (HierarchicalKartAgent.cs:319-325) —
        // are latency-bound by the number of copies, not by bytes: gather the six operand arrays into one pinned staging
        // buffer, one H2D copy, one launch, one D2H copy of every requested output.
        const size_t in_b = in_elems * sizeof(double) * batch, out_b = out_elems * sizeof(double) * batch + sizeof(int) * batch;
        char* h = (char*)hscratch(c, 0, in_b + out_b);
        char* d = (char*)dscratch(c, 2, in_b + out_b + 64);
        if (!h || !d) return HK_ERR_OUT_OF_MEMORY;
        const double* src[6] = {A, B, Q, q, R, x0};
        const size_t per[6] = {eA, eB, eQ, eq, eR, (size_t)n};
        // A handful of problems: skip the copies altogether — pinned staging memory is device-accessible (UVA), the kernel reads
        // its operands and writes its results over PCIe directly (one launch + one synchronise).
        static const int zc_max = getenv("HK_ZEROCOPY_MAX") ? atoi(getenv("HK_ZEROCOPY_MAX")) : 4;
        const bool zero_copy = batch <= zc_max;
        if (zero_copy) d = h;
        double* hp = (double*)h;
        double* dp[6];
        double* dcur = (double*)d;
        for (int i = 0; i < 6; ++i) {
            std::memcpy(hp, src[i], per[i] * sizeof(double) * batch);
            dp[i] = dcur;
            hp += per[i] * batch; dcur += per[i] * batch;
        }
        double* du = dcur;
        double* dP = needPa ? du + (size_t)m * batch : nullptr;
        double* da = needPa ? dP + (size_t)T * m * n * batch : nullptr;
        double* dt = traj ? da + (size_t)T * m * batch : nullptr;
        double* dend = traj ? dt + (size_t)(T + 1) * n * batch : (needPa ? da + (size_t)T * m * batch : du + (size_t)m * batch);
        int* dst = (int*)dend;
        cudaStream_t s = c->stream;
        if (!zero_copy) HK_CUDA(cudaMemcpyAsync(d, h, in_b, cudaMemcpyHostToDevice, s));
        rc = lqng_launch(batch, N, horizon, time_varying ? 1 : 0, dp[0], dp[1], dp[2], dp[3], dp[4], dp[5], du, dP, da, dt, dst, s);
        if (rc) return rc;
        if (!zero_copy) HK_CUDA(cudaMemcpyAsync(h + in_b, du, out_b, cudaMemcpyDeviceToHost, s));
        HK_CUDA(cudaStreamSynchronize(s));
        const char* ho = h + in_b;
        std::memcpy(u0, ho, sizeof(double) * m * batch); ho += sizeof(double) * m * batch;
        if (needPa) {
            if (P) std::memcpy(P, ho, sizeof(double) * T * m * n * batch);
            ho += sizeof(double) * T * m * n * batch;
            if (alpha) std::memcpy(alpha, ho, sizeof(double) * T * m * batch);
            ho += sizeof(double) * T * m * batch;
        }
        if (traj) { std::memcpy(traj, ho, sizeof(double) * (T + 1) * n * batch); ho += sizeof(double) * (T + 1) * n * batch; }
        if (status) std::memcpy(status, ho, sizeof(int) * batch);
        drain_guard.ok = true;
        return HK_OK;
    }
    const int nchunks = batch >= 8192 ? 4 : 1;
    const int chunk = (batch + nchunks - 1) / nchunks;
    // two device buffers (double buffering) sized for one chunk each
    const size_t chunk_bytes = ((in_elems + out_elems) * sizeof(double) + sizeof(int)) * (size_t)chunk + 1024;
    char* dbufs[2];
    for (int k = 0; k < 2; ++k) {
        dbufs[k] = (char*)dscratch(c, 2 + k, nchunks > 1 || k == 0 ? chunk_bytes : 0);
        if (!dbufs[k] && (nchunks > 1 || k == 0)) return HK_ERR_OUT_OF_MEMORY;
    }
    cudaStream_t streams[2] = {c->stream, c->stream2};
    for (int ci = 0; ci < nchunks; ++ci) {
        const int b0 = ci * chunk, nb = (b0 + chunk <= batch) ? chunk : batch - b0;
        if (nb <= 0) break;
        cudaStream_t s = streams[ci & 1];
        double* d = (double*)dbufs[ci & 1];
        double *dA = d, *dB = dA + eA * nb, *dQ = dB + eB * nb, *dq = dQ + eQ * nb, *dR = dq + eq * nb, *dx = dR + eR * nb;
        double* du = dx + (size_t)n * nb;
        double* dP = needPa ? du + (size_t)m * nb : nullptr;
        double* da = needPa ? dP + (size_t)T * m * n * nb : nullptr;
        double* dt = traj ? da + (size_t)T * m * nb : nullptr;
        double* dend = traj ? dt + (size_t)(T + 1) * n * nb : (needPa ? da + (size_t)T * m * nb : du + (size_t)m * nb);
        int* dst = (int*)dend;
        HK_CUDA(cudaMemcpyAsync(dA, A + eA * b0, sizeof(double) * eA * nb, cudaMemcpyHostToDevice, s));
        HK_CUDA(cudaMemcpyAsync(dB, B + eB * b0, sizeof(double) * eB * nb, cudaMemcpyHostToDevice, s));
        HK_CUDA(cudaMemcpyAsync(dQ, Q + eQ * b0, sizeof(double) * eQ * nb, cudaMemcpyHostToDevice, s));
        HK_CUDA(cudaMemcpyAsync(dq, q + eq * b0, sizeof(double) * eq * nb, cudaMemcpyHostToDevice, s));
        HK_CUDA(cudaMemcpyAsync(dR, R + eR * b0, sizeof(double) * eR * nb, cudaMemcpyHostToDevice, s));
        HK_CUDA(cudaMemcpyAsync(dx, x0 + (size_t)n * b0, sizeof(double) * n * nb, cudaMemcpyHostToDevice, s));
        rc = lqng_launch(nb, N, horizon, time_varying ? 1 : 0, dA, dB, dQ, dq, dR, dx, du, dP, da, dt, dst, s);
        if (rc) return rc;
        HK_CUDA(cudaMemcpyAsync(u0 + (size_t)m * b0, du, sizeof(double) * m * nb, cudaMemcpyDeviceToHost, s));
        if (P) HK_CUDA(cudaMemcpyAsync(P + (size_t)T * m * n * b0, dP, sizeof(double) * T * m * n * nb, cudaMemcpyDeviceToHost, s));
        if (alpha) HK_CUDA(cudaMemcpyAsync(alpha + (size_t)T * m * b0, da, sizeof(double) * T * m * nb, cudaMemcpyDeviceToHost, s));
        if (traj) HK_CUDA(cudaMemcpyAsync(traj + (size_t)(T + 1) * n * b0, dt, sizeof(double) * (T + 1) * n * nb, cudaMemcpyDeviceToHost, s));
        if (status) HK_CUDA(cudaMemcpyAsync(status + b0, dst, sizeof(int) * nb, cudaMemcpyDeviceToHost, s));
        // before chunk ci+2 reuses this buffer its stream order already serialises (same stream)
    }
    HK_CUDA(cudaStreamSynchronize(c->stream));
    HK_CUDA(cudaStreamSynchronize(c->stream2));
    drain_guard.ok = true;
    return HK_OK;
}

extern "C" int hk_lqng_solve_one(int n_players, int horizon, const double* A, const double* B, const double* Q, const double* q,
                                 const double* R, const double* x0, double* u0)
{
    return hk_lqng_solve_batch(1, n_players, horizon, 0, A, B, Q, q, R, x0, u0, nullptr, nullptr, nullptr, nullptr);
}

// Ingest kernel of the zero-copy mode (HK_E2E_COPY_MODE=2, an experiment kept behind its knob): the caller's pinned arrays are
// device-accessible under UVA, so one kernel pulls a chunk of all seven over PCIe with 16-byte loads instead of seven copies.
struct IngestArgs { const double2* src[7]; double2* dst[7]; long long n2[7]; };
__global__ void lqng_ingest_kernel(IngestArgs a)
{
    const long long stride = (long long)gridDim.x * blockDim.x, id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 7; ++i)
        for (long long e = id; e < a.n2[i]; e += stride) a.dst[i][e] = a.src[i][e];
}

// Chunk schedule of the host-pointer pipelines.  The copies are the critical path (PCIe), so what a schedule can lose is (a) a fixed cost per copy
// and (b) the tail after the last copy — that chunk's solve and its D2H.  Equal chunks trade one against the other; a TAPERED schedule does not:
// every chunk is half the one before (the solve runs ~3x as fast as the copy, so a chunk's solve still hides behind the next, smaller copy)
// down to `floor_n` problems (measured optimum 8,192: 32,768 + 16,384 + 8,192 + 8,192 for the 65,536-problem call, 0.565 -> 0.53 ms together with a
// SINGLE copy stream — copies on several streams share the link and every chunk lands late).  Only for the 2-kart game: a 4-kart solve takes as long
// as its copy, so there equal chunks on four copy streams stay.  starts[k] .. starts[k + 1] is chunk k; sizes are even (16-byte aligned sub-arrays).
// HK_E2E_TAPER=0: equal chunks.
static int chunk_schedule(int batch, int nchunks_equal, int* starts, int max_chunks, int* largest, bool allow_taper)
{
    static const bool taper = !(getenv("HK_E2E_TAPER") && atoi(getenv("HK_E2E_TAPER")) == 0);
    static const int floor_env = getenv("HK_E2E_TAPER_FLOOR") ? atoi(getenv("HK_E2E_TAPER_FLOOR")) : 8192;
    static const double ratio = getenv("HK_E2E_TAPER_RATIO") ? atof(getenv("HK_E2E_TAPER_RATIO")) : 0.5;
    int n = 0;
    starts[0] = 0;
    if (taper && allow_taper && nchunks_equal > 2 && batch <= 65536 * 2) {
        const int floor_n = floor_env < 256 ? 256 : floor_env;
        int left = batch;
        while (left > 0 && n < max_chunks - 1) {
            int c = ((int)(left * ratio) + 1) & ~1;
            if (c < floor_n || left - c < floor_n) c = left;
            starts[n + 1] = starts[n] + c;
            left -= c;
            ++n;
        }
        if (left > 0) { starts[n + 1] = starts[n] + left; ++n; }
    } else {
        const int chunk = ((batch + nchunks_equal - 1) / nchunks_equal + 1) & ~1;
        for (int b0 = 0; b0 < batch && n < max_chunks; b0 += chunk) { starts[n + 1] = b0 + chunk < batch ? b0 + chunk : batch; ++n; }
    }
    int big = 0;
    for (int k = 0; k < n; ++k) big = starts[k + 1] - starts[k] > big ? starts[k + 1] - starts[k] : big;
    *largest = big;
    return n;
}

// Host-pointer entry of the compact description.  Pipeline over chunks of the batch: the seven H2D copies of a chunk go to
// dedicated copy streams (round robin) and never queue behind a solve or a D2H; the assembly + solve of chunk k (compute
// stream k & 1) waits on the chunk's "copied in" event, its results go back on the same compute stream.  Chunk buffers form a
// ring of four, a copy into a ring slot waits for the slot's previous tenant to have drained.
extern "C" int hk_lqng_assemble_solve_batch(int batch, int n_players, int horizon, double dt, const double* x0, const double* target,
                                            const double* tw, const double* cw, const double* aw, const double* otgt, const double* otw,
                                            double* u0, int* status)
{
    const int N = n_players;
    if (batch < 0 || N < 1 || N > HK_MAX_PLAYERS || horizon < 0 || horizon > HK_MAX_HORIZON) { set_error("hk_lqng_assemble_solve_batch: invalid dimensions"); return HK_ERR_INVALID_ARGUMENT; }
    if (batch > 0 && (!x0 || !target || !tw || !cw || !u0 || (N > 1 && (!aw || !otgt || !otw)))) { set_error("hk_lqng_assemble_solve_batch: null operand"); return HK_ERR_INVALID_ARGUMENT; }
    if (batch == 0) return HK_OK;
    ThreadCtx* c = ctx();
    if (!c) return HK_ERR_NO_DEVICE;
    struct DrainOnError { ThreadCtx* c; bool ok = false; ~DrainOnError() { if (!ok) drain_ctx(c); } } drain_guard{c};   // an error return must not leave copies in flight on the caller's buffers
    const int K = N - 1, m = 2 * N;
    const size_t e[7] = {(size_t)N * 4, (size_t)N * 4, (size_t)N * 4, (size_t)N, (size_t)N * K * 2, (size_t)N * K * 4, (size_t)N * K * 3};
    const double* src[7] = {x0, target, tw, cw, aw, otgt, otw};
    size_t per = 0;
    for (int i = 0; i < 7; ++i) per += e[i];
    static const int chunks_env = getenv("HK_E2E_CHUNKS") ? atoi(getenv("HK_E2E_CHUNKS")) : 0;    // tuning knobs
    static const int ncs_env = getenv("HK_E2E_COPY_STREAMS") ? atoi(getenv("HK_E2E_COPY_STREAMS")) : 0;     // 0: one stream for the 2-kart game, four otherwise
    // 0: seven cudaMemcpyAsync per chunk; 1 (default): one cudaMemcpyBatchAsync per chunk — the pipeline is bound by the host's
    // issue rate (~3 us per call) before it is bound by PCIe; 2: ingest kernel (pinned sources only)
    static const int mode_env = getenv("HK_E2E_COPY_MODE") ? atoi(getenv("HK_E2E_COPY_MODE")) : 1;
    static std::atomic<int> batch_copy_ok{1};
    const int mode = (mode_env == 1 && !batch_copy_ok.load()) ? 0 : mode_env;
    constexpr int RING = 4;
    const int ncs = ncs_env < 1 ? (N <= 2 ? 1 : 4) : ncs_env > 4 ? 4 : ncs_env;
    int nchunks = batch >= 16384 ? (chunks_env > 0 ? chunks_env : (batch >= 65536 ? 8 : batch / 8192)) : 1;
    const int by_size = (int)(((long long)batch + 65535) / 65536);                                // chunks of at most 65,536 problems
    if (nchunks < by_size) nchunks = by_size;
    int starts[40], chunk = 0;                                                                    // chunk: the largest one (slot layout)
    nchunks = chunk_schedule(batch, nchunks, starts, 39, &chunk, N <= 2);
    const size_t slot_bytes = (((per + m) * sizeof(double) + sizeof(int)) * (size_t)chunk + 255) & ~(size_t)255;
    const int ring = nchunks < RING ? nchunks : RING;
    char* dring = (char*)dscratch(c, 4, slot_bytes * ring);
    if (!dring) return HK_ERR_OUT_OF_MEMORY;
    static const int trig_slot[RING] = {5, 7, 10, 11};
    {   // grow the per-slot scratch of lqng_assemble_launch before anything is in flight ((cos, sin) pairs, or dense records for N != 2)
        const size_t n = 4 * (size_t)N, dense = (size_t)N * 16 + N * 8 + N * n * n + N * n + N * 4 + n;
        for (int r = 0; r < ring; ++r)
            if (!dscratch(c, trig_slot[r], sizeof(double) * (N == 2 ? 4 : dense) * (size_t)chunk)) return HK_ERR_OUT_OF_MEMORY;
    }
    cudaStream_t compute[2] = {c->stream, c->stream2};
    for (int ci = 0; ci < nchunks; ++ci) {
        const int b0 = starts[ci], nb = starts[ci + 1] - starts[ci];
        if (nb <= 0) break;
        const int r = ci % ring;
        cudaStream_t cs = nchunks > 1 ? c->cstream[ci % ncs] : compute[0];
        cudaStream_t s = compute[ci & 1];
        double* dp[7];
        double* cur = (double*)(dring + slot_bytes * r);
        if (ci >= ring) HK_CUDA(cudaStreamWaitEvent(cs, c->pev[8 + r], 0));                       // the slot's previous tenant has drained
        void* dsts[7]; void* srcs[7]; size_t sizes[7]; size_t cnt = 0;
        IngestArgs ia;
        bool even = true;
        for (int i = 0; i < 7; ++i) even = even && ((e[i] * nb) & 1) == 0;
        const int md = (mode == 2 && !even) ? 0 : mode;
        for (int i = 0; i < 7; ++i) {
            dp[i] = cur;
            ia.src[i] = (const double2*)(src[i] + e[i] * b0); ia.dst[i] = (double2*)cur; ia.n2[i] = (long long)(e[i] * nb + 1) / 2;
            if (e[i]) {
                if (md == 0) HK_CUDA(cudaMemcpyAsync(cur, src[i] + e[i] * b0, sizeof(double) * e[i] * nb, cudaMemcpyHostToDevice, cs));
                else { dsts[cnt] = cur; srcs[cnt] = (void*)(src[i] + e[i] * b0); sizes[cnt] = sizeof(double) * e[i] * nb; ++cnt; }
            }
            cur += e[i] * chunk;
        }
        if (md == 1 && cnt) {
            cudaMemcpyAttributes at = {};
            at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
            size_t idx = 0, fail = 0;
            if (cudaMemcpyBatchAsync(dsts, srcs, sizes, cnt, &at, &idx, 1, &fail, cs) != cudaSuccess) {
                cudaGetLastError();                                                               // driver without batched copies: plain copies from now on
                batch_copy_ok.store(0);
                for (size_t k = 0; k < cnt; ++k) HK_CUDA(cudaMemcpyAsync(dsts[k], srcs[k], sizes[k], cudaMemcpyHostToDevice, cs));
            }
        } else if (md == 2) {
            count_launch(); lqng_ingest_kernel<<<296, 512, 0, cs>>>(ia);
            HK_CUDA(cudaGetLastError());
        }
        double* du = cur;
        int* dst = (int*)(du + (size_t)m * chunk);
        if (nchunks > 1) {
            HK_CUDA(cudaEventRecord(c->pev[r], cs));
            HK_CUDA(cudaStreamWaitEvent(s, c->pev[r], 0));
        }
        int rc = lqng_assemble_launch(nb, N, horizon, dt, dp[0], dp[1], dp[2], dp[3], dp[4], dp[5], dp[6], du, dst, s, trig_slot[r]);
        if (rc) return rc;
        HK_CUDA(cudaMemcpyAsync(u0 + (size_t)m * b0, du, sizeof(double) * m * nb, cudaMemcpyDeviceToHost, s));
        if (status) HK_CUDA(cudaMemcpyAsync(status + b0, dst, sizeof(int) * nb, cudaMemcpyDeviceToHost, s));
        if (ci + ring < nchunks) HK_CUDA(cudaEventRecord(c->pev[8 + r], s));
    }
    HK_CUDA(cudaStreamSynchronize(c->stream));
    HK_CUDA(cudaStreamSynchronize(c->stream2));
    drain_guard.ok = true;
    return HK_OK;
}

// The same call with the seven arrays interleaved per problem (include/hk_abi.h): ONE host-to-device copy per chunk instead of seven — the
// copy engine's per-copy cost, not the host's issue rate, was what kept the seven-array pipeline above its PCIe floor — and one TMA bulk copy
// per problem in the 2-kart solve kernel instead of seven.  Same chunk pipeline: copies on dedicated streams, assembly + solve + D2H of chunk
// k on compute stream k & 1 behind the chunk's "copied in" event, a ring of four chunk buffers.
extern "C" int hk_lqng_assemble_solve_packed(int batch, int n_players, int horizon, double dt, const double* records, double* u0, int* status)
{
    const int N = n_players;
    if (batch < 0 || N < 1 || N > HK_MAX_PLAYERS || horizon < 0 || horizon > HK_MAX_HORIZON) { set_error("hk_lqng_assemble_solve_packed: invalid dimensions"); return HK_ERR_INVALID_ARGUMENT; }
    if (batch > 0 && (!records || !u0)) { set_error("hk_lqng_assemble_solve_packed: null operand"); return HK_ERR_INVALID_ARGUMENT; }
    if (batch == 0) return HK_OK;
    ThreadCtx* c = ctx();
    if (!c) return HK_ERR_NO_DEVICE;
    struct DrainOnError { ThreadCtx* c; bool ok = false; ~DrainOnError() { if (!ok) drain_ctx(c); } } drain_guard{c};
    const int K = N - 1, m = 2 * N, P = 13 * N + 9 * N * K;
    // Zero-copy form (2-kart game, records / u0 / status all in pinned host memory, which is device-accessible under UVA): ONE launch — the solve
    // kernel's bulk copies pull each 352-byte record over PCIe exactly once while other warps solve, results are written straight back.  No
    // staging copies, no chunk pipeline, no (cos h, sin h) pre-pass.  Measured (round 2) and left OFF: 0.545-0.554 ms per 65,536-problem call against
    // 0.512-0.516 ms for the chunk pipeline — 352-byte reads over PCIe reach ~42 GB/s where the copy engine's large transfers reach 54.
    static const int zc_env = getenv("HK_E2E_ZEROCOPY") ? atoi(getenv("HK_E2E_ZEROCOPY")) : 0;
    if (zc_env && N == 2 && (reinterpret_cast<uintptr_t>(records) & 15) == 0) {
        auto mapped = [](const void* ptr) -> void* {
            cudaPointerAttributes a{};
            if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return nullptr; }
            return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
        };
        void* drec = mapped(records);
        void* du = mapped(u0);
        void* dst = status ? mapped(status) : nullptr;
        if (drec && du && (!status || dst)) {
            int rc = lqng_solve_packed_in_place(batch, horizon, dt, (const double*)drec, (double*)du, (int*)dst, c->stream);
            if (rc) return rc;
            HK_CUDA(cudaStreamSynchronize(c->stream));
            drain_guard.ok = true;
            return HK_OK;
        }
    }
    static const int chunks_env = getenv("HK_E2E_CHUNKS") ? atoi(getenv("HK_E2E_CHUNKS")) : 0;
    constexpr int RING = 4;
    int nchunks = batch >= 16384 ? (chunks_env > 0 ? chunks_env : (batch >= 65536 ? 8 : batch / 8192)) : 1;
    const int by_size = (int)(((long long)batch + 65535) / 65536);
    if (nchunks < by_size) nchunks = by_size;
    static const int ncs_env_p = getenv("HK_E2E_COPY_STREAMS") ? atoi(getenv("HK_E2E_COPY_STREAMS")) : 0;
    const int ncs_p = ncs_env_p < 1 ? (N <= 2 ? 1 : 4) : ncs_env_p > 4 ? 4 : ncs_env_p;
    int starts[40], chunk = 0;
    nchunks = chunk_schedule(batch, nchunks, starts, 39, &chunk, N <= 2);
    const size_t slot_bytes = ((((size_t)P + m) * sizeof(double) + sizeof(int)) * (size_t)chunk + 255) & ~(size_t)255;
    const int ring = nchunks < RING ? nchunks : RING;
    char* dring = (char*)dscratch(c, 4, slot_bytes * ring);
    if (!dring) return HK_ERR_OUT_OF_MEMORY;
    static const int trig_slot[RING] = {5, 7, 10, 11};
    {
        const size_t n = 4 * (size_t)N, dense = (size_t)N * 16 + N * 8 + N * n * n + N * n + N * 4 + n;
        for (int r = 0; r < ring; ++r)
            if (!dscratch(c, trig_slot[r], sizeof(double) * (N == 2 ? 4 : dense + P) * (size_t)chunk)) return HK_ERR_OUT_OF_MEMORY;
    }
    cudaStream_t compute[2] = {c->stream, c->stream2};
    for (int ci = 0; ci < nchunks; ++ci) {
        const int b0 = starts[ci], nb = starts[ci + 1] - starts[ci];
        if (nb <= 0) break;
        const int r = ci % ring;
        cudaStream_t cs = nchunks > 1 ? c->cstream[ci % ncs_p] : compute[0];
        cudaStream_t s = compute[ci & 1];
        double* drec = (double*)(dring + slot_bytes * r);
        double* du = drec + (size_t)P * chunk;
        int* dst = (int*)(du + (size_t)m * chunk);
        if (ci >= ring) HK_CUDA(cudaStreamWaitEvent(cs, c->pev[8 + r], 0));                       // the slot's previous tenant has drained
        // one plain copy per chunk; cudaMemcpyBatchAsync for it (whole or split in four) measured the same (0.550 / 0.558 / 0.553 ms per call)
        HK_CUDA(cudaMemcpyAsync(drec, records + (size_t)P * b0, sizeof(double) * P * nb, cudaMemcpyHostToDevice, cs));
        if (nchunks > 1) {
            HK_CUDA(cudaEventRecord(c->pev[r], cs));
            HK_CUDA(cudaStreamWaitEvent(s, c->pev[r], 0));
        }
        int rc = lqng_assemble_launch_packed(nb, N, horizon, dt, drec, du, dst, s, trig_slot[r]);
        if (rc) return rc;
        HK_CUDA(cudaMemcpyAsync(u0 + (size_t)m * b0, du, sizeof(double) * m * nb, cudaMemcpyDeviceToHost, s));
        if (status) HK_CUDA(cudaMemcpyAsync(status + b0, dst, sizeof(int) * nb, cudaMemcpyDeviceToHost, s));
        if (ci + ring < nchunks) HK_CUDA(cudaEventRecord(c->pev[8 + r], s));
    }
    HK_CUDA(cudaStreamSynchronize(c->stream));
    HK_CUDA(cudaStreamSynchronize(c->stream2));
    drain_guard.ok = true;
    return HK_OK;
}
