// mcalf_api.cu -- the C ABI of libmcalf_b200.so (include/mcalf_b200.h): context construction from
// the als_fitter state, batch entry points, the pipelined host-pointer path.  No computation happens
// on the host here beyond the one-off per-pixel set-up of mcalf_create (the analogue of
// als_fitter.__init__, hires_fitter.py:65-200); without a CUDA device every entry point fails.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>

#include "mcalf_device.h"
#include "host_setup.h"

using namespace mcalf;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(MCALF_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

// Every entry point runs on the context's device and leaves the calling thread's current device as it found it.
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); cur = -1; }
        if (cur != dev) {
            err = cudaSetDevice(dev);
            prev = cur;
        }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ON_DEVICE(dev)                                                                                  \
    DeviceGuard guard_(dev);                                                                            \
    if (guard_.err != cudaSuccess) return fail(MCALF_E_CUDA, "cudaSetDevice(%d): %s", dev, cudaGetErrorString(guard_.err))

constexpr int NBUF = 2;
constexpr int NRING = 3;
constexpr size_t ZC_PARAM_BYTES = 256u << 10;  // zero-copy path: at most 256 KB of parameters ...
constexpr long long ZC_MAX_ROWS = 1024;       // ... and 1024 rows
constexpr double FWHM_TO_SIGMA_H = 2.354820;   // hires_fitter.py:454
constexpr double TRUNC_SIGMAS_H = 3.0348;      // hires_fitter.py:458
constexpr double A_MAX_LIMIT = 0.02;      // beyond this the a^4 term of the core series is > 1e-6
constexpr double A_MAX_DEFAULT = 0.01;    // default hand-over to the fp64 kernel: fp32 forms good to 3e-7 below it
constexpr int FF_NC_HOST = 8;          // far-field coefficients per chunk (FF_DEG + 1 in voigt_math.cuh)

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr, k0 = nullptr, k1 = nullptr;
    double *h_params = nullptr, *d_params = nullptr;   // pinned / device, slice * ld_cap doubles
    double *h_out = nullptr, *d_out = nullptr;         // 2 * slice doubles (logL, chi2)
    void *h_flux = nullptr, *d_flux = nullptr;
    size_t flux_cap = 0, params_cap = 0, hparams_cap = 0, out_cap = 0;
    unsigned int *counters = nullptr;                  // [2][2]: {work counter, fallback count} x launch parity
    int parity = 0;
    BatchArgs pending{};                               // the last deferred fix-up (zero-copy path)
    unsigned int *pending_count = nullptr;
    int *fallback = nullptr;                           // [fallback_cap]
    size_t fallback_cap = 0;
};

}  // namespace

struct mcalf_ctx {
    int device = 0, sm_count = 0;
    DevProblem P{};
    std::vector<void *> allocs;
    Slot slot[NBUF];
    Slot dev_slot;                    // counters / fallback list of the MCALF_F_ON_DEVICE path
    Slot ring[NRING];                 // the copy-stream / compute-stream pipeline of large logL calls
    cudaStream_t copy_stream = nullptr, comp_stream = nullptr;
    cudaEvent_t ring_h2d[NRING] = {nullptr, nullptr, nullptr};
    double *pipe_dout = nullptr, *pipe_hout = nullptr;
    double *zc_buf = nullptr;         // mapped pinned memory for small host calls: params in, results out, no copies
    Slot zc_slot;
    size_t pipe_dout_cap = 0, pipe_hout_cap = 0;
    unsigned long long *d_stats = nullptr;
    int threads = 0, threads_small = 0, ctas_per_sm = 0, threads_opt = 0, ctas_opt = 0, dense = 0, dense_opt = -1;
    size_t smem_fast = 0, smem_fp64 = 0;
    long long slice = 65536;
    int collect_stats = 0;
    uint64_t kernel_launches = 0, samples = 0, samples_fp64 = 0;
    int last_slot = -1, last_ring = -1;
    // one call at a time per context: the slots' counters, staging buffers and events are not re-entrant
    std::atomic_flag busy = ATOMIC_FLAG_INIT;
    // MCALF_F_ON_DEVICE calls may arrive on different streams: each one is ordered after the previous
    cudaStream_t dev_last_stream = nullptr;
    bool dev_has_last = false;
    cudaEvent_t dev_order = nullptr;
    // staging of the small utility calls (prior transform from host pointers)
    cudaStream_t util_stream = nullptr;
    double *util_dev = nullptr;
    size_t util_cap = 0;
    // samplers hand over the same buffers call after call: the (slow, driver-locked) pointer query is remembered
    struct PinEntry { const void *p = nullptr; bool pinned = false; } pin_cache[8];
    int pin_next = 0;
    bool is_pinned(const void *p);
};

namespace {
bool query_pinned(const void *p);
}
bool mcalf_ctx::is_pinned(const void *p) {
    if (!p) return false;
    for (const PinEntry &e : pin_cache)
        if (e.p == p) return e.pinned;
    PinEntry &slot = pin_cache[pin_next];
    pin_next = (pin_next + 1) % 8;
    slot.p = p;
    slot.pinned = query_pinned(p);
    return slot.pinned;
}

namespace {
struct BusyGuard {
    mcalf_ctx *c;
    bool mine;
    explicit BusyGuard(mcalf_ctx *ctx) : c(ctx), mine(!ctx->busy.test_and_set(std::memory_order_acquire)) {}
    ~BusyGuard() { if (mine) c->busy.clear(std::memory_order_release); }
};
}  // namespace
#define ONE_CALL(ctx)                                                                                          \
    BusyGuard busy_(ctx);                                                                                      \
    if (!busy_.mine) return fail(MCALF_E_INVALID, "context busy: another call on this context is in flight (use one context per thread)")

namespace {

template <typename T>
int upload(mcalf_ctx *c, const std::vector<T> &v, const T **out) {
    void *d = nullptr;
    const size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
    CU(cudaMalloc(&d, bytes));
    c->allocs.push_back(d);
    if (!v.empty()) CU(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T *)d;
    return MCALF_OK;
}

int choose_launch(mcalf_ctx *c) {
    const DevProblem &P = c->P;
    int nwarps = c->threads_opt > 0 ? c->threads_opt / 32 : std::min(std::max((P.nchunks + 1) / 2, 2), 8);   // measured: about two chunks per warp balances best
    if (nwarps < 1) nwarps = 1;
    if (nwarps > 32) nwarps = 32;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, c->device));
    c->P.lay = fast_smem_layout(P);
    size_t smem = (size_t)c->P.lay.bytes;
    size_t static_smem = 0;            // the kernel's fixed-address H1 table
    // (also raises every kernel's dynamic shared-memory limit to what the device allows: the attribute is
    // held per function and device, not per context, so it is never set to one problem's size)
    CU(configure_kernels((size_t)prop.sharedMemPerBlockOptin, &static_smem));
    while (smem + static_smem > (size_t)prop.sharedMemPerBlockOptin && nwarps > 1) {
        nwarps = (nwarps + 1) / 2;
    }
    if (smem + static_smem > (size_t)prop.sharedMemPerBlockOptin)
        return fail(MCALF_E_RESOURCE, "problem needs %zu B of shared memory per CTA (limit %zu): too many pixels/lines", smem,
                    (size_t)prop.sharedMemPerBlockOptin);
    c->smem_fp64 = fp64_smem_bytes(P);
    if (c->smem_fp64 > (size_t)prop.sharedMemPerBlockOptin)
        return fail(MCALF_E_RESOURCE, "fp64 kernel needs %zu B of shared memory per CTA (limit %zu)", c->smem_fp64,
                    (size_t)prop.sharedMemPerBlockOptin);
    c->threads = nwarps * 32;
    c->smem_fast = smem;
    int occ = 0, occ_dense = 0;
    CU(fast_occupancy(c->threads, c->smem_fast, 0, &occ));   // accounts for the kernel's static shared memory too
    if (c->threads <= 256 && c->dense_opt != 0) CU(fast_occupancy(c->threads, c->smem_fast, 1, &occ_dense));
    // the 48-register build is ~4 % slower per warp: it pays only where it seats more CTAs AND the samples are long
    // (measured: cfg 4, 8192 px, 5 instead of 4 CTAs: +3 %; cfg 2, 1998 px, 10 instead of 8 CTAs: -3 %)
    c->dense = (occ_dense > occ && (c->dense_opt == 1 || P.nchunks >= 16)) ? 1 : 0;
    if (c->dense) occ = occ_dense;
    if (occ < 1) return fail(MCALF_E_RESOURCE, "fp32 kernel does not fit an SM (threads %d, smem %zu)", c->threads, smem);
    c->ctas_per_sm = c->ctas_opt > 0 ? std::min(c->ctas_opt, occ) : occ;
    c->threads_small = 0;
    {
        const int ts = 32 * std::min(std::max(P.nchunks, 1), 32);
        int occ_s = 0;
        if (ts > c->threads && fast_occupancy(ts, c->smem_fast, 0, &occ_s) == cudaSuccess && occ_s >= 1) c->threads_small = ts;
    }
    return MCALF_OK;
}

// stage_in / stage_out: pinned staging buffers are only needed when the caller's buffers are pageable
int ensure_slot(mcalf_ctx *c, Slot &s, long long n, long long ld, size_t flux_bytes, bool host_io, bool stage_in = true,
                bool stage_out = true) {
    if (!s.stream) {
        CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        CU(cudaEventCreate(&s.k0));
        CU(cudaEventCreate(&s.k1));
        CU(cudaMalloc((void **)&s.counters, 4 * sizeof(unsigned int)));
        CU(cudaMemsetAsync(s.counters, 0, 4 * sizeof(unsigned int), s.stream));
        CU(cudaStreamSynchronize(s.stream));      // done before the first launch, whatever stream that is on
    }
    if ((size_t)n > s.fallback_cap) {
        if (s.fallback) CU(cudaFree(s.fallback));
        CU(cudaMalloc((void **)&s.fallback, sizeof(int) * (size_t)n));
        s.fallback_cap = (size_t)n;
    }
    if (!host_io) return MCALF_OK;
    const size_t pbytes = sizeof(double) * (size_t)n * (size_t)ld;
    if (pbytes > s.params_cap) {
        if (s.d_params) CU(cudaFree(s.d_params));
        CU(cudaMalloc((void **)&s.d_params, pbytes));
        s.params_cap = pbytes;
    }
    if (stage_in && pbytes > s.hparams_cap) {
        if (s.h_params) CU(cudaFreeHost(s.h_params));
        CU(cudaMallocHost((void **)&s.h_params, pbytes));
        s.hparams_cap = pbytes;
    }
    const size_t obytes = sizeof(double) * 2 * (size_t)n;
    if (stage_out && obytes > s.out_cap) {
        if (s.h_out) CU(cudaFreeHost(s.h_out));
        if (s.d_out) CU(cudaFree(s.d_out));
        CU(cudaMallocHost((void **)&s.h_out, obytes));
        CU(cudaMalloc((void **)&s.d_out, obytes));
        s.out_cap = obytes;
    }
    if (flux_bytes > s.flux_cap) {
        if (s.h_flux) CU(cudaFreeHost(s.h_flux));
        if (s.d_flux) CU(cudaFree(s.d_flux));
        CU(cudaMallocHost(&s.h_flux, flux_bytes));
        CU(cudaMalloc(&s.d_flux, flux_bytes));
        s.flux_cap = flux_bytes;
    }
    return MCALF_OK;
}

// true when `p` is page-locked host memory the copy engines can reach directly
bool query_pinned(const void *p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

// enqueue the kernels of one slice on `st`; all pointers are device pointers
// fallback_flag (mapped host memory, nullable): when given, the fp64 fix-up kernel is NOT launched here; the
// caller synchronises, looks at the flag and calls enqueue_fixup only if a sample was re-routed
int enqueue(mcalf_ctx *c, Slot &s, cudaStream_t st, const double *d_params, long long n, long long ld, uint32_t flags,
            double *d_logl, double *d_chi2, void *d_flux, int *fallback_flag = nullptr, double *const *peers = nullptr,
            int npeers = 0) {
    BatchArgs a{};
    a.npeers = npeers;
    for (int p = 0; p < npeers; ++p) a.logl_peer[p] = peers[p];
    a.params = d_params;
    a.B = n;
    a.ld = ld;
    a.flags = flags;
    a.logl_out = d_logl;
    a.chi2_out = d_chi2;
    a.flux_out = d_flux;
    // the fast kernel of launch k clears the counter pair launch k+1 will use: no memset per call
    unsigned int *cnt = s.counters + 2 * s.parity;
    a.work_counter = cnt;
    a.fallback_count = cnt + 1;
    a.clear_counters = s.counters + 2 * (s.parity ^ 1);
    a.fallback_list = s.fallback;
    a.fallback_flag = fallback_flag;
    a.stats = c->collect_stats ? c->d_stats : nullptr;
    const bool timed = c->collect_stats || &s == &c->dev_slot;   // last_kernel_ms; host-pointer calls skip the two records
    if (timed) CU(cudaEventRecord(s.k0, st));
    const int fp64_grid = (int)std::min<long long>(n, (long long)c->sm_count * 8);
    if (flags & MCALF_F_FP64) {
        CU(launch_fp64(c->P, a, nullptr, nullptr, fp64_grid, c->smem_fp64, st));
        c->kernel_launches += 1;
        c->samples_fp64 += (uint64_t)n;
    } else {
        int grid = (int)std::min<long long>(n, (long long)c->sm_count * c->ctas_per_sm);
        int threads = c->threads;
        // fewer samples than SMs: one CTA per sample, as many warps as the sample has chunks (results do
        // not depend on the CTA size)
        if (n <= c->sm_count && c->threads_opt == 0 && c->threads_small > threads) threads = c->threads_small;
        CU(launch_fast(c->P, a, grid, threads, c->smem_fast, threads == c->threads ? c->dense : 0, st));
        // samples outside the fp32 domain were listed by the fast kernel; the fp64 kernel finishes them
        // (exits at once when the list is empty)
        c->kernel_launches += 1;
        s.parity ^= 1;
        if (!fallback_flag) {
            const int fgrid = (int)std::min<long long>(n, (long long)c->sm_count * 2);
            CU(launch_fp64(c->P, a, s.fallback, cnt + 1, fgrid, c->smem_fp64, st));
            c->kernel_launches += 1;
        } else {
            s.pending = a;               // for enqueue_fixup
            s.pending_count = cnt + 1;
        }
    }
    if (timed) CU(cudaEventRecord(s.k1, st));
    c->samples += (uint64_t)n;
    return MCALF_OK;
}

int run_batch_inner(mcalf_ctx *c, const double *params, long long B, long long ld, uint32_t flags, void *stream, double *logl,
                    double *chi2, void *flux, double *const *peers, int npeers);

int run_batch(mcalf_ctx *c, const double *params, long long B, long long ld, uint32_t flags, void *stream, double *logl,
              double *chi2, void *flux, double *const *peers = nullptr, int npeers = 0) {
    const int rc = run_batch_inner(c, params, B, ld, flags, stream, logl, chi2, flux, peers, npeers);
#if defined(MCALF_CHECK)
    // bounds-checked build: every batch call ends with a device synchronisation and a look at the violation mask
    if (rc == MCALF_OK && c) {
        ON_DEVICE(c->device);
        unsigned int mask = 0u;
        CU(check_flag_fetch(&mask));
        if (mask) return fail(MCALF_E_CUDA, "bounds check failed in the fp32 kernel (mask 0x%x; bits in voigt_math.cuh)", mask);
    }
#endif
    return rc;
}

int run_batch_inner(mcalf_ctx *c, const double *params, long long B, long long ld, uint32_t flags, void *stream, double *logl,
                    double *chi2, void *flux, double *const *peers, int npeers) {
    if (!c) return fail(MCALF_E_INVALID, "null context");
    if (B < 0) return fail(MCALF_E_INVALID, "negative batch size");
    if (B == 0) return MCALF_OK;
    if (B > (1LL << 30)) return fail(MCALF_E_INVALID, "batch of %lld rows: split calls above 2^30 rows", B);
    if (!params) return fail(MCALF_E_INVALID, "null params");
    const uint32_t row5 = MCALF_F_ONECOMP | MCALF_F_ONECOMP_FILL | MCALF_F_ONELINE;
    const int need = (flags & MCALF_F_ONELINE) ? 6 : (flags & row5) ? 5 : c->P.ndim;
    if (ld < need) return fail(MCALF_E_INVALID, "ld (%lld) smaller than the row length (%d)", ld, need);
    if (((flags & row5) & ((flags & row5) - 1)) != 0) return fail(MCALF_E_INVALID, "ONECOMP, ONECOMP_FILL and ONELINE are exclusive");
    if ((flags & row5) && (flags & MCALF_F_UNIT_CUBE))
        return fail(MCALF_E_INVALID, "ONECOMP / ONELINE rows are physical parameters, not unit-cube draws");
    ON_DEVICE(c->device);
    ONE_CALL(c);
    const size_t esize = (flags & MCALF_F_FLUX_F64) ? 8 : 4;

    if (flags & MCALF_F_ON_DEVICE) {
        Slot &s = c->dev_slot;
        int rc = ensure_slot(c, s, B, ld, 0, false);
        if (rc) return rc;
        c->last_slot = NBUF;             // = dev_slot
        c->last_ring = -1;
        // the slot's counter pairs are handed from launch to launch: a call on another stream than the
        // previous one waits for it (same stream: stream order already does)
        cudaStream_t st = (cudaStream_t)stream;
        if (!c->dev_order) CU(cudaEventCreateWithFlags(&c->dev_order, cudaEventDisableTiming));
        if (c->dev_has_last && c->dev_last_stream != st) {
            CU(cudaEventRecord(c->dev_order, c->dev_last_stream));
            CU(cudaStreamWaitEvent(st, c->dev_order, 0));
        }
        c->dev_last_stream = st;
        c->dev_has_last = true;
        return enqueue(c, s, st, params, B, ld, flags, logl, chi2, flux, nullptr, peers, npeers);
    }

    // Small calls (the scalar callbacks of a CPU sampler, one live set): the kernel reads the parameters
    // from, and writes the results to, mapped pinned host memory -- no copy operations, one stream sync.
    if (!flux && (size_t)B * (size_t)ld * sizeof(double) <= ZC_PARAM_BYTES && B <= ZC_MAX_ROWS) {
        Slot &s = c->zc_slot;
        int rc = ensure_slot(c, s, ZC_MAX_ROWS, ld, 0, false);
        if (rc) return rc;
        if (!c->zc_buf) CU(cudaHostAlloc((void **)&c->zc_buf, ZC_PARAM_BYTES + 2 * ZC_MAX_ROWS * sizeof(double) + 64, cudaHostAllocMapped));
        double *zp = c->zc_buf, *zo = c->zc_buf + ZC_PARAM_BYTES / sizeof(double);
        int *zflag = (int *)(zo + 2 * ZC_MAX_ROWS);
        *zflag = 0;
        memcpy(zp, params, sizeof(double) * (size_t)B * (size_t)ld);
        c->last_slot = NBUF + 1;         // = zc_slot
        c->last_ring = -1;
        rc = enqueue(c, s, s.stream, zp, B, ld, flags, logl ? zo : nullptr, chi2 ? zo + ZC_MAX_ROWS : nullptr, nullptr,
                     (flags & MCALF_F_FP64) ? nullptr : zflag);
        if (rc) return rc;
        CU(cudaStreamSynchronize(s.stream));
        if (*zflag) {                    // rare: some sample lies outside the fp32 kernel's domain
            const int fgrid = (int)std::min<long long>(B, (long long)c->sm_count * 2);
            CU(launch_fp64(c->P, s.pending, s.fallback, s.pending_count, fgrid, c->smem_fp64, s.stream));
            c->kernel_launches += 1;
            CU(cudaStreamSynchronize(s.stream));
        }
        if (logl) memcpy(logl, zo, sizeof(double) * (size_t)B);
        if (chi2) memcpy(chi2, zo + ZC_MAX_ROWS, sizeof(double) * (size_t)B);
        return MCALF_OK;
    }

    // Large logL / chi2 calls: one copy stream feeds a ring of device slices, one compute stream runs the
    // kernels back to back; only the first slice's H2D is exposed, and the results come back in one copy.
    if (!flux && B > std::min<long long>(c->slice, 2048)) {
        const long long slice = c->slice;
        const bool pin_in = c->is_pinned(params), pin_logl = c->is_pinned(logl), pin_chi2 = c->is_pinned(chi2);
        if (!c->copy_stream) {
            CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
            CU(cudaStreamCreateWithFlags(&c->comp_stream, cudaStreamNonBlocking));
            for (int r = 0; r < NRING; ++r) CU(cudaEventCreateWithFlags(&c->ring_h2d[r], cudaEventDisableTiming));
        }
        const size_t obytes = sizeof(double) * 2 * (size_t)B;
        if (obytes > c->pipe_dout_cap) {
            if (c->pipe_dout) CU(cudaFree(c->pipe_dout));
            CU(cudaMalloc((void **)&c->pipe_dout, obytes));
            c->pipe_dout_cap = obytes;
        }
        if (((logl && !pin_logl) || (chi2 && !pin_chi2)) && obytes > c->pipe_hout_cap) {
            if (c->pipe_hout) CU(cudaFreeHost(c->pipe_hout));
            CU(cudaMallocHost((void **)&c->pipe_hout, obytes));
            c->pipe_hout_cap = obytes;
        }
        double *d_logl = c->pipe_dout, *d_chi2 = c->pipe_dout + B;
        long long k = 0, n = 0;
        for (long long off = 0; off < B; off += n, ++k) {
            // ramp the first slices up (1024, 2048, ...): only the first staging copy + H2D is exposed, so keep it short
            n = std::min(std::min(slice, (long long)1024 << std::min<long long>(k, 10)), B - off);
            const int r = (int)(k % NRING);
            Slot &s = c->ring[r];
            int rc = ensure_slot(c, s, slice, ld, 0, true, !pin_in, false);
            if (rc) return rc;
            if (k >= NRING) CU(cudaEventSynchronize(s.done));          // the ring slot's previous slice has been consumed
            const double *src = params + off * ld;
            if (!pin_in) {
                memcpy(s.h_params, src, sizeof(double) * (size_t)n * (size_t)ld);
                src = s.h_params;
            }
            CU(cudaMemcpyAsync(s.d_params, src, sizeof(double) * (size_t)n * (size_t)ld, cudaMemcpyHostToDevice, c->copy_stream));
            CU(cudaEventRecord(c->ring_h2d[r], c->copy_stream));
            CU(cudaStreamWaitEvent(c->comp_stream, c->ring_h2d[r], 0));
            rc = enqueue(c, s, c->comp_stream, s.d_params, n, ld, flags, logl ? d_logl + off : nullptr, chi2 ? d_chi2 + off : nullptr, nullptr);
            if (rc) return rc;
            CU(cudaEventRecord(s.done, c->comp_stream));
            c->last_slot = -1;
            c->last_ring = r;
        }
        if (logl) CU(cudaMemcpyAsync(pin_logl ? logl : c->pipe_hout, d_logl, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, c->comp_stream));
        if (chi2) CU(cudaMemcpyAsync(pin_chi2 ? chi2 : c->pipe_hout + B, d_chi2, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, c->comp_stream));
        CU(cudaStreamSynchronize(c->comp_stream));
        if (logl && !pin_logl) memcpy(logl, c->pipe_hout, sizeof(double) * (size_t)B);
        if (chi2 && !pin_chi2) memcpy(chi2, c->pipe_hout + B, sizeof(double) * (size_t)B);
        return MCALF_OK;
    }

    // host pointers: slices pipelined through two pinned staging slots, one stream each, so the
    // H2D of slice k+1 and the D2H of slice k-1 overlap the kernel of slice k
    long long slice = c->slice;
    // pageable input: the staging memcpy runs on this thread, so smaller slices overlap it better with the kernels
    if (B > 8192 && !c->is_pinned(params)) slice = std::min<long long>(slice, 8192);
    if (flux) slice = std::max<long long>(1, std::min<long long>(slice, (long long)((64u << 20) / ((size_t)c->P.npix * esize))));
    slice = std::min(slice, B);
    const size_t flux_bytes = flux ? (size_t)slice * c->P.npix * esize : 0;
    // caller buffers that are already page-locked (mcalf_host_alloc, torch pin_memory) skip the staging copy
    // (small calls are always staged: the pointer queries would cost more than the copies)
    const bool small = (size_t)B * (size_t)ld * sizeof(double) <= (64u << 10) && !flux;
    const bool pin_in = !small && c->is_pinned(params);
    const bool pin_logl = !small && c->is_pinned(logl), pin_chi2 = !small && c->is_pinned(chi2), pin_flux = !small && c->is_pinned(flux);
    struct Pending { long long off = 0, n = 0; bool live = false; } pend[NBUF];
    auto drain = [&](int k) -> int {
        Slot &s = c->slot[k];
        if (!pend[k].live) return MCALF_OK;
        CU(cudaEventSynchronize(s.done));
        if (logl && !pin_logl) memcpy(logl + pend[k].off, s.h_out, sizeof(double) * (size_t)pend[k].n);
        if (chi2 && !pin_chi2) memcpy(chi2 + pend[k].off, s.h_out + pend[k].n, sizeof(double) * (size_t)pend[k].n);
        if (flux && !pin_flux) memcpy((char *)flux + (size_t)pend[k].off * c->P.npix * esize, s.h_flux, (size_t)pend[k].n * c->P.npix * esize);
        pend[k].live = false;
        return MCALF_OK;
    };
    int k = 0;
    for (long long off = 0; off < B; off += slice, k ^= 1) {
        const long long n = std::min(slice, B - off);
        Slot &s = c->slot[k];
        int rc = drain(k);
        if (rc) return rc;
        rc = ensure_slot(c, s, slice, ld, flux_bytes, true);
        if (rc) return rc;
        const double *src = params + off * ld;
        if (!pin_in) {
            memcpy(s.h_params, src, sizeof(double) * (size_t)n * (size_t)ld);
            src = s.h_params;
        }
        CU(cudaMemcpyAsync(s.d_params, src, sizeof(double) * (size_t)n * (size_t)ld, cudaMemcpyHostToDevice, s.stream));
        rc = enqueue(c, s, s.stream, s.d_params, n, ld, flags, logl ? s.d_out : nullptr, chi2 ? s.d_out + n : nullptr,
                     flux ? s.d_flux : nullptr);
        if (rc) return rc;
        if (logl) CU(cudaMemcpyAsync(pin_logl ? logl + off : s.h_out, s.d_out, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s.stream));
        if (chi2) CU(cudaMemcpyAsync(pin_chi2 ? chi2 + off : s.h_out + n, s.d_out + n, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s.stream));
        if (flux) CU(cudaMemcpyAsync(pin_flux ? (void *)((char *)flux + (size_t)off * c->P.npix * esize) : s.h_flux, s.d_flux,
                                     (size_t)n * c->P.npix * esize, cudaMemcpyDeviceToHost, s.stream));
        CU(cudaEventRecord(s.done, s.stream));
        pend[k].off = off;
        pend[k].n = n;
        pend[k].live = true;
        c->last_slot = k;
        c->last_ring = -1;
    }
    for (int j = 0; j < NBUF; ++j) {
        int rc = drain(j);
        if (rc) return rc;
    }
    return MCALF_OK;
}

}  // namespace

extern "C" {

int mcalf_abi_version(void) { return MCALF_ABI_VERSION; }
int mcalf_is_checked_build(void) {
#if defined(MCALF_CHECK)
    return 1;
#else
    return 0;
#endif
}
const char *mcalf_last_error(void) { return g_err; }

int mcalf_create(const mcalf_problem_t *p, int device, mcalf_ctx **out) {
    if (!p || !out) return fail(MCALF_E_INVALID, "null argument");
    *out = nullptr;
    if (p->abi_version != MCALF_ABI_VERSION) return fail(MCALF_E_INVALID, "ABI version %d, library is %d", p->abi_version, MCALF_ABI_VERSION);
    if (p->npix < 1 || !p->wave || !p->flux || !p->err) return fail(MCALF_E_INVALID, "empty spectrum");
    if (p->nlines < 1 || !p->line_wrest || !p->line_f || !p->line_gamma) return fail(MCALF_E_INVALID, "no lines");
    if (p->ncompmax < 0 || p->nfill < 0) return fail(MCALF_E_INVALID, "negative component count");
    const int startind = (p->free_specres ? 1 : 0) + (p->free_cont ? 1 : 0);          // hires_fitter.py:169-174
    const int ndim = startind + 1 + 3 * (p->ncompmax + p->nfill);                     // :200
    if (p->ndim != ndim) return fail(MCALF_E_INVALID, "ndim %d inconsistent with the layout (%d)", p->ndim, ndim);
    if (!p->bounds_lo || !p->bounds_hi) return fail(MCALF_E_INVALID, "null bounds");
    if (!(p->velstep > 0.0)) return fail(MCALF_E_INVALID, "velstep must be positive");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1)
        return fail(MCALF_E_NODEVICE, "no CUDA device (%s): this library never computes on the host", e == cudaSuccess ? "count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(MCALF_E_INVALID, "device %d out of range (%d devices)", device, ndev);
    ON_DEVICE(device);

    mcalf_ctx *c = new mcalf_ctx();
    struct Guard {                      // every early error return below releases the half-built context
        mcalf_ctx *p;
        ~Guard() { if (p) mcalf_destroy(p); }
    } guard{c};
    c->device = device;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    DevProblem &P = c->P;
    const int npix = p->npix;
    P.npix = npix;
    P.npix4 = (npix + 7) & ~7;             // two float4 groups per stencil thread
    P.nlines = p->nlines;
    P.ncompmax = p->ncompmax;
    P.nfill = p->nfill;
    P.ndim = ndim;
    P.ndim_pad = (std::max(ndim, 5) + 1) & ~1;
    P.startind = startind;
    P.endind = startind + 3 * p->ncompmax + 1;                                         // :176
    P.free_specres = p->free_specres ? 1 : 0;
    P.free_cont = p->free_cont ? 1 : 0;
    P.asymmlike = p->asymmlike ? 1 : 0;
    P.fixed_specres = p->fixed_specres;
    P.fixed_cont = p->fixed_cont;
    P.velstep = p->velstep;
    P.asym_t5 = p->asym_thresh5;
    P.asym_t4 = p->asym_thresh4;
    P.a_max = A_MAX_DEFAULT;
    P.eps_cull = 0.0f;
    P.eps_far = 3e-9f;      // per (line, chunk) pair; measured: 1e-9 -> 3e-9 is +1 % with no change of the achieved accuracy
    P.Lmax = std::max(p->ncompmax * p->nlines + p->nfill, std::max(p->nlines, 1));
    P.lam_ref = p->wave[npix / 2];
    if (!(P.lam_ref > 0.0)) return fail(MCALF_E_INVALID, "non-positive wavelength");

    // largest LSF half-width any sample may need (hires_fitter.py:454-459)
    double max_res = p->max_specres;
    if (!(max_res > 0.0)) {
        max_res = p->fixed_specres;
        if (p->free_specres) max_res = std::max(p->bounds_lo[0], p->bounds_hi[0]);
    }
    int nmax = (int)ceil(TRUNC_SIGMAS_H * (max_res / FWHM_TO_SIGMA_H) / p->velstep) + 1;
    if (nmax < 1) nmax = 1;
    P.nmax = nmax;
    P.nmax4 = (nmax + 3) & ~3;
    P.halo = P.nmax4;

    // per-pixel tables
    std::vector<double> wave(p->wave, p->wave + npix), obj(npix), w(npix), obj_raw(p->flux, p->flux + npix), isig(npix);
    std::vector<float> obj_hi(P.npix4, 0.0f), obj_lo(P.npix4, 0.0f), w32(P.npix4, 0.0f);
    double csum = 0.0;
    P.chi2_add = 0.0;
    for (int i = 0; i < npix; ++i) {
        if (!(wave[i] > 0.0)) return fail(MCALF_E_INVALID, "non-positive wavelength at pixel %d", i);
        const double f = p->flux[i], er = p->err[i];
        isig[i] = 1.0 / er;
        // nansum (:294) drops the pixel when its term is NaN: NaN flux, NaN error, or zero error
        // (w = inf gives inf - inf)
        const bool valid = !(f != f) && !(er != er) && er != 0.0;
        if (valid) {
            const double wi = 1.0 / (er * er);
            obj[i] = f;
            w[i] = wi;
            csum += -log(wi) + log(2.0 * 3.14159265358979323846);
        } else {
            obj[i] = 0.0;
            w[i] = 0.0;
            // chi2 (:246) has no -log(w) term: a zero-error pixel contributes inf*(resid^2) = +inf there
            if (er == 0.0 && !(f != f)) P.chi2_add = INFINITY;
        }
        obj_hi[i] = (float)obj[i];
        obj_lo[i] = (float)(obj[i] - (double)obj_hi[i]);
        w32[i] = (float)w[i];
    }
    P.logC = -0.5 * csum;

    std::vector<ChunkDesc> chunks;
    std::vector<float> dhi, dlo;
    build_chunks(wave.data(), npix, P.lam_ref, chunks, dhi, dlo);
    P.nchunks = (int)chunks.size();
    {   // pass-A geometry: lane groups of cslot_w chunks
        int w = 1, lw = 0;
        while (w < 32 && w < P.nchunks) { w <<= 1; ++lw; }
        P.cslot_w = w;
        P.cslot_lw = lw;
        P.vwarps = std::max(1, 8 / (32 / w));      // about eight slots whatever the chunk count
        P.nslots = P.vwarps * (32 / w);
        P.mwords = (P.Lmax + 31) / 32;
        int minlen = 1 << 30;
        for (const ChunkDesc &cd : chunks) minlen = std::min(minlen, cd.len);
        P.scratch_in_flux = minlen >= (FF_NC_HOST * P.nslots + 1) + P.Lmax + 32 ? 1 : 0;   // + the bank skew
    }

    // periodic halo of the depth buffer (astropy boundary='wrap', hires_fitter.py:463-464): cell H - 1 - j holds pixel
    // (npix - 1 - j) mod npix, cell H + npix + j holds pixel j mod npix
    std::vector<int> halo_src;
    {
        const int H = P.halo, tail = H + 12 + (P.npix4 - npix);
        for (int cell = 0; cell < H; ++cell) {
            const int j = H - 1 - cell;
            int src = (npix - 1 - j) % npix;
            if (src < 0) src += npix;
            halo_src.push_back(src);
        }
        for (int j = 0; j < tail; ++j) halo_src.push_back(j % npix);
        P.nhalo = (int)halo_src.size();
    }

    std::vector<double> lw(p->line_wrest, p->line_wrest + p->nlines), lf(p->line_f, p->line_f + p->nlines),
        lg(p->line_gamma, p->line_gamma + p->nlines);
    lw.push_back(p->fill_wrest);
    lf.push_back(p->fill_f);
    lg.push_back(p->fill_gamma);
    std::vector<double> blo(p->bounds_lo, p->bounds_lo + ndim), bhi(p->bounds_hi, p->bounds_hi + ndim);

    int rc;
#define UPF4(vec, field) { const float *tmp_ = nullptr; if ((rc = upload(c, vec, &tmp_)) != MCALF_OK) return rc; P.field = reinterpret_cast<const float4 *>(tmp_); }
#define UP(vec, field) if ((rc = upload(c, vec, &P.field)) != MCALF_OK) return rc;
    std::vector<PairF> ph, pl;
    build_pair_table(chunks, dhi, ph);
    build_pair_table(chunks, dlo, pl);
    static_assert(sizeof(PairF) == sizeof(float2), "pair table layout");
#define UPF2(vec, field) { const PairF *tmp_ = nullptr; if ((rc = upload(c, vec, &tmp_)) != MCALF_OK) return rc; P.field = reinterpret_cast<const float2 *>(tmp_); }
    UPF2(ph, dhi2) UPF2(pl, dlo2) UPF4(obj_hi, obj_hi4) UPF4(obj_lo, obj_lo4) UPF4(w32, w4) UP(chunks, chunks) UP(halo_src, halo_src) UP(wave, wave) UP(obj, obj) UP(w, w)
    UP(obj_raw, obj_raw) UP(isig, isig) UP(lw, line_wrest) UP(lf, line_f) UP(lg, line_gamma) UP(blo, blo) UP(bhi, bhi)
#undef UP
#undef UPF4
#undef UPF2
    e = cudaMalloc((void **)&c->d_stats, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(c->d_stats, 0, 8 * sizeof(unsigned long long));
    if (e != cudaSuccess) return fail(MCALF_E_CUDA, "stats buffer: %s", cudaGetErrorString(e));
    if ((rc = choose_launch(c)) != MCALF_OK) return rc;
    guard.p = nullptr;
    *out = c;
    return MCALF_OK;
}

void mcalf_destroy(mcalf_ctx *c) {
    if (!c) return;
    DeviceGuard guard_(c->device);
    for (Slot &s : c->slot) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        if (s.h_params) cudaFreeHost(s.h_params);
        if (s.d_params) cudaFree(s.d_params);
        if (s.h_out) cudaFreeHost(s.h_out);
        if (s.d_out) cudaFree(s.d_out);
        if (s.h_flux) cudaFreeHost(s.h_flux);
        if (s.d_flux) cudaFree(s.d_flux);
        if (s.counters) cudaFree(s.counters);
        if (s.fallback) cudaFree(s.fallback);
        if (s.done) cudaEventDestroy(s.done);
        if (s.k0) cudaEventDestroy(s.k0);
        if (s.k1) cudaEventDestroy(s.k1);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    if (c->zc_buf) cudaFreeHost(c->zc_buf);
    for (Slot *sp : {&c->dev_slot, &c->zc_slot}) {
        Slot &s = *sp;
        if (s.stream) cudaStreamSynchronize(s.stream);
        if (s.counters) cudaFree(s.counters);
        if (s.fallback) cudaFree(s.fallback);
        if (s.done) cudaEventDestroy(s.done);
        if (s.k0) cudaEventDestroy(s.k0);
        if (s.k1) cudaEventDestroy(s.k1);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    for (Slot &s : c->ring) {
        if (s.h_params) cudaFreeHost(s.h_params);
        if (s.d_params) cudaFree(s.d_params);
        if (s.h_out) cudaFreeHost(s.h_out);
        if (s.d_out) cudaFree(s.d_out);
        if (s.counters) cudaFree(s.counters);
        if (s.fallback) cudaFree(s.fallback);
        if (s.done) cudaEventDestroy(s.done);
        if (s.k0) cudaEventDestroy(s.k0);
        if (s.k1) cudaEventDestroy(s.k1);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    for (int r = 0; r < NRING; ++r) if (c->ring_h2d[r]) cudaEventDestroy(c->ring_h2d[r]);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->comp_stream) cudaStreamDestroy(c->comp_stream);
    if (c->pipe_dout) cudaFree(c->pipe_dout);
    if (c->pipe_hout) cudaFreeHost(c->pipe_hout);
    for (void *d : c->allocs) cudaFree(d);
    if (c->d_stats) cudaFree(c->d_stats);
    if (c->dev_order) cudaEventDestroy(c->dev_order);
    if (c->util_dev) cudaFree(c->util_dev);
    if (c->util_stream) cudaStreamDestroy(c->util_stream);
    delete c;
}

int mcalf_loglike_batch(mcalf_ctx *ctx, const double *params, int64_t B, int64_t ld, uint32_t flags, void *stream,
                        double *logl_out, double *chi2_out) {
    if (!logl_out && !chi2_out && B > 0) return fail(MCALF_E_INVALID, "no output buffer");
    const uint32_t allowed = MCALF_F_UNIT_CUBE | MCALF_F_ON_DEVICE | MCALF_F_FP64 | MCALF_F_TARGONLY | MCALF_F_NO_TRUNC;
    if (flags & ~allowed) return fail(MCALF_E_INVALID, "flag 0x%x not valid for mcalf_loglike_batch", flags & ~allowed);
    return run_batch(ctx, params, B, ld, flags, stream, logl_out, chi2_out, nullptr);
}

int mcalf_loglike_batch_peers(mcalf_ctx *ctx, const double *params, int64_t B, int64_t ld, uint32_t flags, void *stream,
                              double *const *logl_peers, int npeers) {
    if (npeers < 1 || npeers > MCALF_MAX_PEERS) return fail(MCALF_E_INVALID, "npeers must be in [1, %d]", MCALF_MAX_PEERS);
    if (!logl_peers) return fail(MCALF_E_INVALID, "null peer list");
    for (int p = 0; p < npeers; ++p)
        if (!logl_peers[p]) return fail(MCALF_E_INVALID, "null peer buffer %d", p);
    const uint32_t allowed = MCALF_F_UNIT_CUBE | MCALF_F_ON_DEVICE | MCALF_F_FP64 | MCALF_F_TARGONLY | MCALF_F_NO_TRUNC;
    if (flags & ~allowed) return fail(MCALF_E_INVALID, "flag 0x%x not valid for mcalf_loglike_batch_peers", flags & ~allowed);
    if (!(flags & MCALF_F_ON_DEVICE)) return fail(MCALF_E_INVALID, "mcalf_loglike_batch_peers takes device pointers (MCALF_F_ON_DEVICE)");
    return run_batch(ctx, params, B, ld, flags, stream, nullptr, nullptr, nullptr, logl_peers, npeers);
}

int mcalf_model_batch(mcalf_ctx *ctx, const double *params, int64_t B, int64_t ld, uint32_t flags, void *stream, void *flux_out) {
    if (!flux_out && B > 0) return fail(MCALF_E_INVALID, "null flux_out");
    return run_batch(ctx, params, B, ld, flags, stream, nullptr, nullptr, flux_out);
}

int mcalf_prior_transform_batch(mcalf_ctx *c, const double *cube, int64_t B, int64_t ld, uint32_t flags, void *stream, double *theta_out) {
    if (!c) return fail(MCALF_E_INVALID, "null context");
    if (B < 0) return fail(MCALF_E_INVALID, "negative batch size");
    if (B == 0) return MCALF_OK;
    if (!cube || !theta_out) return fail(MCALF_E_INVALID, "null buffer");
    if (ld < c->P.ndim) return fail(MCALF_E_INVALID, "ld (%lld) smaller than ndim (%d)", (long long)ld, c->P.ndim);
    ON_DEVICE(c->device);
    ONE_CALL(c);
    if (flags & MCALF_F_ON_DEVICE) {
        CU(launch_prior(c->P, cube, B, ld, flags, theta_out, (cudaStream_t)stream));
        c->kernel_launches += 1;
        return MCALF_OK;
    }
    // host pointers: one device staging buffer per context, grown on demand and kept
    const size_t nin = (size_t)B * (size_t)ld, nout = (size_t)B * (size_t)c->P.ndim;
    if (!c->util_stream) CU(cudaStreamCreateWithFlags(&c->util_stream, cudaStreamNonBlocking));
    if (nin + nout > c->util_cap) {
        if (c->util_dev) CU(cudaFree(c->util_dev));
        c->util_dev = nullptr;
        c->util_cap = 0;
        const size_t cap = std::max<size_t>(nin + nout, 1u << 16);
        CU(cudaMalloc((void **)&c->util_dev, sizeof(double) * cap));
        c->util_cap = cap;
    }
    double *d_in = c->util_dev, *d_out = c->util_dev + nin;
    CU(cudaMemcpyAsync(d_in, cube, sizeof(double) * nin, cudaMemcpyHostToDevice, c->util_stream));
    CU(launch_prior(c->P, d_in, B, ld, flags, d_out, c->util_stream));
    CU(cudaMemcpyAsync(theta_out, d_out, sizeof(double) * nout, cudaMemcpyDeviceToHost, c->util_stream));
    CU(cudaStreamSynchronize(c->util_stream));
    c->kernel_launches += 1;
    return MCALF_OK;
}

int mcalf_voigt_h(int device, int mode, const double *u, const double *a, int64_t n, double *h_out) {
    if (n < 0 || (n > 0 && (!u || !a || !h_out))) return fail(MCALF_E_INVALID, "bad argument");
    if (n == 0) return MCALF_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail(MCALF_E_NODEVICE, "no CUDA device");
    ON_DEVICE(device);
    double *d = nullptr;
    CU(cudaMalloc((void **)&d, sizeof(double) * 3 * (size_t)n));
    cudaError_t e = cudaMemcpy(d, u, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + n, a, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_voigt_h(mode, d, d + n, n, d + 2 * n, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(h_out, d + 2 * n, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(MCALF_E_CUDA, "voigt_h: %s", cudaGetErrorString(e));
    return MCALF_OK;
}

int mcalf_ffma_peak(int device, double *tflops_out) {
    if (!tflops_out) return fail(MCALF_E_INVALID, "null output");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail(MCALF_E_NODEVICE, "no CUDA device");
    ON_DEVICE(device);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    const int threads = 1024, grid = prop.multiProcessorCount * 2, iters = 4096;
    float *d = nullptr;
    CU(cudaMalloc((void **)&d, sizeof(float) * (size_t)grid * threads));
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a));
    CU(cudaEventCreate(&b));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CU(cudaEventRecord(a));
        CU(launch_ffma_peak(d, grid, threads, iters, nullptr));
        CU(cudaEventRecord(b));
        CU(cudaEventSynchronize(b));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, a, b));
        const double flop = 2.0 * 128.0 * (double)iters * (double)grid * threads;
        if (rep > 0) best = std::max(best, flop / (ms * 1e-3) * 1e-12);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    *tflops_out = best;
    return MCALF_OK;
}

int mcalf_get_stats(mcalf_ctx *c, mcalf_stats_t *out) {
    if (!c || !out) return fail(MCALF_E_INVALID, "null argument");
    ON_DEVICE(c->device);
    memset(out, 0, sizeof(*out));
    unsigned long long h[8];
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(h, c->d_stats, sizeof(h), cudaMemcpyDeviceToHost));
    out->kernel_launches = c->kernel_launches;
    out->samples = c->samples;
    out->samples_fp64 = c->samples_fp64;
    out->evals_total = h[0];
    out->evals_wing = h[1];
    out->evals_mixed = h[2];
    out->evals_core = h[3];
    out->evals_culled = h[4];
    out->evals_far = h[5];
    out->evals_core_precise = h[6];
    out->evals_core_straddle = h[7];
    if (c->last_slot >= 0 || c->last_ring >= 0) {
        float ms = 0.f;
        Slot &s = c->last_slot == NBUF + 1 ? c->zc_slot : c->last_slot == NBUF ? c->dev_slot : (c->last_slot >= 0 ? c->slot[c->last_slot] : c->ring[c->last_ring]);
        if (cudaEventElapsedTime(&ms, s.k0, s.k1) == cudaSuccess) out->last_kernel_ms = ms;
        else cudaGetLastError();      // pipelined host slices are only timed while collect_stats is on
    }
    return MCALF_OK;
}

int mcalf_reset_stats(mcalf_ctx *c) {
    if (!c) return fail(MCALF_E_INVALID, "null context");
    ON_DEVICE(c->device);
    CU(cudaDeviceSynchronize());
    CU(cudaMemset(c->d_stats, 0, 8 * sizeof(unsigned long long)));
    c->kernel_launches = c->samples = c->samples_fp64 = 0;
    return MCALF_OK;
}

int mcalf_set_option(mcalf_ctx *c, const char *name, double value) {
    if (!c || !name) return fail(MCALF_E_INVALID, "null argument");
    ON_DEVICE(c->device);
    if (!strcmp(name, "cull_eps")) {
        if (!(value >= 0.0)) return fail(MCALF_E_INVALID, "cull_eps must be >= 0");
        c->P.eps_cull = (float)value;
    } else if (!strcmp(name, "far_eps")) {
        if (!(value >= 0.0) || value > 1e-6) return fail(MCALF_E_INVALID, "far_eps must be in [0, 1e-6]");
        c->P.eps_far = (float)value;
    } else if (!strcmp(name, "a_max")) {
        if (!(value >= 0.0) || value > A_MAX_LIMIT) return fail(MCALF_E_INVALID, "a_max must be in [0, %g]", A_MAX_LIMIT);
        c->P.a_max = value;
    } else if (!strcmp(name, "collect_stats")) {
        c->collect_stats = value != 0.0;
    } else if (!strcmp(name, "check_selftest")) {
        c->P.check_selftest = value != 0.0;      // only the -DMCALF_CHECK build looks at it
    } else if (!strcmp(name, "threads")) {
        const int t = (int)value;
        if (t < 0 || t > 1024 || (t % 32)) return fail(MCALF_E_INVALID, "threads must be a multiple of 32 in [0, 1024]");
        const int old = c->threads_opt;
        c->threads_opt = t;
        int rc = choose_launch(c);
        if (rc) { c->threads_opt = old; choose_launch(c); return rc; }
    } else if (!strcmp(name, "ctas_per_sm")) {
        if (value < 0) return fail(MCALF_E_INVALID, "ctas_per_sm must be >= 0");
        c->ctas_opt = (int)value;
        return choose_launch(c);
    } else if (!strcmp(name, "dense")) {
        c->dense_opt = value < 0 ? -1 : (value != 0.0);
        return choose_launch(c);
    } else if (!strcmp(name, "slice")) {
        if (value < 1) return fail(MCALF_E_INVALID, "slice must be >= 1");
        c->slice = (long long)value;
    } else {
        return fail(MCALF_E_INVALID, "unknown option '%s'", name);
    }
    return MCALF_OK;
}

int mcalf_get_option(mcalf_ctx *c, const char *name, double *value) {
    if (!c || !name || !value) return fail(MCALF_E_INVALID, "null argument");
    if (!strcmp(name, "cull_eps")) *value = c->P.eps_cull;
    else if (!strcmp(name, "far_eps")) *value = c->P.eps_far;
    else if (!strcmp(name, "a_max")) *value = c->P.a_max;
    else if (!strcmp(name, "collect_stats")) *value = c->collect_stats;
    else if (!strcmp(name, "threads")) *value = c->threads;
    else if (!strcmp(name, "ctas_per_sm")) *value = c->ctas_per_sm;
    else if (!strcmp(name, "dense")) *value = c->dense;
    else if (!strcmp(name, "slice")) *value = (double)c->slice;
    else return fail(MCALF_E_INVALID, "unknown option '%s'", name);
    return MCALF_OK;
}

int mcalf_get_geometry(mcalf_ctx *c, int64_t *out) {
    if (!c || !out) return fail(MCALF_E_INVALID, "null argument");
    out[0] = c->P.npix;
    out[1] = c->P.nchunks;
    out[2] = c->threads;
    out[3] = c->ctas_per_sm;
    out[4] = c->sm_count;
    out[5] = (int64_t)c->smem_fast;
    out[6] = c->P.halo;
    out[7] = c->P.Lmax;
    return MCALF_OK;
}

int mcalf_host_alloc(void **ptr, uint64_t bytes) {
    if (!ptr) return fail(MCALF_E_INVALID, "null argument");
    CU(cudaMallocHost(ptr, bytes ? bytes : 1));
    return MCALF_OK;
}

int mcalf_host_free(void *ptr) {
    if (ptr) CU(cudaFreeHost(ptr));
    return MCALF_OK;
}

}  // extern "C"
