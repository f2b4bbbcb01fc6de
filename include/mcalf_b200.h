/*
 * mcalf_b200.h -- C ABI of the B200-native MC-ALF likelihood hot path (libmcalf_b200.so).
 *
 * The reference (matteofox/MC-ALF) is pure Python and has no FFI of its own; the boundary this
 * library sits behind is the set of bound methods of `als_fitter` that the nested samplers call
 * (mcalf/routines/hires_fitter.py).  Each entry point below names the reference code it replaces.
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * MCALF_E_* code, with a thread-local message available from mcalf_last_error().  The caller owns
 * every in/out buffer.  A context is bound to one CUDA device and serves one call at a time: a second
 * thread entering a batch call while one is in flight gets MCALF_E_INVALID ("context busy"); use one
 * context per thread.  Every call leaves the calling thread's current CUDA device as it found it.
 * There is no CPU fallback: without a usable CUDA device mcalf_create() fails.
 */
#ifndef MCALF_B200_H
#define MCALF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCALF_ABI_VERSION 2

/* error codes */
#define MCALF_OK 0
#define MCALF_E_INVALID (-1)   /* bad argument / inconsistent problem description            */
#define MCALF_E_CUDA (-2)      /* CUDA runtime error (message carries cudaGetErrorString)     */
#define MCALF_E_NODEVICE (-3)  /* no CUDA device: this library never computes on the host     */
#define MCALF_E_RESOURCE (-4)  /* problem does not fit the kernel's shared-memory/line limits */

/* flags for the batch calls (bitwise or) */
#define MCALF_F_UNIT_CUBE 0x01  /* params are unit-cube draws: apply _scale_cube_pc first (hires_fitter.py:202-209) */
#define MCALF_F_ON_DEVICE 0x02  /* params and outputs are device pointers on the context's device                */
#define MCALF_F_FP64 0x04       /* use the fp64 check kernel (trapezoid-rule Faddeeva, ~1e-13)                   */
#define MCALF_F_TARGONLY 0x08   /* skip filler lines: reconstruct_spec(p, targonly=True) (hires_fitter.py:437)   */
#define MCALF_F_ONECOMP 0x10    /* params rows are [specres, continuum, N, z, b]: reconstruct_onecomp (:379-392) */
#define MCALF_F_ONECOMP_FILL 0x20 /* as ONECOMP but with the filler line: reconstruct_onecomp_fill (:394-406)    */
#define MCALF_F_NO_TRUNC 0x40   /* with UNIT_CUBE: _scale_cube_mn semantics, no int() on the ncomp slot (:211-216) */
#define MCALF_F_FLUX_F64 0x80   /* mcalf_model_batch: flux_out is double[B*npix] instead of float[B*npix]        */
#define MCALF_F_ONELINE 0x100   /* params rows are [specres, continuum, N, z, b, line]: ONE line of the line table
                                   (index `line`, nlines = the filler) of one component -- the single-line
                                   voigt_model the reference's calc_w integrates (:467-491)                      */

/*
 * Everything als_fitter.__init__ (hires_fitter.py:32-200) leaves behind that the likelihood reads.
 * All arrays are host pointers, copied by mcalf_create().
 */
typedef struct mcalf_problem {
    int32_t abi_version;      /* MCALF_ABI_VERSION */
    int32_t npix;             /* pixels after the wavefit mask, windows concatenated (:75-82) */
    const double *wave;       /* obj_wl  [npix], Angstrom                                      */
    const double *flux;       /* obj     [npix] (NaN allowed: pixel dropped, nansum :294)      */
    const double *err;        /* obj_noise [npix] (0 / NaN allowed: pixel dropped)             */
    double velstep;           /* km/s per pixel, sigma-clipped median (:84-87)                 */
    int32_t nlines;           /* numlines (:91)                                                */
    const double *line_wrest; /* [nlines] Angstrom (:104-113)                                  */
    const double *line_f;     /* [nlines] oscillator strengths                                 */
    const double *line_gamma; /* [nlines] damping constants, 1/s                               */
    double fill_wrest, fill_f, fill_gamma; /* linefill (:120-121): wrest = 250 A               */
    int32_t ncompmax;         /* ncomp[1] (:48)                                                */
    int32_t nfill;            /* (:44)                                                         */
    int32_t free_specres;     /* len(specres) > 1 (:59-62)                                     */
    int32_t free_cont;        /* len(contval) > 1 (:54-57)                                     */
    double fixed_specres;     /* max(specres) when not free (:417)                             */
    double fixed_cont;        /* contval[0] when not free (:425)                               */
    int32_t ndim;             /* startind + 1 + 3*(ncompmax+nfill) (:200)                      */
    int32_t asymmlike;        /* Asymmlike (:296)                                              */
    const double *bounds_lo;  /* [ndim] min(bounds[i]) (:184-198)                              */
    const double *bounds_hi;  /* [ndim] max(bounds[i])                                         */
    double asym_thresh5;      /* gauss_cdf[2] + gracenum (:300)                                */
    double asym_thresh4;      /* gauss_cdf[1] + gracenum (:302)                                */
    double max_specres;       /* largest specres any call may carry; sizes the LSF halo (0: from bounds/fixed) */
} mcalf_problem_t;

typedef struct mcalf_ctx mcalf_ctx;

/* counters accumulated over the calls of one context (reset with mcalf_reset_stats) */
typedef struct mcalf_stats {
    uint64_t kernel_launches;  /* CUDA kernels this library launched                                    */
    uint64_t samples;          /* parameter vectors evaluated                                           */
    uint64_t samples_fp64;     /* ... of which by the fp64 kernel (MCALF_F_FP64, or re-routed: a > a_max) */
    /* the next seven only advance while option "collect_stats" is 1 (fp32 kernel)                       */
    uint64_t evals_total;      /* (line, pixel) Voigt evaluations the reference would perform           */
    uint64_t evals_wing;       /* ... in (line, chunk) pairs served by the wing-only form               */
    uint64_t evals_mixed;      /* ... in pairs that may contain line-core pixels                        */
    uint64_t evals_core;       /* (line, pixel) evaluations executed with a line-core form (whole 64-pixel row pairs) */
    uint64_t evals_culled;     /* ... skipped under the proven tau < cull_eps bound                     */
    uint64_t evals_far;        /* ... in pairs folded into the chunk's far-field polynomial             */
    uint64_t evals_core_precise; /* ... of the core ones that took the two-float form (kappa > 8)          */
    uint64_t evals_core_straddle; /* ... of the core ones in row pairs that also needed the wing form (per-pixel select) */
    double last_kernel_ms;     /* device time of the last batch call's kernels (CUDA events): device-pointer calls always,
                                  host-pointer calls only while collect_stats is on */
} mcalf_stats_t;

/* Build a context on CUDA device `device` (replaces als_fitter.__init__ state, :65-200). */
int mcalf_create(const mcalf_problem_t *problem, int device, mcalf_ctx **out);
void mcalf_destroy(mcalf_ctx *ctx);

/*
 * logL for B parameter vectors (row b at params + b*ld): replaces lnlhood_worker (:287-328) =
 * reconstruct_spec (:409-449) -> voigt_model/voigt_tau (:331-377) -> convolve_model (:452-464) ->
 * -0.5*nansum(...).  chi2_out may be NULL (else: chi2(p), :236-248).  stream: cudaStream_t or NULL.
 * With host pointers the call returns after the results landed (small calls go through mapped pinned
 * memory, large ones through a copy-stream / compute-stream pipeline over pinned staging slices);
 * with MCALF_F_ON_DEVICE it only enqueues work on `stream` (successive device calls of one context may use
 * different streams: each is ordered after the previous one with an event, because they share the
 * context's work counters).
 */
int mcalf_loglike_batch(mcalf_ctx *ctx, const double *params, int64_t B, int64_t ld, uint32_t flags,
                        void *stream, double *logl_out, double *chi2_out);

/*
 * The sharded form (MPI ranks each owning an als_fitter, cli.py:37-41,156-158, become one process
 * per GPU each evaluating a contiguous shard of one global batch).  As mcalf_loglike_batch with
 * MCALF_F_ON_DEVICE, but the kernel stores logL of sample b to logl_peers[p][b] for every p < npeers (<= 8):
 * the buffers are the ranks' gather buffers -- this GPU's own and the other GPUs' mapped into this device's
 * address space (CUDA IPC / symmetric memory over NVLink) -- each already offset to where this shard starts.
 * The logL gather is thus fused into the kernel's tail as peer stores; no collective follows, only the
 * caller's cross-rank barrier.
 */
int mcalf_loglike_batch_peers(mcalf_ctx *ctx, const double *params, int64_t B, int64_t ld, uint32_t flags,
                              void *stream, double *const *logl_peers, int npeers);

/* Model flux for B parameter vectors, [B, npix] row-major: replaces reconstruct_spec (:409-449),
 * reconstruct_onecomp (:379-392) and reconstruct_onecomp_fill (:394-406) via flags. */
int mcalf_model_batch(mcalf_ctx *ctx, const double *params, int64_t B, int64_t ld, uint32_t flags,
                      void *stream, void *flux_out);

/* Unit cube -> physical parameters for B rows: replaces _scale_cube_pc (:202-209), or
 * _scale_cube_mn (:211-216) with MCALF_F_NO_TRUNC. */
int mcalf_prior_transform_batch(mcalf_ctx *ctx, const double *cube, int64_t B, int64_t ld, uint32_t flags,
                                void *stream, double *theta_out);

/* Re w(u + i a) element-wise with the kernels' own device code (mode 0: fp32 wing / two-float core
 * forms, mode 2: fp32 wing / short core form of weak lines, mode 3: the weak-line forms with the core
 * boundary at s = 16 and the wide-interval wing polynomial, mode 1: fp64 check path); for unit tests
 * against scipy.special.wofz (:365). Host pointers. */
int mcalf_voigt_h(int device, int mode, const double *u, const double *a, int64_t n, double *h_out);

int mcalf_get_stats(mcalf_ctx *ctx, mcalf_stats_t *out);
int mcalf_reset_stats(mcalf_ctx *ctx);
/* Options: "cull_eps"  (line, chunk) pairs whose optical depth is provably below it are skipped
 *                      (default 0 = never: the reference never skips);
 *          "far_eps"   optical-depth error allowed to a (line, chunk) pair that is folded into the
 *                      chunk's far-field expansion of the Lorentzian wings (default 3e-9; 0 = never);
 *          "a_max"     damping parameters above it route the sample to the fp64 kernel (default 0.01,
 *                      upper limit 0.02: the validity range of the fp32 line-core series);
 *          "collect_stats" 0/1; "check_selftest" 0/1 (checked build only: the next launches report a violation on purpose);
 *          "threads" CTA size of the fp32 kernel (multiple of 32, 0 = automatic);
 *          "ctas_per_sm" persistent CTAs per SM (0 = what the occupancy calculator allows);
 *          "dense" 1/0/-1: force / forbid / choose automatically (default) the 48-register build of the fp32 kernel
 *                  (CTAs of <= 256 threads; chosen when it seats more CTAs per SM on a long spectrum);
 *          "slice" samples per pipelined slice of the host-pointer path. */
int mcalf_set_option(mcalf_ctx *ctx, const char *name, double value);
int mcalf_get_option(mcalf_ctx *ctx, const char *name, double *value);

/* Geometry the context chose: npix, number of chunks, CTA size, CTAs per SM, SM count, dynamic shared
 * memory per CTA, LSF halo, maximum line count.  out[8]. */
int mcalf_get_geometry(mcalf_ctx *ctx, int64_t *out);

/* Measured FP32 peak of `device`: an FFMA-only kernel, CUDA-event timed; returns TFLOP/s (FMA = 2). */
int mcalf_ffma_peak(int device, double *tflops_out);

/* pinned host buffers for callers that want the fast host path */
int mcalf_host_alloc(void **ptr, uint64_t bytes);
int mcalf_host_free(void *ptr);

const char *mcalf_last_error(void);
int mcalf_abi_version(void);
/* 1 for libmcalf_b200_check.so (-DMCALF_CHECK: every index the fp32 kernel forms is validated on the device and a
 * violation makes the batch call fail with MCALF_E_CUDA), 0 for the product library. */
int mcalf_is_checked_build(void);

#ifdef __cplusplus
}
#endif
#endif /* MCALF_B200_H */
