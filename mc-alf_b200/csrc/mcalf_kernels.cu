// mcalf_kernels.cu -- sm_100a kernels of the MC-ALF likelihood hot path.
//
// Replaces, per parameter vector ("sample"), the reference chain
//   lnlhood_worker (hires_fitter.py:287-328) -> reconstruct_spec (:409-449) -> voigt_model/voigt_tau
//   (:331-377, scipy.special.wofz) -> convolve_model (:452-464, astropy convolve) -> -0.5*nansum(...)
// with ONE persistent kernel: a CTA takes a sample, stages its line table in shared memory, every
// warp synthesises exp(-sum tau) for 256-pixel chunks with the optical depth held in registers,
// the transmission goes to shared memory only, the periodic Gaussian LSF stencil and the chi-square
// reduction run from there, and one double per sample reaches HBM.
//
// Two kernels:
//   mcalf_fast_kernel   fp32 arithmetic (voigt_math.cuh), fp64 per-line set-up and final reduction
//   mcalf_fp64_kernel   everything in fp64 (trapezoid-rule Faddeeva): the check path, and the
//                       landing place for samples outside the fast path's domain (a > a_max)
#include <cuda_runtime.h>
#include <stdint.h>

#include "mcalf_device.h"
#include "voigt_math.cuh"

namespace mcalf {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Shared-memory loads through an explicit 32-bit shared-window address (one register, computed once): in the hot
// per-line loop the compiler otherwise re-derives the window base and the array offset for every access.
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds128(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
// One lane's fetch-and-increment on shared memory (atom.inc with a bound that is never reached).  For atomicAdd --
// also for atom.shared.add written as PTX -- ptxas emits its warp-aggregation sequence (vote, leader election, two
// population counts, a shuffle: nine more instructions) although the caller has already elected one lane.
__device__ __forceinline__ int atoms_inc(unsigned a) {
    int r;
    asm volatile("atom.shared.inc.u32 %0, [%1], %2;" : "=r"(r) : "r"(a), "r"(0x7fffffff) : "memory");
    return r;
}
__device__ __forceinline__ float lds32(unsigned a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}

// one sample's results: local buffers, and the peers' gather buffers when the batch is one shard of a global one
__device__ __forceinline__ void store_results(const BatchArgs &Bt, long long b, double logl, double chi2) {
    if (Bt.logl_out) Bt.logl_out[b] = logl;
    if (Bt.chi2_out) Bt.chi2_out[b] = chi2;
    for (int p = 0; p < Bt.npeers; ++p) Bt.logl_peer[p][b] = logl;       // st.global on mapped peer memory (NVLink)
}

__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Parse one parameter row exactly as reconstruct_spec does (hires_fitter.py:412-428), after the
// optional prior transform (:202-216).  Executed by the first ndim threads; theta lands in smem.
__device__ __forceinline__ double load_theta(const DevProblem &P, const double *row, int i, uint32_t flags) {
    double v = row[i];
    if (flags & MCALF_F_UNIT_CUBE) {
        // cube*ptp(bounds)+min(bounds): two roundings, as numpy does it (no FMA contraction)
        v = __dadd_rn(__dmul_rn(v, __dsub_rn(P.bhi[i], P.blo[i])), P.blo[i]);
        if (i == P.startind && !(flags & MCALF_F_NO_TRUNC)) v = trunc(v);   // int() at :207
    }
    return v;
}

struct SampleHead {
    double specres, cont;
    int nact;       // active (component, line) pairs + fillers
    int ncomp;      // int(p[startind])
    int nfill;      // fillers evaluated
    int onecomp;    // 0: full vector; 1: reconstruct_onecomp; 2: reconstruct_onecomp_fill; 3: one line of one component
};

constexpr uint32_t MCALF_F_ROW5 = MCALF_F_ONECOMP | MCALF_F_ONECOMP_FILL | MCALF_F_ONELINE;   // rows [specres, cont, N, z, b(, line)]
__device__ __forceinline__ int row_length(const DevProblem &P, uint32_t flags) {
    return (flags & MCALF_F_ONELINE) ? 6 : (flags & MCALF_F_ROW5) ? 5 : P.ndim;
}

__device__ __forceinline__ SampleHead parse_head(const DevProblem &P, const double *th, uint32_t flags) {
    SampleHead h;
    if (flags & MCALF_F_ROW5) {
        h.specres = th[0];
        h.cont = th[1];
        h.onecomp = (flags & MCALF_F_ONELINE) ? 3 : (flags & MCALF_F_ONECOMP_FILL) ? 2 : 1;
        h.ncomp = 1;
        h.nfill = 0;
        h.nact = h.onecomp == 1 ? P.nlines : 1;
        return h;
    }
    h.onecomp = 0;
    h.specres = P.free_specres ? th[0] : P.fixed_specres;
    h.cont = P.free_cont ? th[P.free_specres ? 1 : 0] : P.fixed_cont;
    double nc = th[P.startind];
    int n = (nc != nc) ? 0 : (nc >= (double)P.ncompmax ? P.ncompmax : (nc <= 0.0 ? 0 : (int)nc));
    h.ncomp = n;
    h.nfill = (flags & MCALF_F_TARGONLY) ? 0 : P.nfill;
    h.nact = n * P.nlines + h.nfill;
    return h;
}

// (logN, z, b, atomic line index) of active line t
__device__ __forceinline__ void line_source(const DevProblem &P, const SampleHead &h, const double *th, int t,
                                            double &logN, double &z, double &b, int &li) {
    if (h.onecomp) {
        logN = th[2]; z = th[3]; b = th[4];
        if (h.onecomp == 3) {          // th[5]: index into the line table (nlines = the filler line)
            const double v = th[5];
            li = (v >= 0.0 && v <= (double)P.nlines) ? (int)v : 0;
        } else {
            li = h.onecomp == 2 ? P.nlines : t;
        }
        return;
    }
    const int ntarget = h.ncomp * P.nlines;
    if (t < ntarget) {
        const int comp = t / P.nlines;
        li = t - comp * P.nlines;
        const int o = 1 + 3 * comp + P.startind;
        logN = th[o]; z = th[o + 1]; b = th[o + 2];
    } else {
        const int k = t - ntarget;
        li = P.nlines;   // filler
        const int o = 3 * k + P.endind;
        logN = th[o]; z = th[o + 1]; b = th[o + 2];
    }
}

// Normalised LSF taps into smem as floats, laid out for the register-blocked stencil:
// G[m] = g_{m-n4}, m = 0 .. 2 n4, zero where |m - n4| > n; zero padded to 2 n4 + 4.   One warp.
__device__ __forceinline__ int build_taps(const DevProblem &P, double specres, float *G, int lane, int &nhalf) {
    int n = 0;
    double sigma = 1.0;
    const bool conv = specres > P.velstep;                       // hires_fitter.py:445
    if (conv) lsf_geometry(specres, P.velstep, sigma, n);
    nhalf = n;
    if (n > P.nmax || n < 0) return -1;                          // wider than the halo the host sized from the bounds
    const int n4 = (n + 3) & ~3;
    const double inv2s2 = conv ? 0.5 / (sigma * sigma) : 0.0;
    double part = 0.0;
    for (int k = lane; k <= n; k += 32) {
        const double g = exp(-(double)(k * k) * inv2s2);
        part += (k == 0) ? g : 2.0 * g;
    }
    const double norm = 1.0 / warp_sum(part);
    for (int m = lane; m < 2 * n4 + 4; m += 32) {
        const int k = m - n4;
        const int ak = k < 0 ? -k : k;
        G[m] = (ak <= n) ? (float)(exp(-(double)(ak * ak) * inv2s2) * norm) : 0.0f;
    }
    return n4;
}

// ---------------------------------------------------------------------------------------------
// fast kernel
// ---------------------------------------------------------------------------------------------
constexpr int PX = 8;            // pixels per lane per chunk (chunk = 32 lanes x 8 = 256 pixels)

constexpr int FF_NC = FF_DEG + 1;
// The classification pass runs on P.vwarps "virtual warps" (a problem constant, not the CTA size): its
// summation and list order, hence every result bit, is independent of the launch geometry.

struct FastSmem {
    double *theta;     // [ndim_pad]
    double *A64;       // [Lmax]
    double *rc64;      // [Lmax]
    LineP *lp;         // [Lmax]
    // uarr and farp live in the chunk's own slice of `flux` instead (which nothing else uses until pass B
    // writes the chunk's depth) when every chunk is long enough: P.scratch_in_flux
    float *uarr;       // [nchunks][Lmax]     U_hi of the near (line, chunk) pairs
    unsigned *nmask;   // [nchunks][mwords]   bit t: line t is near chunk c (evaluated in the direct form)
    unsigned *cmask;   // [nchunks][mwords]   bit t: the core of line t may reach chunk c
    float *farp;       // [nchunks][FF_NC * nslots + 1] partial far-field coefficients, [n][slot] within a chunk
    float *taps;       // [2*nmax4 + 8]
    float *flux;       // [halo + npix4 + halo + 12]
    double *red;       // [64]
    int *misc;         // [8]
    ChunkDesc *chk;    // [nchunks]  copy of the chunk descriptors (48-register build)
    size_t bytes;
};

// The byte offsets are computed once on the host (fast_smem_layout, at context creation) and travel in DevProblem:
// the kernel only adds them to the shared-memory base instead of re-deriving them from the problem sizes.
static SmemLayout make_layout(const DevProblem &P) {
    SmemLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 15) & ~(size_t)15; return (int)at; };
    L.lp = take(sizeof(LineP) * P.Lmax);            // hottest first: offset 0
    L.nmask = take(sizeof(unsigned) * (size_t)P.nchunks * P.mwords);
    L.cmask = take(sizeof(unsigned) * (size_t)P.nchunks * P.mwords);
    L.A64 = take(sizeof(double) * P.Lmax);
    L.rc64 = take(sizeof(double) * P.Lmax);
    L.theta = take(sizeof(double) * P.ndim_pad);
    L.uarr = take(P.scratch_in_flux ? 0 : sizeof(float) * (size_t)P.nchunks * P.Lmax);
    L.farp = take(P.scratch_in_flux ? 0 : sizeof(float) * (size_t)P.nchunks * (FF_NC * P.nslots + 1));
    L.taps = take(sizeof(float) * (2 * P.nmax4 + 8));
    L.red = take(sizeof(double) * 64);
    L.misc = take(sizeof(int) * 8);
    L.chk = take(sizeof(ChunkDesc) * P.nchunks);
    L.flux = take(sizeof(float) * (2 * P.halo + P.npix4 + 12));
    L.bytes = (int)o;
    return L;
}

__device__ __forceinline__ FastSmem carve(unsigned char *base, const DevProblem &P) {
    FastSmem s;
    s.theta = (double *)(base + P.lay.theta);
    s.A64 = (double *)(base + P.lay.A64);
    s.rc64 = (double *)(base + P.lay.rc64);
    s.lp = (LineP *)(base + P.lay.lp);
    s.uarr = (float *)(base + P.lay.uarr);
    s.nmask = (unsigned *)(base + P.lay.nmask);
    s.cmask = (unsigned *)(base + P.lay.cmask);
    s.farp = (float *)(base + P.lay.farp);
    s.taps = (float *)(base + P.lay.taps);
    s.flux = (float *)(base + P.lay.flux);
    s.red = (double *)(base + P.lay.red);
    s.misc = (int *)(base + P.lay.misc);
    s.chk = (ChunkDesc *)(base + P.lay.chk);
    s.bytes = (size_t)P.lay.bytes;
    return s;
}

// STATS: path counters (bench / diagnostics).  EXTRAS: the per-pixel outputs (model flux, Asymmlike counts); the
// plain logL instantiation does not carry that code, which keeps the hot kernel smaller in the instruction cache.
// DENSE: the 48-register build (CTAs of at most 256 threads, five per SM).  ptxas fits the kernel into 48 registers
// with 20 bytes of spills and some rematerialisation: 4 % slower per warp, so it only pays where it seats a fifth CTA
// (cfg 4: 40 instead of 32 warps per SM, +3 %; the host picks it, mcalf_api.cu:choose_launch).
// ONE_EACH: the launch has one CTA per sample (small batches, the scalar callbacks): no work queue, the CTA takes
// sample blockIdx.x and leaves -- two global atomics less on the latency path of a single call.  A separate
// instantiation, because even these few instructions in the hand-over cost the big-batch kernel 1-2 % on short spectra.
template <bool STATS, bool EXTRAS, bool DENSE, bool ONE_EACH>
__global__ void __launch_bounds__(DENSE ? 256 : 1024, DENSE ? 5 : 1)
mcalf_fast_kernel(const __grid_constant__ DevProblem P, const __grid_constant__ BatchArgs Bt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;
    FastSmem S = carve(smem_raw, P);
    const uint32_t flags = Bt.flags;

    __shared__ G1Row g1_smem[MCALF_G1_N];       // Taylor rows of H1 (line-core form), fixed address
    for (int i = tid; i < MCALF_G1_N; i += nthreads) g1_smem[i] = g1_tab_dev[i];
    // the chunk descriptors: the 48-register build reads them from a copy in shared memory (a chunk's first instructions
    // wait on them: +1 % at cfg 4); the 64-register build spills with the copy and keeps the global table (L1-resident)
    if (DENSE) for (int i = tid; i < P.nchunks; i += nthreads) S.chk[i] = P.chunks[i];
    const ChunkDesc *const chk = DENSE ? S.chk : P.chunks;
    if (blockIdx.x == 0 && tid == 0) { Bt.clear_counters[0] = 0u; Bt.clear_counters[1] = 0u; }   // for the slot's next launch
    // per-thread statistics (only summed when Bt.stats != nullptr)
    unsigned long long st_wing = 0, st_mixed = 0, st_core = 0, st_cull = 0, st_total = 0, st_far = 0, st_corep = 0, st_both = 0;

    for (int round = 0;; ++round) {
        // ---- next sample (dynamic: the active-component count varies per sample) ----
        __syncthreads();
        if (tid == 0) {
            S.misc[0] = ONE_EACH ? (round ? 0x7fffffff : (int)blockIdx.x) : (int)atomicAdd(Bt.work_counter, 1u);
            S.misc[3] = 0;
        }
        __syncthreads();
        const long long b = S.misc[0];
        if (b >= Bt.B) break;
        const double *row = Bt.params + b * Bt.ld;
        const int nrow = row_length(P, flags);
        MCALF_CHK(nrow <= P.ndim_pad, 0);
        MCALF_CHK(!P.check_selftest, 31);
        for (int i = tid; i < nrow; i += nthreads) S.theta[i] = load_theta(P, row, i, flags);   // rows may be longer than the CTA
        __syncthreads();
        const SampleHead h = parse_head(P, S.theta, flags);

        // ---- per-line set-up in fp64 (one thread per active line) + LSF taps (last warp) ----
        int bad = 0;
        for (int t = tid; t < h.nact; t += nthreads) {
            double logN, z, bk;
            int li;
            line_source(P, h, S.theta, t, logN, z, bk, li);
            MCALF_CHK(t < P.Lmax && li >= 0 && li <= P.nlines, 1);
            const Line64 L = line_setup64(logN, z, bk, P.line_wrest[li], P.line_f[li], P.line_gamma[li], P.lam_ref);
            S.A64[t] = L.A;
            S.rc64[t] = L.rc;
            S.lp[t] = line_pack_full(L);
            // outside the fp32 path's domain: damping too large, or anything non-finite / non-positive
            if (!(L.a <= P.a_max) || !(L.A > 0.0) || !(L.A < 1e30) || !(L.kappa < 1e30) || !(L.kappa >= 0.0)) bad = 1;
        }
        for (int i = tid; i < P.nchunks * P.mwords; i += nthreads) { S.nmask[i] = 0u; S.cmask[i] = 0u; }
        int n4 = 0;
        if (warp == nwarps - 1) {
            int nhalf = 0;
            n4 = build_taps(P, h.specres, S.taps, lane, nhalf);
            if (lane == 0) { S.misc[2] = n4; S.misc[6] = nhalf; }
        }
        bad = __syncthreads_or(bad);
        if (bad || S.misc[2] < 0 || !(h.specres == h.specres) || !(h.cont == h.cont)) {
            // hand the sample to the fp64 kernel
            if (tid == 0) {
                const unsigned int slot = atomicAdd(Bt.fallback_count, 1u);
                MCALF_CHK((long long)slot < Bt.B, 10);
                Bt.fallback_list[slot] = (int)b;
                if (Bt.fallback_flag) *Bt.fallback_flag = 1;
            }
            continue;
        }
        n4 = S.misc[2];

        // ---- pass A: classify every (line, chunk) pair, lane = chunk ----
        // Far pairs fold into per-(chunk, slot) partial expansions (a slot is one lane group of one of
        // P.vwarps virtual warps and owns the lines slot, slot + nslots, ...; the partials are later
        // summed in slot order).  Near pairs set bit `line` of the chunk's near mask (and core mask) and
        // leave their U: pass B walks the masks in line order, so nothing depends on who classified what.
        {
            const int W = P.cslot_w, LW = P.cslot_lw, SUB = 32 >> LW, NS = P.nslots, MW = P.mwords;
            const int fstride = FF_NC * NS + 1;
            for (int vw = warp; vw < P.vwarps; vw += nwarps) {
                const int slot = vw * SUB + (lane >> LW);
                for (int cg = 0; cg < P.nchunks; cg += W) {
                    const int c = cg + (lane & (W - 1));
                    const bool cact = c < P.nchunks;
                    const int cs = cact ? c : 0;
                    const double rho_s = chk[cs].rho_s;
                    const float ds = chk[cs].ds;
                    float *fslice = P.scratch_in_flux ? S.flux + P.halo + chk[cs].start + (cs & 31) : S.farp + (size_t)cs * fstride;   // skewed: lanes hit distinct banks
                    float *uslice = P.scratch_in_flux ? fslice + fstride : S.uarr + (size_t)cs * P.Lmax;
                    F2 C[FF_NC / 2];               // coefficient pairs {C[2m], C[2m+1]}
#pragma unroll
                    for (int m = 0; m < FF_NC / 2; ++m) C[m] = f2(0.0f);
                    for (int t = slot; t < h.nact; t += NS) {
                        const double U = S.A64[t] * (rho_s - S.rc64[t]);
                        const float Uh = (float)U;
                        const float4 L = *reinterpret_cast<const float4 *>(&S.lp[t]);   // A_hi, a2, c1, ucm
                        const int cls = cact ? chunk_class(L.x, Uh, ds, L.z, L.w, P.eps_cull, P.eps_far) : -1;
                        if (cls == 3) farfield_accumulate(L.x, Uh, ds, L.z, L.y, C);
                        if (cls == 1 || cls == 2) {
                            // scratch stays inside the chunk's own slice of the depth buffer (or inside uarr)
                            MCALF_CHK(P.scratch_in_flux ? (uslice + t >= S.flux + P.halo + P.chunks[cs].start && uslice + t < S.flux + P.halo + P.chunks[cs].start + P.chunks[cs].len)
                                                        : (t < P.Lmax), 3);
                            MCALF_CHK(cs * MW + (t >> 5) < P.nchunks * P.mwords, 2);
                            uslice[t] = Uh;
                            atomicOr(&S.nmask[cs * MW + (t >> 5)], 1u << (t & 31));
                            if (cls == 2) atomicOr(&S.cmask[cs * MW + (t >> 5)], 1u << (t & 31));
                        }
                        if (STATS) {
                            const int len = cact ? P.chunks[c].len : 0;
                            st_total += len;
                            st_cull += cls == 0 ? len : 0;
                            st_far += cls == 3 ? len : 0;
                            st_wing += cls == 1 ? len : 0;
                            st_mixed += cls == 2 ? len : 0;
                        }
                    }
                    if (cact) {
                        MCALF_CHK(P.scratch_in_flux ? (fslice >= S.flux + P.halo + P.chunks[cs].start && fslice + (FF_NC - 1) * NS + slot < S.flux + P.halo + P.chunks[cs].start + P.chunks[cs].len)
                                                    : ((FF_NC - 1) * NS + slot < fstride), 3);
#pragma unroll
                        for (int m = 0; m < FF_NC / 2; ++m) {
                            fslice[(2 * m) * NS + slot] = C[m].x;
                            fslice[(2 * m + 1) * NS + slot] = C[m].y;
                        }
                    }
                }
            }
        }
        __syncthreads();

        // ---- pass B: synthesis, each warp takes chunks; tau stays in registers ----
        // chunks are handed out dynamically: a chunk holding several line cores costs many times one
        // that sees only far lines, and the CTA's warps must meet at the barrier below
        for (;;) {
            int c = 0;
            if (lane == 0) c = atoms_inc(smem_addr(&S.misc[3]));
            c = __shfl_sync(0xffffffffu, c, 0);
            if (c >= P.nchunks) break;
            MCALF_CHK(c >= 0 && c < P.nchunks, 13);
            const ChunkDesc cd = chk[c];
            const int NS = P.nslots;
            MCALF_CHK(cd.start >= 0 && cd.len >= 1 && cd.len <= CHUNK_PIXELS && cd.start + cd.len <= P.npix, 13);
            const float *fslice = P.scratch_in_flux ? S.flux + P.halo + cd.start + (c & 31) : S.farp + (size_t)c * (FF_NC * NS + 1);
            const float *uslice = P.scratch_in_flux ? fslice + (FF_NC * NS + 1) : S.uarr + (size_t)c * P.Lmax;
            // rows (32 consecutive pixels) are held in pairs: {row 2j, row 2j+1} in one 64-bit register pair,
            // so the arithmetic below runs on the packed FFMA2 / FMUL2 forms.  The pair tables are padded, so
            // the loads need no bounds test (lanes beyond the chunk compute on finite junk nobody stores).
            constexpr int PX2 = PX / 2;
            F2 d[PX2], tau[PX2];
            const float2 *dh2 = P.dhi2 + cd.start + lane;
            MCALF_CHK(cd.start + lane + 64 * (PX2 - 1) < P.npix + CHUNK_PIXELS, 5);
#pragma unroll
            for (int j = 0; j < PX2; ++j) {
                const float2 v = __ldg(dh2 + 64 * j);
                d[j] = f2(v.x, v.y);
            }
            // far lines: sum the slots' partial expansions in slot order (lane n sums coefficient n),
            // broadcast, one polynomial per pixel
            {
                const float *fp = fslice + (lane < FF_NC ? lane : 0) * NS;
                float cn = 0.0f;
                if (NS == 8) {                   // the usual slot count: no remainder loop
#pragma unroll
                    for (int sidx = 0; sidx < 8; ++sidx) cn += fp[sidx];
                } else {
#pragma unroll 1
                    for (int sidx = 0; sidx < NS; ++sidx) cn += fp[sidx];
                }
                float C[FF_NC];
#pragma unroll
                for (int n = 0; n < FF_NC; ++n) C[n] = __shfl_sync(0xffffffffu, cn, n);
                F2 x[PX2];
#pragma unroll
                for (int j = 0; j < PX2; ++j) {
                    x[j] = mul2(d[j], f2(cd.inv_ds));
                    tau[j] = f2(C[FF_DEG]);
                }
#pragma unroll
                for (int n = FF_DEG - 1; n >= 0; --n) {
                    const F2 cn2 = f2(C[n]);
#pragma unroll
                    for (int j = 0; j < PX2; ++j) tau[j] = fma2(tau[j], x[j], cn2);
                }
            }
            // near lines in line order, everything in registers.  A line whose core cannot reach the chunk
            // takes the direct wing form on all eight rows.  A line that may have core pixels here is taken
            // row pair by row pair: the warp votes whether the pair's 64 pixels are all beyond the line's core
            // boundary (wing form), all inside the H1 table (core form -- it is the full Voigt function there,
            // so it also serves the pixels just outside the boundary), or straddle both (both forms, per-pixel
            // select).  No pixel index is ever dynamic, so tau never leaves the registers.
            const int MW = P.mwords;
            const unsigned lp_sa = smem_addr(S.lp), us_sa = smem_addr(uslice);
            for (int w = 0; w < MW; ++w) {
                const unsigned cmw = S.cmask[c * MW + w];
                for (unsigned m = S.nmask[c * MW + w]; m; m &= m - 1) {
                    const int bit = __ffs(m) - 1;
                    const int t = (w << 5) + bit;
                    MCALF_CHK(t < h.nact && t < P.Lmax, 1);
                    MCALF_CHK(P.scratch_in_flux ? (uslice + t < S.flux + P.halo + cd.start + cd.len) : (t < P.Lmax), 3);
                    const unsigned la = lp_sa + (unsigned)t * (unsigned)sizeof(LineP);
                    const float4 r0 = lds128(la), r1 = lds128(la + 16), r2 = lds128(la + 32);
                    LineP L;
                    L.A_hi = r0.x; L.a2 = r0.y; L.c1 = r0.z; L.ucm = r0.w;
                    L.cw0 = r1.x; L.cw1 = r1.y; L.cw2 = r1.z; L.cw3 = r1.w;
                    L.cw4 = r2.x; L.scut = r2.y; L.kappa = r2.z; L.a = r2.w;
                    const float Uh = lds32(us_sa + 4u * (unsigned)t);
                    const F2 A2 = f2(L.A_hi), U2 = f2(Uh), a22 = f2(L.a2);
                    if (!((cmw >> bit) & 1u)) {
#pragma unroll
                        for (int j = 0; j < PX2; ++j) {
                            const F2 u = fma2(A2, d[j], U2);
                            tau[j] = wing_acc2(tau[j], fma2(u, u, a22), L);
                        }
                        continue;
                    }
                    L.A_lo = lds32(la + 48);
                    float Uh64, Ul;
                    split2(S.A64[t] * (cd.rho_s - S.rc64[t]), Uh64, Ul);      // Uh64 == Uh
                    const float2 *dl2 = P.dlo2 + cd.start + lane;
#pragma unroll
                    for (int j = 0; j < PX2; ++j) {
                        const float2 dlv = __ldg(dl2 + 64 * j);      // issued before the votes: its latency hides behind them
                        const F2 u = fma2(A2, d[j], U2);
                        const F2 s2 = fma2(u, u, a22);
                        const bool all_wing = __all_sync(0xffffffffu, s2.x >= L.scut && s2.y >= L.scut);
                        if (all_wing) {
                            tau[j] = wing_acc2(tau[j], s2, L);
                            continue;
                        }
                        const bool all_tab = __all_sync(0xffffffffu, fabsf(u.x) <= U_TAB && fabsf(u.y) <= U_TAB);
                        tau[j] = mixed_pair_tau(pair_kind(false, all_tab), tau[j], L, u, s2, d[j], f2(dlv.x, dlv.y), Uh64, Ul, g1_smem);
                        if (STATS) { st_core += 2; st_corep += L.kappa > KAPPA_LEAN ? 2 : 0; st_both += all_tab ? 0 : 2; }
                    }
                }
            }
            __syncwarp();        // the chunk's pass-A scratch (read above) lives where its depth goes now
            {
                float *dst = S.flux + P.halo + cd.start + lane;
                MCALF_CHK(P.halo + cd.start + cd.len <= P.halo + P.npix, 6);
                F2 dep[PX2];
#pragma unroll
                for (int j = 0; j < PX2; ++j) dep[j] = depth32_2(tau[j]);
                if (cd.len == CHUNK_PIXELS) {
#pragma unroll
                    for (int j = 0; j < PX2; ++j) {
                        dst[64 * j] = dep[j].x;
                        dst[64 * j + 32] = dep[j].y;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < PX2; ++j) {
                        const int k = 2 * j * 32 + lane;
                        if (k < cd.len) dst[64 * j] = dep[j].x;
                        if (k + 32 < cd.len) dst[64 * j + 32] = dep[j].y;
                    }
                }
            }
            __syncwarp();
        }
        __syncthreads();

        // ---- periodic halo (astropy boundary='wrap' over the concatenated array, :463-464): the cells before pixel 0 and
        // after the last pixel copy the pixels the host listed (P.halo_src: no modulo arithmetic here) ----
        {
            const int H = P.halo, nh = P.nhalo;
            MCALF_CHK(H + P.npix + (nh - H) <= 2 * P.halo + P.npix4 + 12, 7);
            for (int j = tid; j < nh; j += nthreads) {
                const int src = __ldg(P.halo_src + j);
                MCALF_CHK(src >= 0 && src < P.npix, 7);
                S.flux[j < H ? j : P.npix + j] = S.flux[H + src];       // j < H: cell j; else cell H + npix + (j - H)
            }
        }
        __syncthreads();

        // ---- LSF stencil (8 outputs per thread, LDS.128) fused with continuum, residual, chi-square ----
        const float c_hi = (float)h.cont;
        const float c_lo = (float)(h.cont - (double)c_hi);
        double acc = 0.0;
        int cnt5 = 0, cnt4 = 0;
        const int ngroups = P.npix4 >> 3;
        // taps G[m] are non-zero for m in [n4 - n, n4 + n]: blocks of four taps up to the one holding n4 + n
        const int nb = ((n4 + S.misc[6]) >> 2) + 1;
        for (int g = tid; g < ngroups; g += nthreads) {
            const int o0 = g << 3;
            const float4 *xin = reinterpret_cast<const float4 *>(S.flux + (P.halo + o0 - n4));
            const float4 *gin = reinterpret_cast<const float4 *>(S.taps);
            MCALF_CHK(P.halo + o0 - n4 >= 0 && P.halo + o0 - n4 + 4 * (nb + 2) <= 2 * P.halo + P.npix4 + 12, 8);
            MCALF_CHK(4 * nb <= 2 * P.nmax4 + 8, 4);
            MCALF_CHK(g < P.npix4 / 8, 9);
            // eight outputs per thread: every block of four taps costs one window load and one tap load for 32 FMAs.
            // Each output still adds its taps in order, so the result does not depend on the blocking.
            float4 x0 = xin[0], x1 = xin[1];
            float o_0 = 0.f, o_1 = 0.f, o_2 = 0.f, o_3 = 0.f, o_4 = 0.f, o_5 = 0.f, o_6 = 0.f, o_7 = 0.f;
#pragma unroll 3                 // (the window rotates through three registers quads)
            for (int mb = 0; mb < nb; ++mb) {
                const float4 x2 = xin[mb + 2];
                const float4 gg = gin[mb];
                o_0 = fma32(gg.x, x0.x, o_0); o_1 = fma32(gg.x, x0.y, o_1); o_2 = fma32(gg.x, x0.z, o_2); o_3 = fma32(gg.x, x0.w, o_3);
                o_4 = fma32(gg.x, x1.x, o_4); o_5 = fma32(gg.x, x1.y, o_5); o_6 = fma32(gg.x, x1.z, o_6); o_7 = fma32(gg.x, x1.w, o_7);
                o_0 = fma32(gg.y, x0.y, o_0); o_1 = fma32(gg.y, x0.z, o_1); o_2 = fma32(gg.y, x0.w, o_2); o_3 = fma32(gg.y, x1.x, o_3);
                o_4 = fma32(gg.y, x1.y, o_4); o_5 = fma32(gg.y, x1.z, o_5); o_6 = fma32(gg.y, x1.w, o_6); o_7 = fma32(gg.y, x2.x, o_7);
                o_0 = fma32(gg.z, x0.z, o_0); o_1 = fma32(gg.z, x0.w, o_1); o_2 = fma32(gg.z, x1.x, o_2); o_3 = fma32(gg.z, x1.y, o_3);
                o_4 = fma32(gg.z, x1.z, o_4); o_5 = fma32(gg.z, x1.w, o_5); o_6 = fma32(gg.z, x2.x, o_6); o_7 = fma32(gg.z, x2.y, o_7);
                o_0 = fma32(gg.w, x0.w, o_0); o_1 = fma32(gg.w, x1.x, o_1); o_2 = fma32(gg.w, x1.y, o_2); o_3 = fma32(gg.w, x1.z, o_3);
                o_4 = fma32(gg.w, x1.w, o_4); o_5 = fma32(gg.w, x2.x, o_5); o_6 = fma32(gg.w, x2.y, o_6); o_7 = fma32(gg.w, x2.z, o_7);
                x0 = x1;
                x1 = x2;
            }
            // shared memory holds the absorption DEPTH 1 - exp(-tau); with unit-sum taps the convolved
            // model is cont * (1 - conv(depth)), so every rounding error scales with the depth, not with
            // the continuum (a coherent 1e-7 bias of the continuum level would move chi-square by
            // 2 w sum(resid) 1e-7 -- not small for one-signed residuals).  The pixel tables are padded
            // to a multiple of eight with w = 0, so no pixel needs a bounds test here.
            const F2 mch = f2(-c_hi), mcl = f2(-c_lo), ch2 = f2(c_hi), cl2 = f2(c_lo);
            float part;
            {
                const float4 oh = __ldg(P.obj_hi4 + 2 * g), ol = __ldg(P.obj_lo4 + 2 * g), ww = __ldg(P.w4 + 2 * g);
                const F2 oa = f2(o_0, o_1), ob = f2(o_2, o_3);
                const F2 ra = add2(fma2(ch2, oa, add2(add2(f2(oh.x, oh.y), mch), add2(f2(ol.x, ol.y), mcl))), mul2(cl2, oa));
                const F2 rb = add2(fma2(ch2, ob, add2(add2(f2(oh.z, oh.w), mch), add2(f2(ol.z, ol.w), mcl))), mul2(cl2, ob));
                F2 p2 = mul2(mul2(f2(ww.x, ww.y), ra), ra);
                p2 = fma2(mul2(f2(ww.z, ww.w), rb), rb, p2);
                part = p2.x + p2.y;
            }
            float part2;
            {
                const float4 oh = __ldg(P.obj_hi4 + 2 * g + 1), ol = __ldg(P.obj_lo4 + 2 * g + 1), ww = __ldg(P.w4 + 2 * g + 1);
                const F2 oa = f2(o_4, o_5), ob = f2(o_6, o_7);
                const F2 ra = add2(fma2(ch2, oa, add2(add2(f2(oh.x, oh.y), mch), add2(f2(ol.x, ol.y), mcl))), mul2(cl2, oa));
                const F2 rb = add2(fma2(ch2, ob, add2(add2(f2(oh.z, oh.w), mch), add2(f2(ol.z, ol.w), mcl))), mul2(cl2, ob));
                F2 p2 = mul2(mul2(f2(ww.x, ww.y), ra), ra);
                p2 = fma2(mul2(f2(ww.z, ww.w), rb), rb, p2);
                part2 = p2.x + p2.y;
            }
            if (EXTRAS) {
                const float out[8] = {o_0, o_1, o_2, o_3, o_4, o_5, o_6, o_7};
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int o = o0 + r;
                    if (o < P.npix) {
                        MCALF_CHK(b >= 0 && b < Bt.B, 11);
                        const double md = h.cont - h.cont * (double)out[r];
                        if (Bt.flux_out) {
                            if (flags & MCALF_F_FLUX_F64) ((double *)Bt.flux_out)[b * (long long)P.npix + o] = md;
                            else ((float *)Bt.flux_out)[b * (long long)P.npix + o] = (float)md;
                        }
                        if (P.asymmlike) {
                            const double rs = (P.obj_raw[o] - md) * P.isig[o];
                            cnt5 += rs > 5.0;
                            cnt4 += rs > 4.0;
                        }
                    }
                }
            }
            acc += (double)part;
            acc += (double)part2;
        }
        acc = warp_sum(acc);
        if (EXTRAS && P.asymmlike) { cnt5 = warp_sum_int(cnt5); cnt4 = warp_sum_int(cnt4); }
        if (lane == 0) {
            S.red[warp] = acc;
            if (EXTRAS && P.asymmlike) { ((int *)(S.red + 32))[warp] = cnt5; ((int *)(S.red + 32))[32 + warp] = cnt4; }
        }
        __syncthreads();
        if (warp == 0) {
            double v = lane < nwarps ? S.red[lane] : 0.0;
            v = warp_sum(v);
            int c5 = 0, c4 = 0;
            if (EXTRAS && P.asymmlike) {
                c5 = warp_sum_int(lane < nwarps ? ((int *)(S.red + 32))[lane] : 0);
                c4 = warp_sum_int(lane < nwarps ? ((int *)(S.red + 32))[32 + lane] : 0);
            }
            if (lane == 0) {
                double logl = P.logC - 0.5 * v;
                if (EXTRAS && P.asymmlike && ((double)c5 > P.asym_t5 || (double)c4 > P.asym_t4)) logl = -INFINITY;   // :296-303
                store_results(Bt, b, logl, v + P.chi2_add);
            }
        }
    }

    if (STATS) {
        st_wing = (unsigned long long)warp_sum((double)st_wing);   // exact below 2^53
        st_mixed = (unsigned long long)warp_sum((double)st_mixed);
        st_core = (unsigned long long)warp_sum((double)st_core);
        st_cull = (unsigned long long)warp_sum((double)st_cull);
        st_total = (unsigned long long)warp_sum((double)st_total);
        st_far = (unsigned long long)warp_sum((double)st_far);
        st_corep = (unsigned long long)warp_sum((double)st_corep);
        st_both = (unsigned long long)warp_sum((double)st_both);
        if (lane == 0) {
            atomicAdd(Bt.stats + 0, st_total);
            atomicAdd(Bt.stats + 1, st_wing);
            atomicAdd(Bt.stats + 2, st_mixed);
            atomicAdd(Bt.stats + 3, st_core);
            atomicAdd(Bt.stats + 4, st_cull);
            atomicAdd(Bt.stats + 5, st_far);
            atomicAdd(Bt.stats + 6, st_corep);
            atomicAdd(Bt.stats + 7, st_both);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fp64 kernel: one CTA per sample, plain loops.  idx_list/idx_count select the samples (fallback
// use) or are null (check path: all B samples).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2)
mcalf_fp64_kernel(const __grid_constant__ DevProblem P, const __grid_constant__ BatchArgs Bt, const int *idx_list,
                  const unsigned int *idx_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;
    double *theta = (double *)smem_raw;                         // [ndim_pad]
    double *lA = theta + P.ndim_pad;                            // [Lmax] c/b
    double *lC = lA + P.Lmax;                                   // [Lmax] lambda_c
    double *lK = lC + P.Lmax;                                   // [Lmax] kappa
    double *la = lK + P.Lmax;                                   // [Lmax] a
    double *taps = la + P.Lmax;                                 // [nmax + 1]
    double *red = taps + P.nmax + 1;                            // [64]
    double *flux = red + 64;                                    // [npix]
    const uint32_t flags = Bt.flags;
    const long long total = idx_list ? (long long)*idx_count : Bt.B;

    for (long long it = blockIdx.x; it < total; it += gridDim.x) {
        const long long b = idx_list ? idx_list[it] : it;
        __syncthreads();
        const double *row = Bt.params + b * Bt.ld;
        const int nrow = row_length(P, flags);
        for (int i = tid; i < nrow; i += nthreads) theta[i] = load_theta(P, row, i, flags);
        __syncthreads();
        const SampleHead h = parse_head(P, theta, flags);
        for (int t = tid; t < h.nact; t += nthreads) {
            double logN, z, bk;
            int li;
            line_source(P, h, theta, t, logN, z, bk, li);
            const double wrest = P.line_wrest[li];
            lA[t] = C_KMS / bk;
            lC[t] = wrest * (1.0 + z);
            lK[t] = TAU_CONST * exp10(logN) * P.line_f[li] * (wrest * 1e-8) / (bk * 1e5);
            la[t] = P.line_gamma[li] * (wrest * 1e-8) / (4.0 * PI_D * (bk * 1e5));
        }
        int n = 0;
        double sigma = 1.0;
        const bool conv = h.specres > P.velstep;
        if (conv) lsf_geometry(h.specres, P.velstep, sigma, n);
        if (n < 0) n = 0;
        if (tid <= n && tid <= P.nmax) taps[tid] = conv ? exp(-0.5 * (double)(tid * tid) / (sigma * sigma)) : 1.0;
        for (int k = tid + nthreads; k <= n && k <= P.nmax; k += nthreads)
            taps[k] = exp(-0.5 * (double)(k * k) / (sigma * sigma));
        __syncthreads();
        const double ninv2s2 = -0.5 / (sigma * sigma);
        auto tap = [&](int k) { return k <= P.nmax ? taps[k] : exp((double)k * (double)k * ninv2s2); };
        double norm = taps[0];
        for (int k = 1; k <= n; ++k) norm += 2.0 * tap(k);
        norm = 1.0 / norm;

        for (int i = tid; i < P.npix; i += nthreads) {
            const double lam = P.wave[i];
            double tau = 0.0;
            for (int t = 0; t < h.nact; ++t) {
                const double u = lA[t] * (lC[t] - lam) / lam;
                tau += lK[t] * voigt_h64(la[t], u);
            }
            flux[i] = exp(-tau);
        }
        __syncthreads();
        double acc = 0.0;
        int cnt5 = 0, cnt4 = 0;
        for (int i = tid; i < P.npix; i += nthreads) {
            double m = taps[0] * flux[i];
            for (int k = 1; k <= n; ++k) {
                int ip = (i + k) % P.npix, im = (i - k) % P.npix;
                if (im < 0) im += P.npix;
                m += tap(k) * (flux[ip] + flux[im]);
            }
            m = m * norm * h.cont;
            if (Bt.flux_out) {
                if (flags & MCALF_F_FLUX_F64) ((double *)Bt.flux_out)[b * (long long)P.npix + i] = m;
                else ((float *)Bt.flux_out)[b * (long long)P.npix + i] = (float)m;
            }
            const double r = P.obj[i] - m;
            acc += P.w[i] * r * r;
            if (P.asymmlike) {
                const double rs = (P.obj_raw[i] - m) * P.isig[i];
                cnt5 += rs > 5.0;
                cnt4 += rs > 4.0;
            }
        }
        acc = warp_sum(acc);
        cnt5 = warp_sum_int(cnt5);
        cnt4 = warp_sum_int(cnt4);
        if (lane == 0) { red[warp] = acc; ((int *)(red + 32))[warp] = cnt5; ((int *)(red + 32))[32 + warp] = cnt4; }
        __syncthreads();
        if (warp == 0) {
            double v = warp_sum(lane < nwarps ? red[lane] : 0.0);
            const int c5 = warp_sum_int(lane < nwarps ? ((int *)(red + 32))[lane] : 0);
            const int c4 = warp_sum_int(lane < nwarps ? ((int *)(red + 32))[32 + lane] : 0);
            if (lane == 0) {
                double logl = P.logC - 0.5 * v;
                if (P.asymmlike && ((double)c5 > P.asym_t5 || (double)c4 > P.asym_t4)) logl = -INFINITY;
                store_results(Bt, b, logl, v + P.chi2_add);
            }
        }
    }
}

// unit cube -> physical parameters (hires_fitter.py:202-216)
__global__ void mcalf_prior_kernel(const __grid_constant__ DevProblem P, const double *cube, long long B, long long ld,
                                   uint32_t flags, double *out) {
    const long long n = B * (long long)P.ndim;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / P.ndim;
        const int k = (int)(i - b * P.ndim);
        out[b * (long long)P.ndim + k] = load_theta(P, cube + b * ld, k, flags | MCALF_F_UNIT_CUBE);
    }
}

// element-wise Re w(u + i a) with the kernels' own device functions (unit tests)
__global__ void mcalf_voigt_h_kernel(int mode, const double *u, const double *a, long long n, double *out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    {
        if (mode == 1) { out[i] = voigt_h64(a[i], u[i]); continue; }
        const float af = (float)a[i], uf = (float)u[i];
        // mode 0: wing form / two-float core form; mode 2: wing form / the short core form of weak lines;
        // mode 3: the weak-line forms with the core boundary drawn in to s = S_WIDE
        if (mode == 3) { out[i] = (double)voigt_h32_weak(af, uf, MCALF_S_WIDE); continue; }
        if (mode == 2 && fma32(uf, uf, af * af) < S_CUT) out[i] = (double)core_h32_lean(af, af * af, uf);
        else out[i] = (double)voigt_h32(af, uf);
    }
}

// FFMA-only loop: the measured FP32 peak the roofline fraction is also quoted against
__global__ void mcalf_ffma_peak_kernel(float *out, int iters) {
    float x0 = threadIdx.x * 1e-9f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
          x7 = x0 + 7.f;
    const float a = 0.999999f, c = 1e-7f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            x0 = fmaf(x0, a, c); x1 = fmaf(x1, a, c); x2 = fmaf(x2, a, c); x3 = fmaf(x3, a, c);
            x4 = fmaf(x4, a, c); x5 = fmaf(x5, a, c); x6 = fmaf(x6, a, c); x7 = fmaf(x7, a, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

}  // namespace mcalf

// ---------------------------------------------------------------------------------------------
// launchers (called by mcalf_api.cu)
// ---------------------------------------------------------------------------------------------
namespace mcalf {

SmemLayout fast_smem_layout(const DevProblem &P) { return make_layout(P); }

size_t fp64_smem_bytes(const DevProblem &P) {
    return sizeof(double) * ((size_t)P.ndim_pad + 4 * (size_t)P.Lmax + P.nmax + 1 + 64 + P.npix);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is held per function and per device, not per context: raise it
// to what the device allows (opt-in limit minus the kernel's static shared memory) and never to one problem's
// size, so contexts of different sizes on one device cannot lower each other's limit.  Idempotent.
template <typename K>
static cudaError_t raise_smem_limit(K kernel, size_t optin, size_t *static_bytes) {
    cudaFuncAttributes at;
    cudaError_t e = cudaFuncGetAttributes(&at, kernel);
    if (e != cudaSuccess) return e;
    if (static_bytes) *static_bytes = at.sharedSizeBytes;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(optin - at.sharedSizeBytes));
}

cudaError_t configure_kernels(size_t optin_bytes, size_t *fast_static_bytes) {
    cudaError_t e = raise_smem_limit(mcalf_fast_kernel<false, false, false, false>, optin_bytes, fast_static_bytes);
    if (e != cudaSuccess) return e;
    e = raise_smem_limit(mcalf_fast_kernel<false, false, true, false>, optin_bytes, nullptr);
    if (e != cudaSuccess) return e;
    e = raise_smem_limit(mcalf_fast_kernel<false, false, false, true>, optin_bytes, nullptr);
    if (e != cudaSuccess) return e;
    e = raise_smem_limit(mcalf_fast_kernel<false, false, true, true>, optin_bytes, nullptr);
    if (e != cudaSuccess) return e;
    e = raise_smem_limit(mcalf_fast_kernel<false, true, false, false>, optin_bytes, nullptr);
    if (e != cudaSuccess) return e;
    e = raise_smem_limit(mcalf_fast_kernel<true, true, false, false>, optin_bytes, nullptr);
    if (e != cudaSuccess) return e;
    return raise_smem_limit(mcalf_fp64_kernel, optin_bytes, nullptr);
}

cudaError_t fast_occupancy(int threads, size_t smem, int dense, int *ctas_per_sm) {
    if (dense) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, mcalf_fast_kernel<false, false, true, false>, threads, smem);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, mcalf_fast_kernel<false, false, false, false>, threads, smem);
}

cudaError_t launch_fast(const DevProblem &P, const BatchArgs &Bt, int grid, int threads, size_t smem, int dense, cudaStream_t st) {
    const bool one_each = (long long)grid >= Bt.B;          // one CTA per sample: no work queue
    const bool d = dense && threads <= 256;
    if (Bt.stats) mcalf_fast_kernel<true, true, false, false><<<grid, threads, smem, st>>>(P, Bt);
    else if (Bt.flux_out != nullptr || P.asymmlike) mcalf_fast_kernel<false, true, false, false><<<grid, threads, smem, st>>>(P, Bt);
    else if (one_each && d) mcalf_fast_kernel<false, false, true, true><<<grid, threads, smem, st>>>(P, Bt);
    else if (one_each) mcalf_fast_kernel<false, false, false, true><<<grid, threads, smem, st>>>(P, Bt);
    else if (d) mcalf_fast_kernel<false, false, true, false><<<grid, threads, smem, st>>>(P, Bt);
    else mcalf_fast_kernel<false, false, false, false><<<grid, threads, smem, st>>>(P, Bt);
    return cudaGetLastError();
}

cudaError_t launch_fp64(const DevProblem &P, const BatchArgs &Bt, const int *idx_list, const unsigned int *idx_count, int grid,
                        size_t smem, cudaStream_t st) {
    mcalf_fp64_kernel<<<grid, 256, smem, st>>>(P, Bt, idx_list, idx_count);
    return cudaGetLastError();
}

cudaError_t launch_prior(const DevProblem &P, const double *cube, long long B, long long ld, uint32_t flags, double *out,
                         cudaStream_t st) {
    long long n = B * (long long)P.ndim;
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) grid = 1;
    mcalf_prior_kernel<<<grid, 256, 0, st>>>(P, cube, B, ld, flags, out);
    return cudaGetLastError();
}

cudaError_t launch_voigt_h(int mode, const double *u, const double *a, long long n, double *out, cudaStream_t st) {
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) grid = 1;
    mcalf_voigt_h_kernel<<<grid, 256, 0, st>>>(mode, u, a, n, out);
    return cudaGetLastError();
}

// checked build: read and clear the device-side violation mask (0 in the product build)
cudaError_t check_flag_fetch(unsigned int *mask) {
    *mask = 0u;
#if defined(MCALF_CHECK)
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return e;
    e = cudaMemcpyFromSymbol(mask, mcalf_check_flag, sizeof(unsigned int));
    if (e != cudaSuccess) return e;
    const unsigned int zero = 0u;
    return cudaMemcpyToSymbol(mcalf_check_flag, &zero, sizeof(unsigned int));
#else
    return cudaSuccess;
#endif
}

cudaError_t launch_ffma_peak(float *out, int grid, int threads, int iters, cudaStream_t st) {
    mcalf_ffma_peak_kernel<<<grid, threads, 0, st>>>(out, iters);
    return cudaGetLastError();
}

}  // namespace mcalf
