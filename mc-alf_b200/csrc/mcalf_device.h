// mcalf_device.h -- structures shared by the kernels (mcalf_kernels.cu) and the C-ABI host layer
// (mcalf_api.cu).  Internal: the public interface is include/mcalf_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mcalf_b200.h"
#include "host_setup.h"

namespace mcalf {

// byte offsets of the fp32 kernel's dynamic shared-memory arrays (fast_smem_layout)
struct SmemLayout {
    int theta, A64, rc64, lp, uarr, nmask, cmask, farp, taps, flux, red, misc, chk, bytes;
};

// Everything als_fitter.__init__ leaves behind (hires_fitter.py:65-200), in device form.
struct DevProblem {
    int npix, npix4, nchunks, nlines;
    int ncompmax, nfill, ndim, ndim_pad;
    int startind, endind, free_specres, free_cont;
    int asymmlike, halo, nmax, nmax4;
    int Lmax, mwords;                   // mwords: 32-bit words of a per-chunk line mask, ceil(Lmax / 32)
    int scratch_in_flux, check_selftest; // pass-A outputs of a chunk live in its slice of the depth buffer; check_selftest: the
                                        // -DMCALF_CHECK build raises violation bit 31 on purpose (proves the plumbing)
    int cslot_w, cslot_lw, nslots, vwarps; // lanes per chunk group (power of two), its log2, slots = vwarps * 32 / cslot_w
    float eps_cull, eps_far;
    SmemLayout lay;
    int nhalo;                          // halo cells: `halo` before pixel 0, the rest after the last pixel
    double fixed_specres, fixed_cont, velstep, lam_ref;
    double logC, asym_t5, asym_t4, a_max;
    double chi2_add;                    // +inf when a zero-error pixel makes the reference's chi2 infinite, else 0
    const float2 *dhi2, *dlo2;          // [npix + 256] rho_i - rho_s(chunk) as a two-float (hi, lo), each stored as the pixel
                                        // pair {v[i], v[i+32]} a lane holds (host_setup.h: build_pair_table)
    const float4 *obj_hi4, *obj_lo4, *w4; // [npix4/4] flux as a two-float and weight 1/err^2, four pixels per element;
                                        // obj = w = 0 on dropped pixels and on the padding
    const ChunkDesc *chunks;            // [nchunks]
    const int *halo_src;                // [nhalo] source pixel of every halo cell (periodic wrap)
    const double *wave, *obj, *w;       // [npix] fp64 copies for the check kernel (obj = w = 0 on dropped pixels)
    const double *obj_raw, *isig;       // [npix] untouched flux and 1/err for the Asymmlike counts
    const double *line_wrest, *line_f, *line_gamma;   // [nlines + 1], the last entry is the filler line
    const double *blo, *bhi;            // [ndim] prior bounds
};

constexpr int MCALF_MAX_PEERS = 8;

struct BatchArgs {
    const double *params;
    long long B, ld;
    uint32_t flags;
    int npeers;                         // > 0: logL of sample b is ALSO stored to logl_peer[p][b], p < npeers
    double *logl_out, *chi2_out;
    double *logl_peer[MCALF_MAX_PEERS]; // buffers of other GPUs mapped into this device's address space (NVLink peer
                                        // stores at the kernel tail: the logL gather of the sharded path, no collective)
    void *flux_out;
    unsigned int *work_counter;         // zero at launch (cleared by the previous launch, see clear_counters)
    unsigned int *fallback_count;       // zero at launch
    unsigned int *clear_counters;       // [2] the counter pair of the NEXT launch on this slot: this launch zeroes it
    int *fallback_list;                 // [B]
    int *fallback_flag;                 // nullable, mapped host memory: set to 1 when any sample was re-routed
    unsigned long long *stats;          // nullable: {total, wing, mixed, core, culled, far, core-precise} evaluations
};

SmemLayout fast_smem_layout(const DevProblem &P);
size_t fp64_smem_bytes(const DevProblem &P);
cudaError_t configure_kernels(size_t optin_bytes, size_t *fast_static_bytes);
cudaError_t fast_occupancy(int threads, size_t smem, int dense, int *ctas_per_sm);
cudaError_t launch_fast(const DevProblem &P, const BatchArgs &Bt, int grid, int threads, size_t smem, int dense, cudaStream_t st);
cudaError_t launch_fp64(const DevProblem &P, const BatchArgs &Bt, const int *idx_list, const unsigned int *idx_count, int grid,
                        size_t smem, cudaStream_t st);
cudaError_t launch_prior(const DevProblem &P, const double *cube, long long B, long long ld, uint32_t flags, double *out,
                         cudaStream_t st);
cudaError_t launch_voigt_h(int mode, const double *u, const double *a, long long n, double *out, cudaStream_t st);
cudaError_t launch_ffma_peak(float *out, int grid, int threads, int iters, cudaStream_t st);
cudaError_t check_flag_fetch(unsigned int *mask);   // -DMCALF_CHECK build: synchronise, read and clear the bounds-violation mask

}  // namespace mcalf
