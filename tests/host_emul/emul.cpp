// TEST INFRASTRUCTURE: host build of the device arithmetic in mc-alf_b200/csrc/voigt_math.cuh and of
// the per-sample algorithm of mcalf_fast_kernel (same functions, same order of operations), so the
// numerics can be checked against scipy / the oracle without a GPU.  Never linked into the product
// library and never used to produce a result the product returns.
#include <math.h>

#include <vector>

#include "host_setup.h"
#include "voigt_math.cuh"

using namespace mcalf;

extern "C" {

void emul_voigt_h32(long n, const double *a, const double *u, double *out) {
    for (long i = 0; i < n; ++i) out[i] = (double)voigt_h32((float)a[i], (float)u[i]);
}

void emul_voigt_h64(long n, const double *a, const double *u, double *out) {
    for (long i = 0; i < n; ++i) out[i] = voigt_h64(a[i], u[i]);
}

void emul_exp_neg32(long n, const double *x, double *out) {
    for (long i = 0; i < n; ++i) out[i] = (double)exp_neg32((float)x[i], 0.0f);
}

void emul_depth32(long n, const double *x, double *out) {
    for (long i = 0; i < n; ++i) out[i] = (double)depth32((float)x[i]);
}

void emul_lsf_geometry(double fwhm, double velstep, double *sigma_px, int *n) { lsf_geometry(fwhm, velstep, *sigma_px, *n); }

int emul_num_chunks(long npix, const double *wave) {
    std::vector<ChunkDesc> chunks;
    std::vector<float> dhi, dlo;
    build_chunks(wave, (int)npix, wave[npix / 2], chunks, dhi, dlo);
    return (int)chunks.size();
}

// Optical depth of `nlines` lines over the pixel grid exactly as mcalf_fast_kernel accumulates it
// (chunk classification, one-FMA wing coordinate, two-float core coordinate).  lines: rows of
// (logN, z, b_kms, wrest, f, gamma).  cls_out (nullable): [nchunks*nlines] class of each pair.
void emul_tau(long npix, const double *wave, int nlines, const double *lines, double eps_cull, double eps_far,
              double *tau_out, int *cls_out) {
    const double lam_ref = wave[npix / 2];
    std::vector<ChunkDesc> chunks;
    std::vector<float> dhi, dlo;
    build_chunks(wave, (int)npix, lam_ref, chunks, dhi, dlo);
    std::vector<float> tau(npix, 0.0f);
    for (size_t c = 0; c < chunks.size(); ++c) {
        const ChunkDesc &cd = chunks[c];
        // the kernel evaluates the far-field polynomial first, then the wing-only lines, then the mixed ones
        F2 C2[(FF_DEG + 1) / 2] = {};
        int nf = 0;
        for (int pass = 0; pass <= 2; ++pass) {
            for (int t = 0; t < nlines; ++t) {
                const double *l = lines + 6 * t;
                const Line64 L64 = line_setup64(l[0], l[1], l[2], l[3], l[4], l[5], lam_ref);
                const LineP L = line_pack_full(L64);
                const double U = L64.A * (cd.rho_s - L64.rc);
                float Uh, Ul;
                split2(U, Uh, Ul);
                const int cls = chunk_class(L.A_hi, Uh, cd.ds, L.c1, (float)eps_cull, (float)eps_far);
                if (cls_out && pass == 0) cls_out[c * nlines + t] = cls;
                if (pass == 0) {
                    if (cls == 3) { farfield_accumulate(L.A_hi, Uh, cd.ds, L.c1, L.a2, C2); ++nf; }
                    continue;
                }
                if (cls != pass) continue;
                // both classes: direct wing form with s clamped at S_CUT; class 2 then replaces the
                // clamped value by the core form where s < S_CUT (as the kernel's core pass does)
                const float c1w = wing_tau(L.c1, S_CUT);
                for (int i = cd.start; i < cd.start + cd.len; ++i) {
                    const float u = fma32(L.A_hi, dhi[i], Uh);
                    const float s = fma32(u, u, L.a2);
                    tau[i] += wing_tau(L.c1, fmaxf(s, S_CUT));
                    if (cls == 2 && s < S_CUT) {
                        float h;
                        if (L.kappa <= KAPPA_LEAN) {
                            const float uc = u + fma32(L.A_hi, dlo[i], fma32(L.A_lo, dhi[i], Ul));
                            h = core_h32_lean(L.a, L.a2, uc);
                        } else {
                            float uh, ul;
                            core_u2(L.A_hi, L.A_lo, dhi[i], dlo[i], Uh, Ul, uh, ul);
                            h = core_h32(L.a, L.a2, uh, ul);
                        }
                        tau[i] += fma32(L.kappa, h, -c1w);
                    }
                }
            }
            if (pass == 0 && nf) {
                float C[FF_DEG + 1];
                for (int m = 0; m < (FF_DEG + 1) / 2; ++m) { C[2 * m] = C2[m].x; C[2 * m + 1] = C2[m].y; }
                for (int i = cd.start; i < cd.start + cd.len; ++i) tau[i] = farfield_eval(C, dhi[i] * cd.inv_ds);
            }
        }
    }
    for (long i = 0; i < npix; ++i) tau_out[i] = (double)tau[i];
}

// Depth -> LSF stencil -> model and chi-square with the kernel's fp32 arithmetic.
// pix: obj (0 where dropped), w (0 where dropped).  Returns chi2; model_out[npix] as double.
double emul_epilogue(long npix, const double *tau, const double *obj, const double *w, double specres, double velstep,
                     double cont, double *model_out) {
    int n = 0;
    double sigma = 1.0;
    const bool conv = specres > velstep;
    if (conv) lsf_geometry(specres, velstep, sigma, n);
    const double inv2s2 = conv ? 0.5 / (sigma * sigma) : 0.0;
    double norm = 0.0;
    for (int k = 0; k <= n; ++k) norm += (k == 0 ? 1.0 : 2.0) * exp(-(double)(k * k) * inv2s2);
    norm = 1.0 / norm;
    std::vector<float> g(2 * n + 1), dep(npix);
    for (int k = -n; k <= n; ++k) g[k + n] = (float)(exp(-(double)(k * k) * inv2s2) * norm);
    for (long i = 0; i < npix; ++i) dep[i] = depth32((float)tau[i]);
    const float c_hi = (float)cont, c_lo = (float)(cont - (double)c_hi);
    double chi2 = 0.0;
    for (long i = 0; i < npix; ++i) {
        float s = 0.0f;
        for (int k = -n; k <= n; ++k) {
            long j = (i + k) % npix;
            if (j < 0) j += npix;
            s = fma32(g[k + n], dep[j], s);
        }
        const float oh = (float)obj[i], ol = (float)(obj[i] - (double)oh);
        const float base = (oh - c_hi) + (ol - c_lo);
        const float res = fma32(c_hi, s, base) + c_lo * s;
        chi2 += (double)((float)w[i] * res * res);
        model_out[i] = cont - cont * (double)s;
    }
    return chi2;
}

}  // extern "C"
