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

void emul_voigt_h32_weak(long n, const double *a, const double *u, double scut, double *out) {
    for (long i = 0; i < n; ++i) out[i] = (double)voigt_h32_weak((float)a[i], (float)u[i], (float)scut);
}

// scut / wide flag the kernel picks for a line of given kappa, a
int emul_line_cut(double kappa, double a, double *scut) {
    float sc;
    const bool wide = line_cut(kappa, a, sc);
    *scut = (double)sc;
    return wide ? 1 : 0;
}

// Optical depth of `nlines` lines over the pixel grid exactly as mcalf_fast_kernel accumulates it:
// per chunk the far-field polynomial first, then the near lines in line order -- wing-only lines on the
// direct wing form, lines whose core may reach the chunk row pair by row pair (64 pixels: the warp's vote
// is a loop over the pair's 64 slots here, junk slots beyond a short chunk included, as in the kernel).
// lines: rows of (logN, z, b_kms, wrest, f, gamma).  cls_out (nullable): [nchunks*nlines] class of each
// pair.  kind_out (nullable): [4] row pairs taken as wing / core / straddling, and wing-only line-chunks.
void emul_tau(long npix, const double *wave, int nlines, const double *lines, double eps_cull, double eps_far,
              double *tau_out, int *cls_out, long *kind_out) {
    const double lam_ref = wave[npix / 2];
    std::vector<ChunkDesc> chunks;
    std::vector<float> dhi, dlo;
    build_chunks(wave, (int)npix, lam_ref, chunks, dhi, dlo);
    std::vector<PairF> ph, pl;
    build_pair_table(chunks, dhi, ph);
    build_pair_table(chunks, dlo, pl);
    std::vector<float> tau(npix, 0.0f);
    if (kind_out) kind_out[0] = kind_out[1] = kind_out[2] = kind_out[3] = 0;
    for (size_t c = 0; c < chunks.size(); ++c) {
        const ChunkDesc &cd = chunks[c];
        // registers of the warp: tau[j][lane] as pairs
        F2 T[4][32], D[4][32], DL[4][32];
        for (int j = 0; j < 4; ++j)
            for (int l = 0; l < 32; ++l) {
                const PairF v = ph[cd.start + 64 * j + l], w = pl[cd.start + 64 * j + l];
                D[j][l] = f2(v.x, v.y);
                DL[j][l] = f2(w.x, w.y);
                T[j][l] = f2(0.0f);
            }
        // far lines
        F2 C2[(FF_DEG + 1) / 2] = {};
        std::vector<int> cls(nlines);
        std::vector<LineP> lp(nlines);
        std::vector<Line64> l64(nlines);
        for (int t = 0; t < nlines; ++t) {
            const double *l = lines + 6 * t;
            l64[t] = line_setup64(l[0], l[1], l[2], l[3], l[4], l[5], lam_ref);
            lp[t] = line_pack_full(l64[t]);
            const float Uh = (float)(l64[t].A * (cd.rho_s - l64[t].rc));
            cls[t] = chunk_class(lp[t].A_hi, Uh, cd.ds, lp[t].c1, lp[t].ucm, (float)eps_cull, (float)eps_far);
            if (cls_out) cls_out[c * nlines + t] = cls[t];
            if (cls[t] == 3) farfield_accumulate(lp[t].A_hi, Uh, cd.ds, lp[t].c1, lp[t].a2, C2);
        }
        {
            float C[FF_DEG + 1];
            for (int m = 0; m < (FF_DEG + 1) / 2; ++m) { C[2 * m] = C2[m].x; C[2 * m + 1] = C2[m].y; }
            for (int j = 0; j < 4; ++j)
                for (int l = 0; l < 32; ++l)
                    T[j][l] = f2(farfield_eval(C, D[j][l].x * cd.inv_ds), farfield_eval(C, D[j][l].y * cd.inv_ds));
        }
        for (int t = 0; t < nlines; ++t) {
            if (cls[t] != 1 && cls[t] != 2) continue;
            const LineP &L = lp[t];
            float Uh, Ul;
            split2(l64[t].A * (cd.rho_s - l64[t].rc), Uh, Ul);
            if (cls[t] == 1 && kind_out) kind_out[3] += 1;
            for (int j = 0; j < 4; ++j) {
                F2 u[32], s[32];
                bool all_wing = true, all_tab = true;
                for (int l = 0; l < 32; ++l) {
                    u[l] = fma2(f2(L.A_hi), D[j][l], f2(Uh));
                    s[l] = fma2(u[l], u[l], f2(L.a2));
                    all_wing = all_wing && (s[l].x >= L.scut && s[l].y >= L.scut);
                    all_tab = all_tab && (fabsf(u[l].x) <= U_TAB && fabsf(u[l].y) <= U_TAB);
                }
                const int kind = cls[t] == 1 ? PAIR_WING : pair_kind(all_wing, all_tab);
                if (cls[t] == 2 && kind_out) kind_out[kind] += 1;
                for (int l = 0; l < 32; ++l)
                    T[j][l] = mixed_pair_tau(kind, T[j][l], L, u[l], s[l], D[j][l], DL[j][l], Uh, Ul);
            }
        }
        for (int j = 0; j < 4; ++j)
            for (int l = 0; l < 32; ++l) {
                const int k = 64 * j + l;
                if (k < cd.len) tau[cd.start + k] = T[j][l].x;
                if (k + 32 < cd.len) tau[cd.start + k + 32] = T[j][l].y;
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
