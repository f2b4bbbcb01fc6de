// TEST INFRASTRUCTURE: host build of mc-alf_b200/csrc/voigt_math.cuh so the kernels' arithmetic can
// be checked against scipy.special.wofz without a GPU.  Never linked into the product library.
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

// tau of one line over a pixel grid exactly as the fp32 kernel computes it.
// mode 0: near form everywhere (two-float coordinate + core fix-up); mode 1: far form everywhere
// (one FMA, wing only); mode 2: the kernel's own per-256-pixel-segment choice.
// d_near_out receives the line's near/far switch distance in rho units.
void emul_line_tau(long npix, const double *wave, double lam_ref, double logN, double z, double b, double wrest,
                   double f, double gamma, int mode, double eps_far, double *tau_out, double *u_out,
                   double *d_near_out) {
    Line64 L64 = line_setup64(logN, z, b, wrest, f, gamma, lam_ref);
    Line32 L = line_setup32(L64, eps_far, 0.0);
    *d_near_out = L.d_near;
    for (long s0 = 0; s0 < npix; s0 += 256) {
        long s1 = s0 + 256 < npix ? s0 + 256 : npix;
        float rmin = 3e38f, rmax = -3e38f;
        for (long i = s0; i < s1; ++i) {
            float r = (float)(lam_ref / wave[i]);
            rmin = fminf(rmin, r);
            rmax = fmaxf(rmax, r);
        }
        float dist = fmaxf(fmaxf(rmin - L.rc_hi, L.rc_hi - rmax), 0.0f);
        bool far = mode == 1 || (mode == 2 && dist > L.d_near);
        for (long i = s0; i < s1; ++i) {
            float hi, lo;
            split2(lam_ref / wave[i], hi, lo);
            float tau;
            if (far) {
                float u = fma32(L.A_hi, hi, L.U0);
                float s = fma32(u, u, L.a2);
                s = fmaxf(s, S_CUT);
                tau = L.c1 * wing_qp(s);
                u_out[i] = u;
            } else {
                float dh = hi - L.rc_hi, dl = lo - L.rc_lo;
                float t = dh + dl;
                float u = L.A_hi * t;
                float s = fma32(u, u, L.a2);
                float sc = fmaxf(s, S_CUT);
                tau = L.c1 * wing_qp(sc);
                u_out[i] = u;
                if (s < S_CUT) {
                    float e = dl - (t - dh);
                    float ul = fma32(L.A_hi, t, -u) + fma32(L.A_hi, e, L.A_lo * t);
                    tau += fma32(L.kappa, core_h32(L.a, L.a2, u, ul), -L.c1w);
                }
            }
            tau_out[i] = tau;
        }
    }
}

void emul_lsf_geometry(double fwhm, double velstep, double *sigma_px, int *n) { lsf_geometry(fwhm, velstep, *sigma_px, *n); }
}
