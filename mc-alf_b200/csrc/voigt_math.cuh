// voigt_math.cuh -- the arithmetic of the MC-ALF likelihood hot path, written once for device and
// host.  The host build exists only so tests/ can exercise these exact functions without a GPU
// (tests/host_emul); the product never computes on the host.
//
// Reference being replaced: als_fitter.voigt_tau (mcalf/routines/hires_fitter.py:331-367), whose
// kernel is  tau = 0.014971475 * N * f * Re w(u + i a) / dnu  with w from scipy.special.wofz.
//
//   fp32 fast path   H(a,u) = Re w(u+ia) split at s = u^2 + a^2 = scut, a per-line boundary
//                    (S_CUT = 36 for strong lines; drawn in to max(16, ln kappa + 17.5) for weak lines,
//                    whose Gaussian part kappa exp(-s) is below 2.5e-8 beyond it -- line_cut):
//       wing  (s >= scut):   H = (a/sqrt pi) * q * P(q),  q = 1/s, P a degree-4 polynomial
//                            (1 MUFU.RCP + 6 FMA-pipe ops; fitted over s >= 36 to 6e-8 for strong lines,
//                            over s >= 16 to 5e-6 for weak lines where 5e-6 * tau_wing <= 1e-8)
//       core  (s <  scut):   Taylor series in a about the Gaussian (Harris 1948):
//                            H = G0 (1 + a^2 (1-2x)) + a (G1 (1 + a^2 (1 - 2x/3)) + 2 a^2/(3 sqrt pi)),
//                            x = u^2, G0 = exp(-x) from a two-float x, G1 = H1(u) from a Taylor table.
//                            Valid for a <= A_MAX_FAST; larger a is routed to the fp64 path.
//   fp64 check path  trapezoid rule with pole correction on one of two staggered grids
//                    (Hunter & Regan 1972), asymptotic series for s > 100; ~1e-13 relative.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MCALF_HD __host__ __device__ __forceinline__
#define MCALF_HD_NOINLINE __host__ __device__ __noinline__
#else
#define MCALF_HD inline
#define MCALF_HD_NOINLINE inline
#endif

#include "voigt_tables.inc"

// Bounds-checked build (-DMCALF_CHECK, libmcalf_b200_check.so): every shared-memory / global index the fp32
// kernel forms is validated; a violation sets a bit of a device-side flag that the C-ABI turns into
// MCALF_E_CUDA ("bounds check failed").  compute-sanitizer is closed on the GPU pool, so this is the memory-
// safety tool of the test suite (tests/test_gpu_checked.py).  In the product build the macro is empty.
#if defined(MCALF_CHECK) && defined(__CUDACC__)
namespace mcalf { static __device__ unsigned int mcalf_check_flag = 0u; }
#endif
#if defined(MCALF_CHECK) && defined(__CUDA_ARCH__)
#define MCALF_CHK(cond, bit) do { if (!(cond)) atomicOr(&::mcalf::mcalf_check_flag, 1u << (bit)); } while (0)
#else
#define MCALF_CHK(cond, bit) ((void)0)
#endif
// bits: 0 theta  1 line table  2 masks  3 pass-A scratch  4 taps  5 pixel pair tables  6 depth store  7 halo
//       8 stencil window  9 per-pixel tables  10 fallback list  11 flux output  12 H1 table  13 chunk table

namespace mcalf {

constexpr double C_KMS = 2.9979245e5;        // hires_fitter.py:65
constexpr double TAU_CONST = 0.014971475;    // hires_fitter.py:364
constexpr double FWHM_TO_SIGMA = 2.354820;   // hires_fitter.py:454
constexpr double TRUNC_SIGMAS = 3.0348;      // hires_fitter.py:458
constexpr double PI_D = 3.14159265358979323846;
constexpr double SQRTPI_D = 1.77245385090551602730;
constexpr float S_CUT = MCALF_S_CUT;
constexpr double A_MAX_FAST = 0.02;          // beyond this the a^4 term of the core series matters

struct alignas(16) G1Row { float c0, c1, c2, c3; };

#if defined(__CUDACC__)
static __device__ const G1Row g1_tab_dev[MCALF_G1_N] = { MCALF_G1_ROWS };
#endif
static const G1Row g1_tab_host[MCALF_G1_N] = { MCALF_G1_ROWS };

MCALF_HD float fma32(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}

// Packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: fma|mul|add.rn.f32x2, sm_100+).  One instruction
// does two IEEE-rounded fp32 operations -- the same bits as two scalar operations -- at half the warp
// issue rate, which is what an issue-bound kernel needs: the FP work occupies half the issue slots.
struct alignas(8) F2 { float x, y; };
MCALF_HD F2 f2(float a, float b) { F2 r; r.x = a; r.y = b; return r; }
MCALF_HD F2 f2(float a) { F2 r; r.x = a; r.y = a; return r; }
#if defined(__CUDA_ARCH__)
#define MCALF_PK(v) (*reinterpret_cast<const unsigned long long *>(&(v)))
MCALF_HD F2 fma2(F2 a, F2 b, F2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(MCALF_PK(a)), "l"(MCALF_PK(b)), "l"(MCALF_PK(c)));
    return *reinterpret_cast<F2 *>(&d);
}
MCALF_HD F2 mul2(F2 a, F2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(MCALF_PK(a)), "l"(MCALF_PK(b)));
    return *reinterpret_cast<F2 *>(&d);
}
MCALF_HD F2 add2(F2 a, F2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(MCALF_PK(a)), "l"(MCALF_PK(b)));
    return *reinterpret_cast<F2 *>(&d);
}
#else
MCALF_HD F2 fma2(F2 a, F2 b, F2 c) { return f2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
MCALF_HD F2 mul2(F2 a, F2 b) { return f2(a.x * b.x, a.y * b.y); }
MCALF_HD F2 add2(F2 a, F2 b) { return f2(a.x + b.x, a.y + b.y); }
#endif

// Round-to-nearest-integer of a * b (0 <= a*b < 2^22) without the conversion unit: adding 1.5 * 2^23
// leaves the integer in the low mantissa bits.  n: the integer, return value: the same as a float.
constexpr float RN_MAGIC = 12582912.0f;
MCALF_HD float rn_mul(float a, float b, int &n) {
    const float t = fma32(a, b, RN_MAGIC);
    union { float f; int32_t i; } v;
    v.f = t;
    n = v.i - 0x4B400000;
    return t - RN_MAGIC;
}

MCALF_HD float rcp32(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}

// q*P(q) with q = 1/s: the wing shape.  Multiply by c1 = kappa*a/sqrt(pi) to get tau.
MCALF_HD float wing_qp(float s) {
    const float w[5] = MCALF_WING_P;
    float q = rcp32(s);
    float p = fma32(w[4], q, w[3]);
    p = fma32(p, q, w[2]);
    p = fma32(p, q, w[1]);
    p = fma32(p, q, w[0]);
    return p * q;
}

// exp(-(x + xlo)) for x >= 0, |xlo| << 1: Cody-Waite reduction, degree-7 polynomial, exact 2^-n scale.
MCALF_HD float exp_neg32(float x, float xlo) {
    const float c[8] = MCALF_EXPM_C;
    x = fminf(x, 88.0f);
    float n = rintf(x * 1.44269504088896341f);
    float r = fma32(n, -0.693145751953125f, x);
    r = fma32(n, -1.42860676533018702e-06f, r);
    r += xlo;
    float p = fma32(c[7], r, c[6]);
    p = fma32(p, r, c[5]);
    p = fma32(p, r, c[4]);
    p = fma32(p, r, c[3]);
    p = fma32(p, r, c[2]);
    p = fma32(p, r, c[1]);
    p = fma32(p, r, c[0]);
    int e = 127 - (int)n;               // n in [0,127]
    union { int32_t i; float f; } sc;
    sc.i = e << 23;
    return p * sc.f;
}

// 1 - exp(-x) for x >= 0 (the absorption depth of optical depth x), relative accuracy ~1e-7 even for
// x -> 0: with exp(-x) = 2^-n (1 + r q(r)) the depth is (1 - 2^-n) - 2^-n r q(r), and 1 - 2^-n is exact.
MCALF_HD float depth32(float x) {
    const float c[8] = MCALF_EXPM_C;
    x = fminf(x, 88.0f);
    int ni;
    float n = rn_mul(x, 1.44269504088896341f, ni);
    float r = fma32(n, -0.693145751953125f, x);
    r = fma32(n, -1.42860676533018702e-06f, r);
    float q = fma32(c[7], r, c[6]);
    q = fma32(q, r, c[5]);
    q = fma32(q, r, c[4]);
    q = fma32(q, r, c[3]);
    q = fma32(q, r, c[2]);
    q = fma32(q, r, c[1]);
    union { int32_t i; float f; } sc;
    sc.i = (127 - ni) << 23;            // n in [0,127]
    return fma32(-sc.f, r * q, 1.0f - sc.f);
}

// depth32 on a pair of optical depths (packed where both lanes run the same operation)
MCALF_HD F2 depth32_2(F2 x) {
    const float c[8] = MCALF_EXPM_C;
    x = f2(fminf(x.x, 88.0f), fminf(x.y, 88.0f));
    const F2 tm = fma2(x, f2(1.44269504088896341f), f2(RN_MAGIC));      // rn_mul, packed
    const F2 n = add2(tm, f2(-RN_MAGIC));
    F2 r = fma2(n, f2(-0.693145751953125f), x);
    r = fma2(n, f2(-1.42860676533018702e-06f), r);
    F2 q = fma2(f2(c[7]), r, f2(c[6]));
    q = fma2(q, r, f2(c[5]));
    q = fma2(q, r, f2(c[4]));
    q = fma2(q, r, f2(c[3]));
    q = fma2(q, r, f2(c[2]));
    q = fma2(q, r, f2(c[1]));
    union { int32_t i; float f; } sa, sb;
    union { float f; int32_t i; } ta, tb;
    ta.f = tm.x;
    tb.f = tm.y;
    sa.i = (127 - (ta.i - 0x4B400000)) << 23;    // n in [0,127]
    sb.i = (127 - (tb.i - 0x4B400000)) << 23;
    const F2 sc = f2(sa.f, sb.f);
    const F2 msc = f2(-sa.f, -sb.f);
    return fma2(msc, mul2(r, q), add2(f2(1.0f), msc));
}

// tab: the table's copy in shared memory (kernels), or null for the global/host copy
MCALF_HD G1Row g1_row(int j, const G1Row *tab = nullptr) {
    MCALF_CHK(j >= 0 && j < MCALF_G1_N, 12);
#if defined(MCALF_CHECK)
    j = j < 0 ? 0 : (j >= MCALF_G1_N ? MCALF_G1_N - 1 : j);
#endif
#if defined(__CUDA_ARCH__)
    const float4 v = tab ? reinterpret_cast<const float4 *>(tab)[j] : __ldg(reinterpret_cast<const float4 *>(g1_tab_dev) + j);
    G1Row r; r.c0 = v.x; r.c1 = v.y; r.c2 = v.z; r.c3 = v.w;
    return r;
#else
    (void)tab;
    return g1_tab_host[j];
#endif
}

// Line-core H(a,u) for u given as a two-float (uh + ul), s = u^2+a^2 < S_CUT, a <= A_MAX_FAST.
MCALF_HD float core_h32(float a, float a2, float uh, float ul, const G1Row *tab = nullptr) {
    float x = uh * uh;
    float xlo = fma32(uh, uh, -x) + 2.0f * uh * ul;
    float g0 = exp_neg32(x, xlo);
    float au = fabsf(uh);
    float sl = (uh < 0.0f) ? -ul : ul;
    float fj = rintf(au * (float)MCALF_G1_INV_H);
    int j = (int)fj;
    j = j < MCALF_G1_N - 1 ? j : MCALF_G1_N - 1;
    float d = fma32(fj, -1.0f / (float)MCALF_G1_INV_H, au) + sl;
    G1Row t = g1_row(j, tab);
    float g1 = fma32(fma32(fma32(t.c3, d, t.c2), d, t.c1), d, t.c0);
    float k0 = fma32(a2, fma32(-2.0f, x, 1.0f), 1.0f);
    float k1 = fma32(a2, fma32(-0.666666687f, x, 1.0f), 1.0f);
    float inner = fma32(g1, k1, a2 * 0.376126389f);
    return fma32(g0, k0, a * inner);
}

MCALF_HD float ex2_32(float t) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return r;
#else
    return exp2f(t);
#endif
}

// Line-core H(a,u) for WEAK lines (kappa <= KAPPA_LEAN): u as one float accurate to ~1e-7, the
// Gaussian through MUFU.EX2 (relative error 2^-22).  The optical-depth error is <= ~3e-7 tau, i.e. a
// flux error <= 1.5e-7 for kappa <= KAPPA_LEAN (the flux sensitivity F |dtau| peaks near tau = 1).
constexpr float KAPPA_LEAN = 8.0f;
constexpr float U_TAB_END = MCALF_G1_UMAX + 0.03f;   // last H1 table row
constexpr float U_TAB = MCALF_G1_UMAX;               // |u| up to which the core forms are valid
MCALF_HD float core_h32_lean(float a, float a2, float u, const G1Row *tab = nullptr) {
    const float x = u * u;
    const float g0 = ex2_32(x * -1.44269504088896341f);
    const float au = fabsf(u);
    int j;
    const float fj = rn_mul(fminf(au, U_TAB_END), (float)MCALF_G1_INV_H, j);   // (the table ends at U_TAB_END)
    const float d = fma32(fj, -1.0f / (float)MCALF_G1_INV_H, au);
    const G1Row t = g1_row(j, tab);
    const float g1 = fma32(fma32(fma32(t.c3, d, t.c2), d, t.c1), d, t.c0);
    const float k0 = fma32(a2, fma32(-2.0f, x, 1.0f), 1.0f);
    const float k1 = fma32(a2, fma32(-0.666666687f, x, 1.0f), 1.0f);
    const float inner = fma32(g1, k1, a2 * 0.376126389f);
    return fma32(g0, k0, a * inner);
}

// The short core form on a pair of pixels (same arithmetic as core_h32_lean, packed where the two lanes
// of the pair run the same operation; the table look-ups, conversions and MUFU stay scalar).
MCALF_HD F2 core_h32_lean2(float a, float a2, F2 u, const G1Row *tab = nullptr) {
    const F2 x = mul2(u, u);
    const F2 t = mul2(x, f2(-1.44269504088896341f));
    const F2 g0 = f2(ex2_32(t.x), ex2_32(t.y));
    const F2 au = f2(fabsf(u.x), fabsf(u.y));
    const F2 ac = f2(fminf(au.x, U_TAB_END), fminf(au.y, U_TAB_END));        // (the table ends at U_TAB_END)
    const F2 tm = fma2(ac, f2((float)MCALF_G1_INV_H), f2(RN_MAGIC));         // rn_mul, packed
    const F2 fj = add2(tm, f2(-RN_MAGIC));
    union { float f; int32_t i; } ua, ub;
    ua.f = tm.x;
    ub.f = tm.y;
    const int ja = ua.i - 0x4B400000, jb = ub.i - 0x4B400000;
    const F2 d = fma2(fj, f2(-1.0f / (float)MCALF_G1_INV_H), au);
    const G1Row ta = g1_row(ja, tab), tb = g1_row(jb, tab);
    const F2 g1 = f2(fma32(fma32(fma32(ta.c3, d.x, ta.c2), d.x, ta.c1), d.x, ta.c0),
                     fma32(fma32(fma32(tb.c3, d.y, tb.c2), d.y, tb.c1), d.y, tb.c0));
    const F2 a22 = f2(a2), one = f2(1.0f);
    const F2 k0 = fma2(a22, fma2(f2(-2.0f), x, one), one);
    const F2 k1 = fma2(a22, fma2(f2(-0.666666687f), x, one), one);
    const F2 inner = fma2(g1, k1, f2(a2 * 0.376126389f));
    return fma2(g0, k0, mul2(f2(a), inner));
}

// Convenience scalar forms of the fast path (unit tests, mcalf_voigt_h): u, a as floats.
MCALF_HD float voigt_h32(float a, float u) {
    float a2 = a * a;
    float s = fma32(u, u, a2);
    if (s < S_CUT) return core_h32(a, a2, u, 0.0f);
    return a * 0.564189584f * wing_qp(s);
}
// weak-line forms: wide-interval wing polynomial beyond scut (>= S_WIDE), short core form inside
MCALF_HD float voigt_h32_weak(float a, float u, float scut) {
    const float w[5] = MCALF_WING_PW;
    float a2 = a * a;
    float s = fma32(u, u, a2);
    if (s < scut) return core_h32_lean(a, a2, u);
    float q = rcp32(s);
    float p = fma32(w[4], q, w[3]);
    p = fma32(p, q, w[2]);
    p = fma32(p, q, w[1]);
    p = fma32(p, q, w[0]);
    return a * 0.564189584f * (p * q);
}

// ---------------------------------------------------------------------------------------------
// fp64 check path
// ---------------------------------------------------------------------------------------------
MCALF_HD double voigt_h64(double a, double u) {
    const double e0[MCALF_TRAP_N] = MCALF_TRAP_E0;
    const double e1[MCALF_TRAP_N] = MCALF_TRAP_E1;
    const double h = MCALF_TRAP_H;
    u = fabs(u);
    const double a2 = a * a;
    const double s = u * u + a2;
    if (s > 100.0 || a > 5.0) {
        // Re[(i/sqrt pi)/z * (1 + sum_k (2k-1)!!/(2 z^2)^k)], z = u + i a, Horner from k = 12
        const double inv = 1.0 / s;
        const double zr = u * inv, zi = -a * inv;             // 1/z
        const double wr = zr * zr - zi * zi, wi = 2.0 * zr * zi;  // 1/z^2
        double ar = 0.0, ai = 0.0;
        for (int k = 12; k >= 1; --k) {
            const double c = 0.5 * (2 * k - 1);
            const double tr = (ar + 1.0) * c, ti = ai * c;
            ar = tr * wr - ti * wi;
            ai = tr * wi + ti * wr;
        }
        // (i/z)(1+acc): i*(zr + i zi) = -zi + i zr
        const double br = -zi, bi = zr;
        return (br * (1.0 + ar) - bi * ai) / SQRTPI_D;
    }
    const double f = u / h - floor(u / h);
    const bool half = fabs(f - 0.5) > 0.25;   // u close to an integer node: use the staggered grid
    double sum = 0.0;
    if (!half) {
        sum = e0[0] / (u * u + a2);
        for (int k = 1; k < MCALF_TRAP_N; ++k) {
            const double t = k * h;
            sum += e0[k] * (1.0 / ((u - t) * (u - t) + a2) + 1.0 / ((u + t) * (u + t) + a2));
        }
    } else {
        for (int k = 0; k < MCALF_TRAP_N; ++k) {
            const double t = (k + 0.5) * h;
            sum += e1[k] * (1.0 / ((u - t) * (u - t) + a2) + 1.0 / ((u + t) * (u + t) + a2));
        }
    }
    double res = h / PI_D * a * sum;
    // pole correction Re[2 exp(-z^2) / (1 - sg exp(-2 pi i z / h))]
    const double sg = half ? -1.0 : 1.0;
    const double ex = exp(a2 - u * u);
    const double c2 = cos(2.0 * u * a), s2 = sin(2.0 * u * a);
    const double E = exp(2.0 * PI_D * a / h);
    const double th = 2.0 * PI_D * (u / h);
    const double dr = 1.0 - sg * E * cos(th), di = sg * E * sin(th);
    res += 2.0 * ex * (c2 * dr - s2 * di) / (dr * dr + di * di);
    return res;
}

// ---------------------------------------------------------------------------------------------
// Per-line set-up (fp64, once per (sample, line)): voigt_tau's scalars in the well-conditioned form
//   u = A (rho - rho_c),  rho = lam_ref/lambda,  rho_c = lam_ref/(wrest (1+z)),  A = (c/b) wrest (1+z)/lam_ref
//   tau = kappa H(a,u),  kappa = 0.014971475 10^logN f wrest_cm / b_cms,  a = gamma wrest_cm / (4 pi b_cms)
// (algebraically identical to hires_fitter.py:355-365).
//
// Pixel coordinates are stored per CHUNK (<= 256 consecutive pixels of one fit window): the context
// keeps rho_s (fp64, the chunk's reference) and delta_i = rho_i - rho_s as a two-float (hi, lo).
// Per (line, chunk) the kernel forms U = A (rho_s - rho_c) in fp64 once, so that per pixel
//   u = fma(A_hi, delta_hi, U_hi)                                   (one FMA; wing accuracy)
//   u = (A_hi, A_lo) * (delta_hi, delta_lo) + (U_hi, U_lo)          (two-float; line core)
// ---------------------------------------------------------------------------------------------
struct Line64 {
    double A, rc, kappa, a;
};

// 64 bytes per active line, four float4 rows: pass A reads row 0, the direct wing form rows 0-2, the
// line-core forms all four.
struct alignas(16) LineP {
    float A_hi, a2, c1, ucm;     // c1 = kappa*a/sqrt(pi): the wing amplitude; ucm: a chunk whose |u| stays above it is wing only
    float cw0, cw1, cw2, cw3;    // c1 * (wing polynomial of this line): tau_wing = q (cw0 + q (cw1 + ...)), q = 1/s
    float cw4, scut, kappa, a;   // scut: the line's core boundary in s = u^2 + a^2
    float A_lo, pad0, pad1, pad2;
};

// 10^x for x in [-300, 300] to ~2e-8 relative: 2^(i + f) with i = rint(x log2 10) and a degree-7 polynomial
// for 2^f on |f| <= 1/2.  kappa multiplies an fp32 H, so fp32-level accuracy is all it needs; the fp64 library
// exp10 was the longest dependent chain of the per-sample set-up (one thread per line, everybody else waiting).
MCALF_HD double pow10_fast(double x) {
    const double t = x * 3.32192809488736234787;
    const double fi = rint(t);
    const double f = t - fi;                               // exact difference of close doubles
    // 2^f = exp(f ln2), Taylor/minimax to degree 7 in fp64 (cheap: 7 FMAs), |error| < 2e-9
    const double g = f * 0.69314718055994530942;
    double p = 1.0 / 5040.0;
    p = fma(p, g, 1.0 / 720.0);
    p = fma(p, g, 1.0 / 120.0);
    p = fma(p, g, 1.0 / 24.0);
    p = fma(p, g, 1.0 / 6.0);
    p = fma(p, g, 0.5);
    p = fma(p, g, 1.0);
    p = fma(p, g, 1.0);
    int e = (int)fi;
    e = e < -1000 ? -1000 : (e > 1000 ? 1000 : e);
#if defined(__CUDA_ARCH__)
    return scalbn(p, e);
#else
    return ldexp(p, e);
#endif
}

MCALF_HD Line64 line_setup64(double logN, double z, double b_kms, double wrest, double f, double gamma,
                             double lam_ref) {
    Line64 L;
    const double lamc = wrest * (1.0 + z);
    const double inv_b = 1.0 / b_kms;                      // the only divisions: 1/b and lam_ref/lamc
    L.rc = lam_ref / lamc;
    L.A = (C_KMS * inv_b) * (lamc / lam_ref);
    const double cold = pow10_fast(logN);
    const double wb = (wrest * 1e-8) * (inv_b * 1e-5);     // wrest_cm / b_cms
    L.kappa = TAU_CONST * cold * f * wb;
    L.a = gamma * wb * (1.0 / (4.0 * PI_D));
    return L;
}

MCALF_HD void split2(double v, float &hi, float &lo) {
    hi = (float)v;
    lo = (float)(v - (double)hi);
}

// Per-line core boundary.  Outside s = scut the wing form drops the Gaussian part of H, an optical-depth
// error of at most kappa exp(-scut): scut = ln(kappa) + ln(1/EPS_GAUSS) keeps it below EPS_GAUSS.  The
// boundary may only be drawn inside S_CUT when the wide-interval wing polynomial is accurate enough
// for this line (its relative error times the wing's optical depth at the boundary <= EPS_WING);
// otherwise the line keeps S_CUT and the narrow-interval polynomial.
constexpr double EPS_GAUSS = 2.5e-8;
constexpr double LN_INV_EPS_GAUSS = 17.504390;     // ln(1 / 2.5e-8)
constexpr double EPS_WING = 1.0e-8;
constexpr float U_CUT_MARGIN = 0.01f;              // covers the one-FMA coordinate's error
constexpr float U_CORE_MARGIN = 6.01f;             // sqrt(S_CUT) + U_CUT_MARGIN

MCALF_HD bool line_cut(double kappa, double a, float &scut) {
    const double c1 = kappa * a / SQRTPI_D;
#if defined(__CUDA_ARCH__)
    double s = (kappa > 0.0) ? (double)__logf((float)kappa) + LN_INV_EPS_GAUSS : (double)MCALF_S_WIDE;   // (the boundary needs no precision)
#else
    double s = (kappa > 0.0) ? (double)logf((float)kappa) + LN_INV_EPS_GAUSS : (double)MCALF_S_WIDE;
#endif
    if (!(s > (double)MCALF_S_WIDE)) s = (double)MCALF_S_WIDE;
    const bool wide = s < (double)S_CUT - 0.5 && c1 * MCALF_WING_PW_ERR <= EPS_WING * s;
    scut = wide ? (float)s : S_CUT;
    return wide;
}

MCALF_HD LineP line_pack_full(const Line64 &L) {
    const float wn[5] = MCALF_WING_P, ww[5] = MCALF_WING_PW;
    LineP o;
    split2(L.A, o.A_hi, o.A_lo);
    o.a = (float)L.a;
    o.a2 = (float)(L.a * L.a);
    o.c1 = (float)(L.kappa * L.a / SQRTPI_D);
    o.kappa = (float)L.kappa;
    const bool wide = line_cut(L.kappa, L.a, o.scut);
    o.ucm = wide ? sqrtf(o.scut) + U_CUT_MARGIN : U_CORE_MARGIN;
    const float *w = wide ? ww : wn;
    o.cw0 = o.c1 * w[0]; o.cw1 = o.c1 * w[1]; o.cw2 = o.c1 * w[2]; o.cw3 = o.c1 * w[3]; o.cw4 = o.c1 * w[4];
    o.pad0 = o.pad1 = o.pad2 = 0.0f;
    return o;
}

// Far-field ("local expansion") of the Lorentzian wings.  For a line whose centre lies far outside a
// chunk, tau(delta) = c1 g(U + A delta), g(u) = 1/u^2 + (3/2 - a^2)/u^4 + (15/4)/u^6 (the asymptotic
// series of sqrt(pi) H/a), is expanded in x = delta/ds in [-1, 1] (ds = max |delta| of the chunk):
//   1/u^(2m) = U^(-2m) sum_n binom(-2m, n) (r x)^n,   r = A ds / U,  |r| <= FF_RMAX,
// truncated at degree FF_DEG.  All far lines of a chunk are summed into ONE polynomial, evaluated once
// per pixel.  A (line, chunk) pair takes this form only if the proven error bound
//   c1/umin^2 [ 2 (FF_DEG+2) r^(FF_DEG+1)/(1-FF_RMAX)^2 + 14/umin^6 ] <= eps_far
// holds (umin = |U| - A ds, the closest approach), so strong or nearby lines stay on the direct form.
constexpr int FF_DEG = 7;
constexpr float FF_RMAX = 0.45f;
constexpr float FF_UMIN = 10.0f;
constexpr float FF_TRUNC = 2.0f * (FF_DEG + 2) / ((1.0f - 0.45f) * (1.0f - 0.45f));

// Class of a (line, chunk) pair: 0 = culled, 1 = wing only, 2 = mixed (some pixel may have
// u^2 + a^2 < scut, i.e. |u| < ucm), 3 = far field.  ds = max |delta| over the chunk.
MCALF_HD int chunk_class(float A_hi, float U_hi, float ds, float c1, float ucm, float eps_cull, float eps_far) {
    const float hw = A_hi * ds;                  // half-width of the chunk in u
    const float aU = fabsf(U_hi);
    const float umin = aU - hw;
    if (!(umin > ucm)) return 2;                 // also catches NaN
    const float um2 = umin * umin;
    // tau <= c1 q P(q) <= 1.05 c1 / u^2 on the chunk
    if (1.05f * c1 < eps_cull * um2) return 0;
    const float r = hw * rcp32(aU);
    if (r <= FF_RMAX && umin >= FF_UMIN) {
        const float r2 = r * r;
        const float iu2 = rcp32(um2);
        float rp = r2 * r2;                          // r^(FF_DEG + 1), FF_DEG odd
#pragma unroll
        for (int k = 4; k < FF_DEG + 1; k += 2) rp *= r2;
        const float err = c1 * iu2 * fma32(FF_TRUNC, rp, 14.0f * iu2 * iu2 * iu2);
        if (err <= eps_far) return 3;
    }
    return 1;
}

// the binomial pairs {b[2m], b[2m+1]} of the three series: constant-bank operands on the device (the compiler would
// otherwise rebuild the twelve 64-bit constants with two UMOVs each on every trip of the classification loop)
#if defined(__CUDACC__)
static __constant__ float2 ff_b1_dev[5] = {{1.f, 2.f}, {3.f, 4.f}, {5.f, 6.f}, {7.f, 8.f}, {9.f, 10.f}};
static __constant__ float2 ff_b2_dev[5] = {{1.f, 4.f}, {10.f, 20.f}, {35.f, 56.f}, {84.f, 120.f}, {165.f, 220.f}};
static __constant__ float2 ff_b3_dev[5] = {{1.f, 6.f}, {21.f, 56.f}, {126.f, 252.f}, {462.f, 792.f}, {1287.f, 2002.f}};
#endif
#if defined(__CUDA_ARCH__)
#define FF_B1(m) f2(ff_b1_dev[m].x, ff_b1_dev[m].y)
#define FF_B2(m) f2(ff_b2_dev[m].x, ff_b2_dev[m].y)
#define FF_B3(m) f2(ff_b3_dev[m].x, ff_b3_dev[m].y)
#else
static const float ff_b1_host[10] = {1.f, 2.f, 3.f, 4.f, 5.f, 6.f, 7.f, 8.f, 9.f, 10.f};
static const float ff_b2_host[10] = {1.f, 4.f, 10.f, 20.f, 35.f, 56.f, 84.f, 120.f, 165.f, 220.f};
static const float ff_b3_host[10] = {1.f, 6.f, 21.f, 56.f, 126.f, 252.f, 462.f, 792.f, 1287.f, 2002.f};
#define FF_B1(m) f2(ff_b1_host[2 * (m)], ff_b1_host[2 * (m) + 1])
#define FF_B2(m) f2(ff_b2_host[2 * (m)], ff_b2_host[2 * (m) + 1])
#define FF_B3(m) f2(ff_b3_host[2 * (m)], ff_b3_host[2 * (m) + 1])
#endif

// Add one far line's expansion coefficients (in x = delta/ds) to C[0..FF_DEG], held as pairs
// {C[2m], C[2m+1]} (packed arithmetic: two coefficients per instruction).
MCALF_HD void farfield_accumulate(float A_hi, float U_hi, float ds, float c1, float a2, F2 *C2) {
    static_assert(FF_DEG % 2 == 1 && FF_DEG <= 9, "coefficient pairs; chunk_class computes r^(FF_DEG+1) by squaring; tables hold 10 binomials");
    const float iU = rcp32(U_hi);
    const float s = -(A_hi * ds) * iU;           // -r (signed)
    const float v = iU * iU;
    const float T1 = c1 * v;
    const float T2 = T1 * v * (1.5f - a2);
    const float T3 = T1 * v * v * 3.75f;
    // binom(-2,n) = (-1)^n (n+1), binom(-4,n) = (-1)^n C(n+3,3), binom(-6,n) = (-1)^n C(n+5,5)
    const F2 T12 = f2(T1), T22 = f2(T2), T32 = f2(T3), ss = f2(s * s);
    F2 sn = f2(1.0f, s);
#pragma unroll
    for (int m = 0; 2 * m <= FF_DEG; ++m) {
        const F2 t = fma2(FF_B1(m), T12, fma2(FF_B2(m), T22, mul2(FF_B3(m), T32)));
        C2[m] = fma2(sn, t, C2[m]);
        sn = mul2(sn, ss);
    }
}

MCALF_HD float farfield_eval(const float *C, float x) {
    float p = C[FF_DEG];
#pragma unroll
    for (int n = FF_DEG - 1; n >= 0; --n) p = fma32(p, x, C[n]);
    return p;
}

// The direct wing form of one line on a pair of pixels: q (cw0 + q (cw1 + q (cw2 + q (cw3 + q cw4)))), q = 1/s,
// with the line's pre-scaled coefficients.  s >= scut (guaranteed by the caller's classification, or clamped).
MCALF_HD F2 wing_val2(F2 s, const LineP &L, F2 &q) {
    q = f2(rcp32(s.x), rcp32(s.y));
    F2 p = fma2(f2(L.cw4), q, f2(L.cw3));
    p = fma2(p, q, f2(L.cw2));
    p = fma2(p, q, f2(L.cw1));
    p = fma2(p, q, f2(L.cw0));
    return p;
}
// tau + wing (the last multiply fused with the accumulation)
MCALF_HD F2 wing_acc2(F2 tau, F2 s, const LineP &L) {
    F2 q;
    const F2 p = wing_val2(s, L, q);
    return fma2(q, p, tau);
}

// two-float u = A*delta + U for the line core
MCALF_HD void core_u2(float A_hi, float A_lo, float d_hi, float d_lo, float U_hi, float U_lo, float &uh,
                      float &ul) {
    const float ph = A_hi * d_hi;
    float pl = fma32(A_hi, d_hi, -ph);
    pl = fma32(A_hi, d_lo, pl);
    pl = fma32(A_lo, d_hi, pl);
    const float sh = U_hi + ph;
    const float bb = sh - U_hi;
    const float se = (U_hi - (sh - bb)) + (ph - bb);
    uh = sh;
    ul = se + (pl + U_lo);
}

// strong lines (rare): one out-of-line copy, so that the unrolled synthesis loop stays small
MCALF_HD_NOINLINE float core_precise(float A_hi, float A_lo, float d_hi, float d_lo, float U_hi, float U_lo, float a, float a2,
                                     const G1Row *tab) {
    float uh, ul;
    core_u2(A_hi, A_lo, d_hi, d_lo, U_hi, U_lo, uh, ul);
    return core_h32(a, a2, uh, ul, tab);
}

// kappa H(a,u) of one line on a pair of pixels with the line-core forms (valid for |u| <= U_TAB): the short
// form for weak lines, the two-float form for strong ones.  u = fma(A_hi, d_hi, U_hi) is passed in.
MCALF_HD F2 core_val2(const LineP &L, F2 u, F2 dh, F2 dl, float U_hi, float U_lo, const G1Row *tab = nullptr) {
    F2 h;
    if (L.kappa <= KAPPA_LEAN) {
        const F2 uc = add2(u, fma2(f2(L.A_hi), dl, fma2(f2(L.A_lo), dh, f2(U_lo))));
        h = core_h32_lean2(L.a, L.a2, uc, tab);
    } else {
        h.x = core_precise(L.A_hi, L.A_lo, dh.x, dl.x, U_hi, U_lo, L.a, L.a2, tab);
        h.y = core_precise(L.A_hi, L.A_lo, dh.y, dl.y, U_hi, U_lo, L.a, L.a2, tab);
    }
    return mul2(f2(L.kappa), h);
}

// What the kernel does with one mixed line on one pair of 32-pixel rows (the 64 lanes' values are passed as
// arrays by the host emulation; the kernel votes across the warp): all pixels beyond the core boundary ->
// wing form; all pixels inside the H1 table -> core form (valid on both sides of the boundary); a pair that
// straddles both -> both forms and a per-pixel select.  Returns 0 / 1 / 2 for the statistics.
enum { PAIR_WING = 0, PAIR_CORE = 1, PAIR_BOTH = 2 };
MCALF_HD int pair_kind(bool all_wing, bool all_tab) { return all_wing ? PAIR_WING : (all_tab ? PAIR_CORE : PAIR_BOTH); }

// a row pair that straddles the table end and the core boundary (small b; rare): both forms, per-pixel select
// (measured out of line to shrink the unrolled synthesis loop: the call costs more than the code size saves)
MCALF_HD F2 straddle_pair_tau(F2 tau, F2 kh, F2 s, float scut, float cw0, float cw1, float cw2, float cw3, float cw4) {
    LineP L;
    L.cw0 = cw0; L.cw1 = cw1; L.cw2 = cw2; L.cw3 = cw3; L.cw4 = cw4;
    F2 q;
    const F2 sc = f2(fmaxf(s.x, scut), fmaxf(s.y, scut));
    const F2 p = wing_val2(sc, L, q);
    const F2 w = mul2(q, p);
    return add2(tau, f2(s.x < scut ? kh.x : w.x, s.y < scut ? kh.y : w.y));
}

MCALF_HD F2 mixed_pair_tau(int kind, F2 tau, const LineP &L, F2 u, F2 s, F2 dh, F2 dl, float U_hi, float U_lo,
                           const G1Row *tab = nullptr) {
    if (kind == PAIR_WING) return wing_acc2(tau, s, L);
    const F2 kh = core_val2(L, u, dh, dl, U_hi, U_lo, tab);
    if (kind == PAIR_CORE) return add2(tau, kh);
    return straddle_pair_tau(tau, kh, s, L.scut, L.cw0, L.cw1, L.cw2, L.cw3, L.cw4);
}

// LSF geometry (hires_fitter.py:452-459): sigma in pixels and half-width n = ceil(3.0348 sigma).
MCALF_HD void lsf_geometry(double fwhm, double velstep, double &sigma_px, int &n) {
    sigma_px = (fwhm / FWHM_TO_SIGMA) / velstep;
    n = (int)ceil(TRUNC_SIGMAS * sigma_px);
}

}  // namespace mcalf
