// fs2_math.cuh -- device-side fp64 building blocks of the filter step (sm_100a).
//
// Every function names the reference lines it computes (cy-rae/fast-slam).  The gate test is written
// with explicit round-to-nearest intrinsics in exactly the operation order of the CPU oracle
// (oracle/fs2_oracle.c: fs2o_mahalanobis) so that association decisions are bit-identical on identical
// landmark state; everything else is ordinary fp64 (FMA contraction allowed) and agrees to rounding.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FS2_PI 3.141592653589793
#define FS2_TWO_PI 6.283185307179586
#define FS2_LOG_2PI 1.8378770664093453

struct Fs2Lm {  // one landmark: mean and un-symmetrised 2x2 covariance (landmark.py:13-21)
    double x, y, c00, c01, c10, c11;
};

// (a + pi) % (2 pi) - pi with Python/numpy floored-modulo semantics (fast_slam_2.py:84-85, :125).
// fmod is exact, and for |quotient| <= 1 so are the shortcuts below, hence bit-identical to the oracle.
__device__ __forceinline__ double fs2_wrap_pi(double a)
{
    const double v = __dadd_rn(a, FS2_PI);
    // the three cases with |quotient| <= 1 as selects (no branch on the usual path), fmod behind one rare branch
    const bool in0 = (v >= 0.0) && (v < FS2_TWO_PI);
    const bool in1 = (v >= FS2_TWO_PI) && (v < 2.0 * FS2_TWO_PI);
    const bool in2 = (v < 0.0) && (v > -FS2_TWO_PI);
    double m = v;
    m = in1 ? __dadd_rn(v, -FS2_TWO_PI) : m;      // exact (Sterbenz)
    m = in2 ? __dadd_rn(v, FS2_TWO_PI) : m;       // fmod(v) == v, then the reference's "mod += b"
    if (!(in0 || in1 || in2)) {
        m = fmod(v, FS2_TWO_PI);
        if (m != 0.0 && m < 0.0) m = __dadd_rn(m, FS2_TWO_PI);
    }
    if (m == 0.0) m = 0.0;  // copysign(0, 2 pi)
    return __dadd_rn(m, -FS2_PI);
}

// inverse covariance of a landmark as the oracle computes it: adj/det with one reciprocal.
struct Fs2Gate {
    double i00, i01, i10, i11;
    bool singular;  // det == 0: np.linalg.inv raises (geometry_utils.py:22)
};

__device__ __forceinline__ Fs2Gate fs2_gate_prepare(double c00, double c01, double c10, double c11)
{
    Fs2Gate g;
    double det = __dadd_rn(__dmul_rn(c00, c11), -__dmul_rn(c01, c10));
    g.singular = (det == 0.0);
    double r = __ddiv_rn(1.0, det);
    g.i00 = __dmul_rn(c11, r);
    g.i01 = __dmul_rn(-c01, r);
    g.i10 = __dmul_rn(-c10, r);
    g.i11 = __dmul_rn(c00, r);
    return g;
}

// GeometryUtils.mahalanobis_distance(landmark, observed, cov) < gate  (geometry_utils.py:14-23,
// landmark_utils.py:106-115).  delta = observed - landmark.  NaN (negative form) compares false.
__device__ __forceinline__ bool fs2_gate_test(const Fs2Gate &g, double lx, double ly, double ox, double oy,
                                              double gate)
{
    double dx = __dadd_rn(ox, -lx), dy = __dadd_rn(oy, -ly);
    double t0 = __dadd_rn(__dmul_rn(dx, g.i00), __dmul_rn(dy, g.i10));
    double t1 = __dadd_rn(__dmul_rn(dx, g.i01), __dmul_rn(dy, g.i11));
    double s = __dadd_rn(__dmul_rn(t0, dx), __dmul_rn(t1, dy));
    return __dsqrt_rn(s) < gate;
}

// Conservative fp32 screen for the gate.  For a positive definite covariance Cauchy-Schwarz gives
//   dx^2 <= c00 * d^2  and  dy^2 <= c11 * d^2,
// so d < gate implies |dx| < gate*sqrt(c00) and |dy| < gate*sqrt(c11).  The half-widths are widened for
// every fp32 rounding on the way (conversion of the means and observations, approximate sqrt, the
// subtraction) and for the fp64 rounding of the exact test.  The bound needs a covariance that is safely
// positive definite, (numerically) symmetric and not absurdly anisotropic -- then the exact test's own
// rounding error is < 1e-6 relative; anything else gets an infinite box so that the exact test alone
// decides.  A landmark outside the box can never pass the exact test; one inside is re-tested exactly.
struct Fs2Box {
    float mx, my, rx, ry;
};

__device__ __forceinline__ float fs2_sqrt_approx(float v)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ Fs2Box fs2_box(double xd, double yd, double c00, double c01, double c10, double c11,
                                         float gate_f, float slack)
{
    Fs2Box bx;
    const float x = (float)xd, y = (float)yd;
    const float a = (float)c00, b = (float)c01, c = (float)c10, d = (float)c11;
    const float ad = a * d;
    const float det = fmaf(-b, c, ad);          // fp32 error <= 4e-7 * ad for |bc| <= ad
    const float asym = b - c;
    // a > 0 and ad > 1e-30 imply d > 0; fmax(a,d)^2 < 1e4*ad bounds the anisotropy by 1e4
    const float big = fmaxf(a, d);
    const bool finite_pos = fmaxf(fabsf(x), fabsf(y)) < 1e30f;
    const bool safe = (a > 0.f) && (ad > 1e-30f) && (ad < 1e30f) && (det > 4e-6f * ad) &&
                      (asym * asym <= 1e-12f * ad) && (big * big < 1e4f * ad) && finite_pos;
    if (safe) {
        // 1e-5 relative covers sqrt.approx (2^-22), the conversions and the exact test's rounding;
        // 2.4e-7*|m| + slack covers |m - (float)m| + |o - (float)o| + the rounding of the difference
        bx.mx = x;
        bx.my = y;
        bx.rx = fmaf(gate_f * 1.00001f, fs2_sqrt_approx(a), fmaf(fabsf(x), 2.4e-7f, slack));
        bx.ry = fmaf(gate_f * 1.00001f, fs2_sqrt_approx(d), fmaf(fabsf(y), 2.4e-7f, slack));
    } else {
        // infinite box around a finite centre: every finite observation is a candidate, NaN/inf
        // observations still compare false (and fail the exact test too)
        bx.mx = finite_pos ? x : 0.f;
        bx.my = finite_pos ? y : 0.f;
        bx.rx = __int_as_float(0x7f800000);
        bx.ry = bx.rx;
    }
    return bx;
}

// ---- fp64 atan2 and exp for the EKF ------------------------------------------------------------------------------
// libm's versions are accurate to ~1 ulp and so are these, but inlined libm code materialises every polynomial
// coefficient with two moves into uniform registers and drags its special-case paths along: half of the appliers'
// instructions.  The coefficients below live in constant memory (an FMA reads them in place), the polynomials are
// split into two interleaved chains (the appliers are latency-bound), and only what the EKF can produce is handled
// in line.  Fitted to atan and exp at 60 digits (Chebyshev interpolation): max relative error 3.1e-16 / 2.3e-16.
__constant__ double fs2_atan_c[19] = {
    -0.33333333333333315, 0.19999999999986409, -0.14285714284072867, 0.11111111032045885, -0.090909070674963899,
    0.076922759050147713, -0.06666332509211019, 0.058798654898029809, -0.052495185843376632, 0.047052091530712652,
    -0.041652118023505588, 0.035360661072633817, -0.027590832895856107, 0.018805224043642339, -0.01059534604369583,
    0.0046421639380263531, -0.0014627303082646208, 0.00029221739308086863, -2.7633793313601145e-05};
__constant__ double fs2_exp_c[10] = {
    0.50000000000000011, 0.16666666666666669, 0.041666666666624136, 0.0083333333333300633, 0.001388888891720967,
    0.00019841269863050506, 2.4801521302241082e-05, 2.7557268464831627e-06, 2.7620085455018414e-07, 2.5100383195867021e-08};

// atan2(y, x), result in (-pi, pi].  atan(t) = t + t^3 P(t^2) on [0, 1] after folding into the first octant.
__device__ __forceinline__ double fs2_atan2(double y, double x)
{
    const double ax = fabs(x), ay = fabs(y);
    const double mx = fmax(ax, ay), mn = fmin(ax, ay);
    double t = mn / mx;
    if (mx == 0.0) t = 0.0;                                          // atan2(+-0, +-0): angle 0 (or pi below)
    if (mx > 1.7976931348623157e308) t = (mn > 1.7976931348623157e308) ? 1.0 : 0.0;   // infinities, as IEEE atan2 has them
    const double s = t * t, s2 = s * s;
    const double *c = fs2_atan_c;
    double pe = c[18], po = c[17];                                   // even and odd coefficients as two chains in s^2
    pe = fma(pe, s2, c[16]); po = fma(po, s2, c[15]);
    pe = fma(pe, s2, c[14]); po = fma(po, s2, c[13]);
    pe = fma(pe, s2, c[12]); po = fma(po, s2, c[11]);
    pe = fma(pe, s2, c[10]); po = fma(po, s2, c[9]);
    pe = fma(pe, s2, c[8]);  po = fma(po, s2, c[7]);
    pe = fma(pe, s2, c[6]);  po = fma(po, s2, c[5]);
    pe = fma(pe, s2, c[4]);  po = fma(po, s2, c[3]);
    pe = fma(pe, s2, c[2]);  po = fma(po, s2, c[1]);
    pe = fma(pe, s2, c[0]);
    const double p = fma(po, s, pe);
    double r = fma(t * s, p, t);
    if (ay > ax) r = (1.5707963267948966 - r) + 6.123233995736766e-17;
    if (__double2hiint(x) < 0) r = (3.1415926535897931 - r) + 1.2246467991473532e-16;    // sign bit: -0.0 counts as negative
    return copysign(r, y);
}

// exp(x) for -700 < x <= 0 (a normal result); the caller takes libm's exp otherwise
__device__ __forceinline__ double fs2_exp_neg(double x)
{
    const int k = __double2int_rn(x * 1.4426950408889634);
    const double kd = (double)k;
    double y = fma(kd, -0.69314716756343842, x);                     // ln 2, upper bits: exact product
    y = fma(kd, -1.2996506893889889e-08, y);
    const double y2 = y * y;
    const double *c = fs2_exp_c;
    double pe = c[8], po = c[9];
    pe = fma(pe, y2, c[6]); po = fma(po, y2, c[7]);
    pe = fma(pe, y2, c[4]); po = fma(po, y2, c[5]);
    pe = fma(pe, y2, c[2]); po = fma(po, y2, c[3]);
    pe = fma(pe, y2, c[0]); po = fma(po, y2, c[1]);
    const double p = fma(po, y, pe);
    const double r = fma(y2, p, y) + 1.0;
    return r * __hiloint2double((k + 1023) << 20, 0);                // 2^k, k in [-1010, 0]
}

// scipy.stats.multivariate_normal.pdf(nu, 0, Q) for 2x2 (fast_slam_2.py:156), same restatement as
// oracle/fs2_oracle.c: fs2o_mvn_pdf2 (lower triangle, _PSD rejection rule).  false = scipy raises.
__device__ __forceinline__ bool fs2_mvn_pdf2(double n0, double n1, double q00, double q10, double q11, double *out)
{
    if (!(isfinite(q00) && isfinite(q10) && isfinite(q11) && isfinite(n0) && isfinite(n1))) return false;
    const double hm = 0.5 * (q00 + q11);
    const double det = q00 * q11 - q10 * q10;               // = l0 * l1
    // scipy's acceptance rule is l0 > 1e6*eps*l1 (eigenvalues l0 <= l1).  For a PSD matrix hm <= l1 <= 2 hm,
    // so det > 1e6*eps*(2 hm)^2 implies it without computing the eigenvalues; only matrices within a factor
    // of four of the singularity threshold (or indefinite ones) take the exact route.
    const double lim = 1e6 * 2.220446049250313e-16;
    if (!(hm > 0.0 && det > 4.0 * lim * hm * hm)) {
        const double hd = 0.5 * (q00 - q11);
        const double rad = sqrt(hd * hd + q10 * q10);
        double l0 = hm - rad;
        const double l1 = hm + rad;
        if (l1 > 0.0) l0 = det / l1;
        const double eps = lim * fmax(fabs(l0), fabs(l1));
        if (l0 < -eps) return false;
        if (!(l0 > eps)) return false;
    }
    const double rdet = 1.0 / det;
    const double maha = (q11 * n0 * n0 - 2.0 * q10 * n0 * n1 + q00 * n1 * n1) * rdet;
    // exp(-0.5 * (2 log 2pi + log det + maha)) = exp(-maha / 2) / (2 pi sqrt(det)); equal to rounding -- as long as
    // exp(-maha / 2) itself is a normal number.  Past maha ~ 1417 it is not, while scipy's single exp of the whole
    // exponent still is when det is small: shift the exponent by 350 there and undo it after the product.
    if (maha >= 0.0 && maha < 1390.0)
        *out = fs2_exp_neg(-0.5 * maha) * sqrt(rdet) * 0.15915494309189535;
    else if (maha > 1400.0)
        *out = (exp(350.0 - 0.5 * maha) * sqrt(rdet) * 0.15915494309189535) * 9.92959039626498e-153;   // * exp(-350)
    else
        *out = exp(-0.5 * maha) * sqrt(rdet) * 0.15915494309189535;
    return true;
}

// EKF branch of __update_particle (fast_slam_2.py:116-159) for one associated landmark.
// Returns status bits (FS2_ST_*).  On FS2_ST_SINGULAR_Q nothing changes (*out = in, *like = 1).
// On FS2_ST_PDF_FAILED the landmark is replaced and *like = 1 (weight untouched).
__device__ __forceinline__ int fs2_ekf(double px, double py, double pyaw, double zd, double za,
                                       double r00, double r01, double r10, double r11, const Fs2Lm &in,
                                       Fs2Lm *out, double *like)
{
    const double s00 = in.c00, s01 = in.c01, s10 = in.c10, s11 = in.c11;
    double dx = in.x - px, dy = in.y - py;                       // :116-117
    double q = dx * dx + dy * dy;                                // :118
    const double rd = rsqrt(q);                                  // 1/dist: one reciprocal square root serves
    double dist = q * rd;                                        // :119  dist, 1/dist and 1/q (to rounding)
    if (!(q > 0.0)) dist = sqrt(q);                              // q == 0 / NaN: keep IEEE behaviour (0, NaN)
    // The bearing (a long polynomial chain) and the matrix part (H, H S, Q: another chain) do not depend on each other:
    // they are written back to back, without a branch in between, so that the instruction scheduler interleaves them --
    // the appliers are latency-bound, one warp per particle.
    const double ang_raw = fs2_atan2(dy, dx);                    // :120
    const double rq2 = rd * rd;                                  // 1/q
    double h00 = dx * rd, h01 = dy * rd, h10 = -dy * rq2, h11 = dx * rq2;   // :130-133
    double a00 = h00 * s00 + h01 * s10, a01 = h00 * s01 + h01 * s11;        // H S
    double a10 = h10 * s00 + h11 * s10, a11 = h10 * s01 + h11 * s11;
    double q00 = a00 * h00 + a01 * h01 + r00, q01 = a00 * h10 + a01 * h11 + r01;   // :137
    double q10 = a10 * h00 + a11 * h01 + r10, q11 = a10 * h10 + a11 * h11 + r11;
    double detq = q00 * q11 - q01 * q10;
    double ang = ang_raw - pyaw;                                 // :120
    double n0 = zd - dist;                                       // :124
    double n1 = fs2_wrap_pi(za - ang);                           // :125
    *like = 1.0;
    if (detq == 0.0) {                                           // np.linalg.inv raises at :142
        *out = in;
        return 2;  // FS2_ST_SINGULAR_Q
    }
    double rq = 1.0 / detq;
    double v00 = q11 * rq, v01 = -q01 * rq, v10 = -q10 * rq, v11 = q00 * rq;
    double b00 = s00 * h00 + s01 * h01, b01 = s00 * h10 + s01 * h11;        // S H^T
    double b10 = s10 * h00 + s11 * h01, b11 = s10 * h10 + s11 * h11;
    double k00 = b00 * v00 + b01 * v10, k01 = b00 * v01 + b01 * v11;        // :142
    double k10 = b10 * v00 + b11 * v10, k11 = b10 * v01 + b11 * v11;
    out->x = in.x + (k00 * n0 + k01 * n1);                       // :145
    out->y = in.y + (k10 * n0 + k11 * n1);
    double g00 = 1.0 - (k00 * h00 + k01 * h10), g01 = 0.0 - (k00 * h01 + k01 * h11);   // :146
    double g10 = 0.0 - (k10 * h00 + k11 * h10), g11 = 1.0 - (k10 * h01 + k11 * h11);
    out->c00 = g00 * s00 + g01 * s10;
    out->c01 = g00 * s01 + g01 * s11;
    out->c10 = g10 * s00 + g11 * s10;
    out->c11 = g10 * s01 + g11 * s11;
    double l;
    if (!fs2_mvn_pdf2(n0, n1, q00, q10, q11, &l)) return 4;     // FS2_ST_PDF_FAILED
    *like = l;
    return 0;
}

// New-landmark branch (fast_slam_2.py:108-111, landmark.py:13)
__device__ __forceinline__ Fs2Lm fs2_new_landmark(double px, double py, double pyaw, double zd, double za)
{
    Fs2Lm l;
    double s, c;
    sincos(pyaw + za, &s, &c);
    l.x = __dadd_rn(px, __dmul_rn(zd, c));
    l.y = __dadd_rn(py, __dmul_rn(zd, s));
    l.c00 = 0.1; l.c01 = 0.0; l.c10 = 0.0; l.c11 = 0.1;
    return l;
}

// __move_particle (fast_slam_2.py:69-87, quirk Q12)
__device__ __forceinline__ void fs2_move(double &x, double &y, double &yaw, double rotation, double translation,
                                         double noise, double *sin_yaw = nullptr, double *cos_yaw = nullptr)
{
    double nt, nr;
    if (rotation != 0.0) {
        nt = 0.0;
        nr = __dadd_rn(rotation, noise);
    } else {
        nt = __dadd_rn(translation, noise);
        nr = 0.0;
    }
    double a = fs2_wrap_pi(__dadd_rn(yaw, nr));
    double s, c;
    // |a| <= pi after the wrap: sincospi has exact range reduction and no local-memory slow path (sincos's
    // Payne-Hanek fallback both costs registers and trips ptxas inside a setmaxnreg-reduced region)
    sincospi(a * 0.31830988618379067154, &s, &c);
    yaw = a;
    x = __dadd_rn(x, __dmul_rn(nt, c));
    y = __dadd_rn(y, __dmul_rn(nt, s));
    if (sin_yaw) { *sin_yaw = s; *cos_yaw = c; }
}

// Philox4x32-10 (Salmon et al. 2011), the counter-based generator behind fs2_draw_noise.
__device__ __forceinline__ void fs2_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// one standard normal from a 128-bit Philox block: Box-Muller on a 53-bit and a 32-bit uniform
__device__ __forceinline__ double fs2_normal_from_bits(const uint32_t r[4])
{
    uint64_t m = (((uint64_t)r[0] << 32) | r[1]) >> 11;            // 53 bits
    double u1 = ((double)m + 0.5) * (1.0 / 9007199254740992.0);     // (0,1)
    double u2 = ((double)r[2] + 0.5) * (1.0 / 4294967296.0);
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    return sqrt(-2.0 * log(u1)) * c;
}
