/*
 * fs2_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, fp64, closed-form 2x2 algebra) of the reference's FastSLAM filter step,
 * cy-rae/fast-slam `fast_slam_2/algorithms/fast_slam_2.py` + `utils/landmark_utils.py:92-117` +
 * `utils/geometry_utils.py:14-23`.  It is the checker for the CUDA path and the "port" CPU baseline of
 * bench.py; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product (fast_slam_b200/, libfs2.so) never does.
 *
 * Parity status: PINNED BY EXECUTION.  The reference ships no tests or golden vectors (SURVEY.md
 * section 4); this file is validated against the unmodified reference executed in the build container
 * (oracle/ref_harness.py, driven by oracle/gen_golden.py) and against the frozen outputs of those
 * runs in tests/golden/ (tests/test_oracle_golden.py).  Third-party arithmetic the reference calls
 * (numpy.linalg.inv, scipy.stats.multivariate_normal.pdf, libm via numpy/math) is restated in closed
 * form; agreement is to rounding (<= 1e-11 relative observed), association / resampling indices exact.
 *
 * Layout (identical to the HBM store of libfs2.so so states can be memcmp'd):
 *   x, y, yaw, w : double[P]         count : int32[P]
 *   lm           : double[P][lcap][6]  48 B per landmark:  x, y, c00, c01, c10, c11
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off  (no FMA contraction: the gate test below
 * is replicated operation by operation with __dmul_rn/__dadd_rn/__ddiv_rn/__dsqrt_rn on the device).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FS2O_PI 3.141592653589793
#define FS2O_TWO_PI 6.283185307179586
#define FS2O_LOG_2PI 1.8378770664093453 /* np.log(2*np.pi) as scipy's _LOG_2PI */

/* per-particle status bits (reference: exceptions swallowed by the thread pool, fast_slam_2.py:45,53) */
#define FS2O_ST_SINGULAR_LM 1  /* np.linalg.inv raised inside associate: whole update skipped          */
#define FS2O_ST_SINGULAR_Q 2   /* np.linalg.inv(observation_cov) raised: update skipped                 */
#define FS2O_ST_PDF_FAILED 4   /* multivariate_normal.pdf raised: landmark replaced, weight untouched   */
#define FS2O_ST_MAP_FULL 8     /* not in the reference (lists are unbounded): capacity reached, append dropped */

int fs2o_abi_version(void) { return 1; }

int fs2o_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py pins the thread count itself: launchers such as torchrun export OMP_NUM_THREADS=1 */
int fs2o_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* (a + pi) % (2*pi) - pi with Python / numpy floored-modulo semantics.  fast_slam_2.py:84-85, :125 */
double fs2o_wrap_pi(double a)
{
    double v = a + FS2O_PI;
    double m = fmod(v, FS2O_TWO_PI);
    if (m != 0.0) {
        if (m < 0.0) m += FS2O_TWO_PI;
    } else {
        m = 0.0; /* copysign(0, 2*pi) */
    }
    return m - FS2O_PI;
}

/*
 * geometry_utils.py:14-23   sqrt(delta^T inv(cov) delta), delta = b - a.
 * inv() is restated as adj/det with one reciprocal.  *singular is set when det == 0 (numpy raises
 * LinAlgError on an exactly singular matrix).  A negative form gives NaN exactly like np.sqrt.
 * THE OPERATION ORDER BELOW IS THE GATE CONTRACT: the device's exact test repeats it verbatim.
 */
double fs2o_mahalanobis(double ax, double ay, double bx, double by,
                        double c00, double c01, double c10, double c11, int *singular)
{
    double det = c00 * c11 - c01 * c10;
    if (singular) *singular = (det == 0.0);
    double r = 1.0 / det;
    double i00 = c11 * r, i01 = -c01 * r, i10 = -c10 * r, i11 = c00 * r;
    double dx = bx - ax, dy = by - ay;
    double t0 = dx * i00 + dy * i10; /* delta.T @ inv */
    double t1 = dx * i01 + dy * i11;
    double s = t0 * dx + t1 * dy;     /* ... @ delta */
    return sqrt(s);
}

/*
 * landmark_utils.py:92-117   first landmark, in list order, with distance < gate; -1 if none;
 * -2 if a singular covariance was met before any match (the reference raises there).
 */
int fs2o_associate(double ox, double oy, const double *lm_p, int count, int lcap, double gate)
{
    (void)lcap;
    for (int i = 0; i < count; ++i) {
        int sing = 0;
        const double *l = lm_p + 6 * (size_t)i;
        double d = fs2o_mahalanobis(l[0], l[1], ox, oy, l[2], l[3], l[4], l[5], &sing);
        if (sing) return -2;
        if (d < gate) return i;
    }
    return -1;
}

/* fast_slam_2.py:69-87   one Gaussian draw per particle is passed in as noise[i] (already scaled). */
void fs2o_motion(int64_t P, double *x, double *y, double *yaw, double rotation, double translation,
                 const double *noise)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < P; ++i) {
        double nt, nr;
        if (rotation != 0.0) {
            nt = 0.0;
            nr = rotation + noise[i];
        } else {
            nt = translation + noise[i];
            nr = 0.0;
        }
        double a = fs2o_wrap_pi(yaw[i] + nr);
        yaw[i] = a;
        x[i] = x[i] + nt * cos(a);
        y[i] = y[i] + nt * sin(a);
    }
}

/*
 * scipy.stats.multivariate_normal.pdf(nu, mean=0, cov=Q) restated for 2x2 (fast_slam_2.py:156).
 * scipy symmetrises by reading the LOWER triangle (eigh(lower=True)), rejects the matrix unless
 * lambda_min > 1e6*eps*max|lambda| (the _PSD check with allow_singular=False), then evaluates
 * exp(-0.5*(2*log(2pi) + log pdet + nu^T Qs^-1 nu)).  Returns 0 and leaves *out untouched on reject.
 */
int fs2o_mvn_pdf2(double n0, double n1, double q00, double q10, double q11, double *out)
{
    if (!(isfinite(q00) && isfinite(q10) && isfinite(q11) && isfinite(n0) && isfinite(n1))) return 0;
    double hm = 0.5 * (q00 + q11);
    double hd = 0.5 * (q00 - q11);
    double rad = sqrt(hd * hd + q10 * q10);
    double l0 = hm - rad, l1 = hm + rad;
    /* better conditioned small eigenvalue when hm > 0: det / l1 */
    double det = q00 * q11 - q10 * q10;
    if (l1 > 0.0) l0 = det / l1;
    double amax = fmax(fabs(l0), fabs(l1));
    double eps = 1e6 * 2.220446049250313e-16 * amax;
    if (l0 < -eps) return 0;   /* ValueError: not PSD      */
    if (!(l0 > eps)) return 0; /* LinAlgError: singular    */
    double maha = (q11 * n0 * n0 - 2.0 * q10 * n0 * n1 + q00 * n1 * n1) / det;
    double logpdet = log(l0) + log(l1);
    *out = exp(-0.5 * (2.0 * FS2O_LOG_2PI + logpdet + maha));
    return 1;
}

/*
 * fast_slam_2.py:89-159   one (particle, measurement) update.  lm_p points at the particle's
 * [lcap][6] block.  Returns the association index (-1 = new landmark appended, -2 = skipped).
 */
int fs2o_update_one(double px, double py, double pyaw, double *w, int32_t *count, double *lm_p,
                    int lcap, double zd, double za, const double R[4], double gate, int32_t *status)
{
    /* :100-103 observation in the ROBOT frame (quirk Q1) */
    double ox = zd * cos(za), oy = zd * sin(za);
    int idx = fs2o_associate(ox, oy, lm_p, *count, lcap, gate);
    if (idx == -2) {
        *status |= FS2O_ST_SINGULAR_LM;
        return -2;
    }
    if (idx < 0) {
        /* :108-111 new landmark in the world frame, cov = 0.1*I (landmark.py:13), weight untouched */
        if (*count >= lcap) {
            *status |= FS2O_ST_MAP_FULL;
            return -1;
        }
        int j = *count;
        double *n = lm_p + 6 * (size_t)j;
        n[0] = px + zd * cos(pyaw + za);
        n[1] = py + zd * sin(pyaw + za);
        n[2] = 0.1; n[3] = 0.0; n[4] = 0.0; n[5] = 0.1;
        *count = j + 1;
        return -1;
    }
    /* :116-121 predicted measurement */
    double *L = lm_p + 6 * (size_t)idx;
    double s00 = L[2], s01 = L[3], s10 = L[4], s11 = L[5];
    double dx = L[0] - px, dy = L[1] - py;
    double q = dx * dx + dy * dy;
    double dist = sqrt(q);
    double ang = atan2(dy, dx) - pyaw;
    /* :124-125 innovation, angle wrapped */
    double n0 = zd - dist;
    double n1 = fs2o_wrap_pi(za - ang);
    /* :130-133 Jacobian w.r.t. the landmark */
    double h00 = dx / dist, h01 = dy / dist, h10 = -dy / q, h11 = dx / q;
    /* :137 Q = H S H^T + R */
    double a00 = h00 * s00 + h01 * s10, a01 = h00 * s01 + h01 * s11;
    double a10 = h10 * s00 + h11 * s10, a11 = h10 * s01 + h11 * s11;
    double q00 = a00 * h00 + a01 * h01 + R[0], q01 = a00 * h10 + a01 * h11 + R[1];
    double q10 = a10 * h00 + a11 * h01 + R[2], q11 = a10 * h10 + a11 * h11 + R[3];
    /* :142 K = S H^T inv(Q) */
    double detq = q00 * q11 - q01 * q10;
    if (detq == 0.0) {
        *status |= FS2O_ST_SINGULAR_Q;
        return -2;
    }
    double rq = 1.0 / detq;
    double v00 = q11 * rq, v01 = -q01 * rq, v10 = -q10 * rq, v11 = q00 * rq;
    double b00 = s00 * h00 + s01 * h01, b01 = s00 * h10 + s01 * h11;
    double b10 = s10 * h00 + s11 * h01, b11 = s10 * h10 + s11 * h11;
    double k00 = b00 * v00 + b01 * v10, k01 = b00 * v01 + b01 * v11;
    double k10 = b10 * v00 + b11 * v10, k11 = b10 * v01 + b11 * v11;
    /* :145-146 mean and (I - K H) S, stored un-symmetrised (Q5) */
    double mx = L[0] + (k00 * n0 + k01 * n1);
    double my = L[1] + (k10 * n0 + k11 * n1);
    double g00 = 1.0 - (k00 * h00 + k01 * h10), g01 = 0.0 - (k00 * h01 + k01 * h11);
    double g10 = 0.0 - (k10 * h00 + k11 * h10), g11 = 1.0 - (k10 * h01 + k11 * h11);
    /* :149-153 replace the landmark (before the likelihood, so a pdf failure keeps the new landmark) */
    L[0] = mx;
    L[1] = my;
    L[2] = g00 * s00 + g01 * s10;
    L[3] = g00 * s01 + g01 * s11;
    L[4] = g10 * s00 + g11 * s10;
    L[5] = g10 * s01 + g11 * s11;
    /* :156-159 weight *= N(nu; 0, Q) */
    double like;
    if (!fs2o_mvn_pdf2(n0, n1, q00, q10, q11, &like)) {
        *status |= FS2O_ST_PDF_FAILED;
        return idx;
    }
    *w = *w * like;
    return idx;
}

/*
 * fast_slam_2.py:48-53   all measurements, sequentially per particle (Q7).  The reference loops
 * measurement-major; particles are independent so particle-major is the same computation.
 * obs: [M][2] = (distance, yaw).  assoc_out: [M][P] or NULL.  wkind: see fs2o_normalize, or NULL.
 */
void fs2o_update(int64_t P, const double *x, const double *y, const double *yaw, double *w,
                 int32_t *count, double *lm, int lcap, const double *obs, int M, const double R[4],
                 double gate, int32_t *assoc_out, int32_t *status, uint8_t *wkind)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < P; ++i) {
        double *lm_p = lm + (size_t)i * 6 * (size_t)lcap;
        int32_t st = 0;
        for (int k = 0; k < M; ++k) {
            int32_t stk = 0;
            int a = fs2o_update_one(x[i], y[i], yaw[i], &w[i], &count[i], lm_p, lcap, obs[2 * k],
                                    obs[2 * k + 1], R, gate, &stk);
            if (assoc_out) assoc_out[(size_t)k * P + i] = a;
            /* weight *= np.float64 turns a Python float into an np.float64 (see fs2o_normalize) */
            if (wkind && a >= 0 && !(stk & FS2O_ST_PDF_FAILED)) wkind[i] = 0;
            st |= stk;
        }
        if (status) status[i] |= st;
    }
}

/*
 * fast_slam_2.py:161-175 (Q8, Q18).  `sum(p.weight for p in particles)` is CPython's builtin sum:
 * while every item seen so far is an exact Python float it runs the Neumaier-compensated loop
 * (CPython >= 3.12, bltinmodule.c); at the first np.float64 item it folds the compensation in and
 * continues with plain left-to-right additions.  wkind[i] = 1 marks a weight that is an exact
 * Python float in the reference (initial 1/N, or reset by this function), 0 an np.float64
 * (after `weight *= likelihood` or a division by an np.float64 total).  wkind == NULL: all np.float64.
 * Returns the total.
 */
double fs2o_weight_total(int64_t P, const double *w, const uint8_t *wkind, int *total_is_pyfloat)
{
    double f = 0.0, c = 0.0;
    int64_t i = 0;
    int fast = 1;
    /* the int 0 start value plus the first item: exact in both paths */
    if (wkind) {
        for (; i < P && wkind[i]; ++i) {
            double xv = w[i];
            double t = f + xv;
            if (fabs(f) >= fabs(xv)) c += (f - t) + xv; else c += (xv - t) + f;
            f = t;
        }
        if (c != 0.0 && isfinite(c)) f += c;
        fast = (i == P);
    } else {
        fast = (P == 0);
    }
    for (; i < P; ++i) f = f + w[i];
    if (total_is_pyfloat) *total_is_pyfloat = fast;
    return f;
}

double fs2o_normalize(int64_t P, double *w, uint8_t *wkind)
{
    int tot_py = 0;
    double total = fs2o_weight_total(P, w, wkind, &tot_py);
    if (total < 1e-5) {
        double u = 1.0 / (double)P;
        for (int64_t i = 0; i < P; ++i) { w[i] = u; if (wkind) wkind[i] = 1; }
    } else {
        for (int64_t i = 0; i < P; ++i) {
            if (!(w[i] < 1e-5)) {
                w[i] = w[i] / total;
                if (wkind) wkind[i] = (uint8_t)(wkind[i] && tot_py);
            }
        }
    }
    return total;
}

/* numpy's pairwise summation (umath loops_utils.h.src) of v[i] = w[i]^2; np.sum(weights ** 2), :219 */
static double pairwise_sq(const double *a, int64_t n)
{
    if (n < 8) {
        double res = 0.0;
        for (int64_t i = 0; i < n; ++i) res += a[i] * a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j] * a[j];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j] * a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i] * a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sq(a, n2) + pairwise_sq(a + n2, n - n2);
    }
}

/* fast_slam_2.py:212-223 (Q9) */
double fs2o_neff(int64_t P, const double *w)
{
    double s = pairwise_sq(w, P);
    if (s < 1.0 / (double)P) return (double)P;
    return 1.0 / s;
}

/*
 * fast_slam_2.py:183-196 (Q10) low-variance resampling indices.  Sequential fp64 running sum.
 * Returns 0, or 1 if the reference would loop forever (u beyond the total with w[P-1] == 0):
 * the remaining slots then get P-1.
 */
int fs2o_resample_indices(int64_t P, const double *w, double u0, int32_t *idx)
{
    double inv = 1.0 / (double)P;
    double c = w[0];
    int64_t k = 0;
    int stuck = 0;
    for (int64_t m = 0; m < P; ++m) {
        double u = u0 + (double)m * inv;
        while (u > c) {
            if (k == P - 1 && !(w[k] > 0.0)) { stuck = 1; break; }
            k = (k + 1 < P - 1) ? k + 1 : P - 1;
            c += w[k];
        }
        idx[m] = (int32_t)k;
    }
    return stuck;
}

/* The sequential running sum itself (c_k), for stage-wise checks of the device scan. */
void fs2o_cumsum_seq(int64_t P, const double *w, double *c)
{
    double s = 0.0;
    for (int64_t i = 0; i < P; ++i) { s = (i == 0) ? w[0] : s + w[i]; c[i] = s; }
}

/* deepcopy of the survivors, weight included (D6), fast_slam_2.py:196-199 */
void fs2o_gather(int64_t P, const int32_t *idx, int lcap,
                 const double *x, const double *y, const double *yaw, const double *w,
                 const int32_t *count, const double *lm, const uint8_t *wkind,
                 double *x2, double *y2, double *yaw2, double *w2, int32_t *count2, double *lm2,
                 uint8_t *wkind2)
{
    size_t stride = (size_t)6 * (size_t)lcap;
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < P; ++m) {
        int64_t s = idx[m];
        x2[m] = x[s]; y2[m] = y[s]; yaw2[m] = yaw[s]; w2[m] = w[s]; count2[m] = count[s];
        if (wkind && wkind2) wkind2[m] = wkind[s];
        memcpy(lm2 + (size_t)m * stride, lm + (size_t)s * stride, stride * sizeof(double));
    }
}

/* fast_slam_2.py:201-210 (Q11) first arg-max */
int64_t fs2o_argmax(int64_t P, const double *w)
{
    int64_t b = 0;
    for (int64_t i = 1; i < P; ++i)
        if (w[i] > w[b]) b = i;
    return b;
}

/*
 * fast_slam_2.py:33-67   one whole filter step on a state owned by the caller.
 * scratch buffers (x2.. lm2, idx) are only touched when the step resamples; after a resample the
 * results are copied back so the caller's pointers stay valid.
 * out[0..2] = estimate, out[3] = neff, out[4] = resampled (0/1), out[5] = total weight before normalise.
 */
int fs2o_step(int64_t P, int lcap, double *x, double *y, double *yaw, double *w, int32_t *count,
              double *lm, uint8_t *wkind, double rotation, double translation, const double *noise,
              const double *obs, int M, double u0, const double R[4], double gate,
              int32_t *assoc_out, int32_t *idx_out, int32_t *status, double *scratch_pose /*[4P]*/,
              int32_t *scratch_count, double *scratch_lm, double *out)
{
    fs2o_motion(P, x, y, yaw, rotation, translation, noise);
    fs2o_update(P, x, y, yaw, w, count, lm, lcap, obs, M, R, gate, assoc_out, status, wkind);
    double total = fs2o_normalize(P, w, wkind);
    double neff = fs2o_neff(P, w);
    int resampled = 0;
    if (neff < (double)P / 2.0) {
        resampled = 1;
        int32_t *idx = idx_out ? idx_out : (int32_t *)malloc(sizeof(int32_t) * (size_t)P);
        fs2o_resample_indices(P, w, u0, idx);
        uint8_t *wk2 = wkind ? (uint8_t *)malloc((size_t)P) : NULL;
        fs2o_gather(P, idx, lcap, x, y, yaw, w, count, lm, wkind, scratch_pose, scratch_pose + P,
                    scratch_pose + 2 * P, scratch_pose + 3 * P, scratch_count, scratch_lm, wk2);
        memcpy(x, scratch_pose, sizeof(double) * (size_t)P);
        memcpy(y, scratch_pose + P, sizeof(double) * (size_t)P);
        memcpy(yaw, scratch_pose + 2 * P, sizeof(double) * (size_t)P);
        memcpy(w, scratch_pose + 3 * P, sizeof(double) * (size_t)P);
        memcpy(count, scratch_count, sizeof(int32_t) * (size_t)P);
        memcpy(lm, scratch_lm, sizeof(double) * (size_t)P * 6 * (size_t)lcap);
        if (wkind) { memcpy(wkind, wk2, (size_t)P); free(wk2); }
        if (!idx_out) free(idx);
    } else if (idx_out) {
        for (int64_t i = 0; i < P; ++i) idx_out[i] = (int32_t)i;
    }
    int64_t b = fs2o_argmax(P, w);
    out[0] = x[b]; out[1] = y[b]; out[2] = yaw[b];
    out[3] = neff; out[4] = (double)resampled; out[5] = total;
    return 0;
}

/*
 * Update-only timing leg for bench.py's cpu_baseline: motion + update + normalise + neff + argmax on a
 * bounded sample of particles; no resample (the sample's weights are not the full set's).
 */
void fs2o_step_noresample(int64_t P, int lcap, double *x, double *y, double *yaw, double *w,
                          int32_t *count, double *lm, double rotation, double translation,
                          const double *noise, const double *obs, int M, const double R[4],
                          double gate, double *out)
{
    fs2o_motion(P, x, y, yaw, rotation, translation, noise);
    fs2o_update(P, x, y, yaw, w, count, lm, lcap, obs, M, R, gate, NULL, NULL, NULL);
    double total = fs2o_normalize(P, w, NULL);
    double neff = fs2o_neff(P, w);
    int64_t b = fs2o_argmax(P, w);
    out[0] = x[b]; out[1] = y[b]; out[2] = yaw[b];
    out[3] = neff; out[4] = 0.0; out[5] = total;
}
