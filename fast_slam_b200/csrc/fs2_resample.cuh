// fs2_resample.cuh -- low-variance resampling (row A9 of SURVEY.md 8a; reference fast_slam_2.py:177-199).
//
// The reference walks a SEQUENTIAL fp64 running sum c_k = fl(c_{k-1} + w_k) and picks, for every slot m,
// the first k with c_k >= u0 + m/N (quirk Q10).  Resampling indices must be bit-exact, and a parallel
// prefix sum rounds differently (expected ~0.1 flipped indices per resample at N = 2^20), so the scan
// below reproduces the sequential rounding exactly:
//
//   While the running sum stays inside one binade [2^e, 2^(e+1)) its ulp is fixed, u = 2^(e-52), and
//   fl(c + w) = c + round(w/u)*u, where round() is to nearest and a tie goes to the neighbour that makes
//   c/u even.  So with C = c/u an integer, adding w is  C -> C + a + t(C)  with a = floor(w/u) and
//   t in {0, 1} depending only on frac(w/u) and, for an exact tie, on the parity of C + a.  Such a step
//   is a function of the PARITY of C only:  f = (inc if C even, inc if C odd), and these functions
//   compose associatively:  (f then g)_p = f_p + g_{(p + f_p) mod 2}.  That is a parallel scan.
//
//   1. fs2_scan_blocksum     plain fp64 block sums (256 weights per block) and a validity flag
//   2. fs2_scan_blockprefix  their exclusive prefix (approximate) -> the binade e_b each block starts in
//   3. fs2_scan_groupfunc    the composed parity function (A0, A1) of each block in units of 2^(e_b-52), and of
//                            each GROUP of 32 blocks; then, in the last CTA to finish,
//   4.   fs2_scan_walk       one warp walks the groups with the EXACT running sum, 32 group functions per step:
//                            groups whose assumed binade holds at entry and exit are applied at once; a group in
//                            which the sum crosses a power of two (~20 per scan) is walked block by block and
//                            its straddling block element by element
//   5. fs2_scan_emit         exact c_k for every element (block starts from the group's start and the block
//                            functions, parity scan inside the block, serial for the straddling blocks)
//   6. fs2_resample_search   k(m) by binary search over c (monotone), clamped to N-1
//
// Negative / non-finite weights (never produced by a healthy filter) take fs2_resample_serial, one
// thread repeating the reference's loop literally.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FS2_SCAN_B 256        // weights per scan block (a block that straddles a power of two is walked serially)
#define FS2_SCAN_T 256        // threads per scan block
#define FS2_SCAN_EPT (FS2_SCAN_B / FS2_SCAN_T)   // consecutive weights per thread
#define FS2_MODE_PARITY 0
#define FS2_MODE_SERIAL 1
#define FS2_MODE_ZERO 2       // all weights of the block are +0: identity

// control words of the fused step (fs2_step.cuh): written on the device, read by every kernel of the resample chain
#define FS2_CTL_RES 0       // this step resamples
#define FS2_CTL_ANOMALY 1   // weights the exact scan does not accept
#define FS2_CTL_SERIAL 2    // ancestors already written by the literal serial loop
#define FS2_CTL_STUCK 3     // a slot beyond the running total (quirk Q10: the reference would not terminate)
#define FS2_CTL_LEN 8

struct Fs2Par {   // increment of C when the incoming C is even (e) / odd (o)
    unsigned long long e, o;
};

__device__ __forceinline__ Fs2Par fs2_par_compose(Fs2Par f, Fs2Par g)
{
    Fs2Par h;
    h.e = f.e + ((f.e & 1ull) ? g.o : g.e);          // parity after f from even = f.e mod 2
    h.o = f.o + (((1ull + f.o) & 1ull) ? g.o : g.e);
    return h;
}

// parity function of adding w while the running sum sits in the binade of exponent e (unit 2^(e-52)).
// *big is set when w >= 2^(e+1), i.e. the sum cannot stay in the binade.
__device__ __forceinline__ Fs2Par fs2_par_of(double w, int e, bool *big)
{
    Fs2Par f;
    f.e = f.o = 0ull;
    unsigned long long bits = (unsigned long long)__double_as_longlong(w);
    int ew = (int)((bits >> 52) & 0x7ffull);
    unsigned long long m = bits & 0x000fffffffffffffull;
    if (ew) m |= 0x0010000000000000ull; else ew = 1;   // subnormal: no hidden bit, exponent of 2^-1074
    if (m == 0ull) return f;                            // +0
    int x = ew - 1075;                                  // w = m * 2^x
    int s = (e - 52) - x;                               // bits of m below the unit
    if (s <= 0) {                                       // w is a multiple of the unit
        if (s < -1) { *big = true; return f; }
        unsigned long long a = m << (-s);
        if (a >= 0x0020000000000000ull) { *big = true; return f; }
        f.e = f.o = a;
        return f;
    }
    if (s >= 64) return f;                              // far below half a unit: rounds away
    unsigned long long a = m >> s;
    unsigned long long rem = m & ((1ull << s) - 1ull);
    unsigned long long half = 1ull << (s - 1);
    if (rem < half) { f.e = f.o = a; }
    else if (rem > half) { f.e = f.o = a + 1ull; }
    else {                                              // exact tie: to even
        f.e = a + (a & 1ull);
        f.o = a + ((a + 1ull) & 1ull);
    }
    return f;
}

__device__ __forceinline__ int fs2_exponent(double c)  // unbiased exponent of a positive normal double, else INT_MIN
{
    unsigned long long bits = (unsigned long long)__double_as_longlong(c);
    int ec = (int)((bits >> 52) & 0x7ffull);
    if ((bits >> 63) || ec == 0 || ec == 0x7ff) return INT_MIN;
    return ec - 1023;
}

__global__ void __launch_bounds__(FS2_SCAN_T)
fs2_scan_blocksum(const double *__restrict__ w, int64_t n, double *bsum, int *anomaly)
{
    __shared__ double ws[FS2_SCAN_T / 32];
    int64_t base = (int64_t)blockIdx.x * FS2_SCAN_B;
    double s = 0.0;
    bool bad = false;
#pragma unroll
    for (int j = 0; j < FS2_SCAN_B / FS2_SCAN_T; ++j) {
        int64_t i = base + threadIdx.x + j * FS2_SCAN_T;
        if (i < n) {
            double v = w[i];
            bad |= !(v >= 0.0) || !(v < 1.797e308);
            s += v;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(anomaly, 1);
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < FS2_SCAN_T / 32; ++k) t += ws[k];
        bsum[blockIdx.x] = t;
    }
}

// single block: exclusive prefix of bsum[nb] (plain fp64, approximate by design)
__global__ void __launch_bounds__(1024) fs2_scan_blockprefix(const double *bsum, int nb, double *bpre)
{
    __shared__ double ws[32];
    __shared__ double carry_s;
    if (threadIdx.x == 0) carry_s = 0.0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        int i = base + threadIdx.x;
        double v = (i < nb) ? bsum[i] : 0.0;
        double incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            double t = ws[threadIdx.x];
            double ti = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                double u = __shfl_up_sync(0xffffffffu, ti, o);
                if (threadIdx.x >= o) ti += u;
            }
            ws[threadIdx.x] = ti - t;  // exclusive over warps
        }
        __syncthreads();
        double carry = carry_s;
        double excl = carry + ws[threadIdx.x >> 5] + (incl - v);
        if (i < nb) bpre[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + ws[31] + incl;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Block functions, GROUP functions and the exact walk.
//
// A group is FS2_GRP = 32 consecutive scan blocks (8192 weights), handled by one CTA of 8 warps: every warp computes
// the parity functions of four blocks (a lane composes 8 consecutive weights, then an ordered warp reduction), warp 0
// composes the 32 block functions into the group's function if they were all built for the same binade.  The exact
// running sum is then walked by ONE warp, 32 GROUPS per step (fs2_walk32): a prefix of groups whose assumed binade
// holds at entry and exit is accepted at once; the first group that does not -- the sum crosses a power of two in
// it, or it holds blocks that need the literal sum -- is walked block by block, and the one block in it that
// straddles the power of two weight by weight.  2^20 weights are 128 groups: four steps plus ~20 binade crossings,
// instead of 128 steps over blocks.  The walk runs in the last CTA to finish (no extra launch).
// ------------------------------------------------------------------------------------------------
#define FS2_GRP 32
#define FS2_GRP_T 256
#define FS2_KIND_ZERO 0      // every block of the group is all-zero: identity
#define FS2_KIND_PARITY 1    // one function for the whole group, valid in binade gE
#define FS2_KIND_MIXED 2     // must be walked block by block

struct Fs2Scan {             // scratch of the exact scan (sized for the global particle count)
    unsigned long long *A0, *A1;    // per block: increment if the incoming sum is even / odd
    int *eb, *mode;                 // per block: assumed binade, FS2_MODE_*
    double *cstart;                 // per block: exact sum before it (written by the walk for groups it descended into)
    unsigned long long *gA0, *gA1;  // per group
    int *gE, *gK, *gV;              // per group: binade, FS2_KIND_*, 1 = accepted as a whole (cg valid)
    double *cg;                     // per group: exact sum before it
    double *total;
};

__device__ __forceinline__ double fs2_sum_from(unsigned long long bits, unsigned long long C)
{
    return __longlong_as_double((long long)((bits & 0xfff0000000000000ull) | (C & 0x000fffffffffffffull)));
}

// parity function of scan block b, by one warp: lane l composes weights 8l .. 8l+7 of the block, lane 0 returns it
__device__ __forceinline__ Fs2Par fs2_block_func(const double *__restrict__ w, int64_t n, int64_t b, int e, int lane, bool *big_any)
{
    const int64_t base = b * FS2_SCAN_B + 8 * lane;
    double v[8];
    if (base + 8 <= n) {
        const double2 *p = reinterpret_cast<const double2 *>(w + base);
        const double2 a0 = p[0], a1 = p[1], a2 = p[2], a3 = p[3];
        v[0] = a0.x; v[1] = a0.y; v[2] = a1.x; v[3] = a1.y; v[4] = a2.x; v[5] = a2.y; v[6] = a3.x; v[7] = a3.y;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (base + j < n) ? w[base + j] : 0.0;     // +0 composes as the identity
    }
    Fs2Par f;
    f.e = f.o = 0ull;
    bool big = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) f = fs2_par_compose(f, fs2_par_of(v[j], e, &big));
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {       // ordered reduction (lane order = element order)
        Fs2Par g;
        g.e = __shfl_down_sync(0xffffffffu, f.e, o);
        g.o = __shfl_down_sync(0xffffffffu, f.o, o);
        if ((lane % (2 * o)) == 0) f = fs2_par_compose(f, g);
    }
    *big_any = __any_sync(0xffffffffu, big);
    return f;
}

// One step of the walk over <= 32 consecutive functions (lane order): accepts the longest prefix that can be applied
// to the exact running sum c at once, returns its length, advances c, and hands every accepted lane the sum BEFORE
// its function in *start.  kind: FS2_KIND_ZERO (identity), FS2_KIND_PARITY (f valid in binade e), anything else never
// accepted.
__device__ __forceinline__ int fs2_walk32(double &c, Fs2Par f, int e, int kind, bool valid, int lane, double *start)
{
    const unsigned full = 0xffffffffu;
    if (!(valid && kind == FS2_KIND_PARITY)) f.e = f.o = 0ull;
    Fs2Par inc = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Fs2Par g;
        g.e = __shfl_up_sync(full, inc.e, o);
        g.o = __shfl_up_sync(full, inc.o, o);
        if (lane >= o) inc = fs2_par_compose(g, inc);
    }
    Fs2Par ex;
    ex.e = __shfl_up_sync(full, inc.e, 1);
    ex.o = __shfl_up_sync(full, inc.o, 1);
    if (lane == 0) ex.e = ex.o = 0ull;
    const int ec = fs2_exponent(c);
    const unsigned long long bits = (unsigned long long)__double_as_longlong(c);
    const unsigned long long C = (bits & 0x000fffffffffffffull) | 0x0010000000000000ull;
    const bool odd = (C & 1ull) != 0ull;
    const unsigned long long Cin = C + (odd ? ex.o : ex.e), Cout = C + (odd ? inc.o : inc.e);
    // fine if it changes nothing, or if it was composed for the binade the sum is in and the sum stays there
    const bool ok = valid && (kind == FS2_KIND_ZERO ||
                              (kind == FS2_KIND_PARITY && ec != INT_MIN && e == ec && Cout < 0x0020000000000000ull));
    const unsigned okm = __ballot_sync(full, ok);
    const int nacc = (okm == full) ? 32 : (__ffs(~okm) - 1);
    *start = (ec == INT_MIN) ? c : fs2_sum_from(bits, Cin);           // all-zero functions before the sum starts
    if (nacc > 0) {
        const unsigned long long Cl = __shfl_sync(full, Cout, nacc - 1);
        if (ec != INT_MIN) c = fs2_sum_from(bits, Cl);
    }
    return nacc;
}

// the walk itself: warp 0 of the calling CTA
__device__ __forceinline__ void fs2_scan_walk(const double *w, int64_t n, int nb, const Fs2Scan &sc, int lane)
{
    const unsigned full = 0xffffffffu;
    const int ng = (nb + FS2_GRP - 1) / FS2_GRP;
    double c = 0.0;                                   // the exact running sum, identical in every lane
    int g0 = 0;
    while (g0 < ng) {
        const int g = g0 + lane;
        const bool valid = g < ng;
        Fs2Par f;
        f.e = f.o = 0ull;
        int e = INT_MIN, kind = FS2_KIND_MIXED;
        if (valid) {
            kind = ((volatile int *)sc.gK)[g]; e = ((volatile int *)sc.gE)[g];
            f.e = ((volatile unsigned long long *)sc.gA0)[g]; f.o = ((volatile unsigned long long *)sc.gA1)[g];
        }
        double start;
        const int nacc = fs2_walk32(c, f, e, kind, valid, lane, &start);
        if (lane < nacc) { sc.cg[g] = start; sc.gV[g] = 1; }
        g0 += nacc;
        if (nacc < 32 && g0 < ng) {
            // group g0 cannot be applied as a whole: its blocks, 32 at a time, and weight by weight where needed
            if (lane == 0) sc.gV[g0] = 0;
            const int bfirst = g0 * FS2_GRP;
            const int nbg = min(FS2_GRP, nb - bfirst);
            int b0 = 0;
            while (b0 < nbg) {
                const int i = b0 + lane;
                const bool bvalid = i < nbg;
                Fs2Par bf;
                bf.e = bf.o = 0ull;
                int be = INT_MIN, bk = FS2_KIND_MIXED;
                if (bvalid) {
                    const int md = ((volatile int *)sc.mode)[bfirst + i];
                    be = ((volatile int *)sc.eb)[bfirst + i];
                    bk = (md == FS2_MODE_ZERO) ? FS2_KIND_ZERO : (md == FS2_MODE_PARITY ? FS2_KIND_PARITY : FS2_KIND_MIXED);
                    if (md == FS2_MODE_PARITY) {
                        bf.e = ((volatile unsigned long long *)sc.A0)[bfirst + i];
                        bf.o = ((volatile unsigned long long *)sc.A1)[bfirst + i];
                    }
                }
                double bstart;
                const int na = fs2_walk32(c, bf, be, bk, bvalid, lane, &bstart);
                if (lane < na) sc.cstart[bfirst + i] = bstart;
                b0 += na;
                if (na < 32 && b0 < nbg) {
                    // block bfirst + b0, exactly: the reference's "particle_weight += w[k]"
                    const int64_t lo = (int64_t)(bfirst + b0) * FS2_SCAN_B;
                    const int64_t hi = lo + FS2_SCAN_B < n ? lo + FS2_SCAN_B : n;
                    if (lane == 0) { sc.cstart[bfirst + b0] = c; sc.mode[bfirst + b0] = FS2_MODE_SERIAL; }
                    double r[FS2_SCAN_B / 32];
#pragma unroll
                    for (int k2 = 0; k2 < FS2_SCAN_B / 32; ++k2) {
                        const int64_t k = lo + 32 * k2 + lane;
                        r[k2] = (k < hi) ? w[k] : 0.0;
                    }
#pragma unroll
                    for (int k2 = 0; k2 < FS2_SCAN_B / 32; ++k2) {
                        for (int j = 0; j < 32; ++j) {
                            const double v = __shfl_sync(full, r[k2], j);
                            const int64_t k = lo + 32 * k2 + j;
                            if (k < hi) c = (k == 0) ? v : __dadd_rn(c, v);
                        }
                    }
                    b0 += 1;
                }
            }
            g0 += 1;
        }
    }
    if (lane == 0) *sc.total = c;
}

// What the fused step (fs2_step.cuh) wants done on the way; all null / zero for the stage-wise API.
struct Fs2ScanFused {
    int *ctl;                       // FS2_CTL_*: return at once unless ctl[RES]; literal serial loop if ctl[ANOMALY]
    double u0;                      // resampling start (serial loop only)
    int32_t *ancestor;              // (serial loop only)
    double *cum;                    // (serial loop only)
    int32_t *used;                  // slot-use flags to clear, S of them
    int64_t S;
    unsigned long long *is_state;   // state words of the one-pass slot scan to clear
    int is_tiles;
    unsigned int *is_ticket;
};

__device__ __forceinline__ bool fs2_last_cta(unsigned int *counter)
{
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(counter, 1u);
        last = (t == gridDim.x - 1);
        if (last) *counter = 0;  // re-arm for the next launch
    }
    __syncthreads();
    return last;
}

// grid = number of groups, FS2_GRP_T threads
__global__ void __launch_bounds__(FS2_GRP_T)
fs2_scan_groupfunc(const double *__restrict__ w, int64_t n, int nb, const double *bsum, const double *bpre, const Fs2Scan sc,
                   unsigned int *counter, const Fs2ScanFused fu)
{
    if (fu.ctl && !((volatile int *)fu.ctl)[FS2_CTL_RES]) return;
    __shared__ Fs2Par sF[FS2_GRP];
    __shared__ int sE[FS2_GRP], sM[FS2_GRP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x;
    bool anomaly = false;
    if (fu.ctl) {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < fu.S; i += (int64_t)gridDim.x * blockDim.x) fu.used[i] = 0;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < fu.is_tiles; i += gridDim.x * blockDim.x) fu.is_state[i] = 0ull;
        if (blockIdx.x == 0 && threadIdx.x == 0) *fu.is_ticket = 0u;
        anomaly = ((volatile int *)fu.ctl)[FS2_CTL_ANOMALY] != 0;
    }
    if (!anomaly) {
        const int bfirst = g * FS2_GRP;
        const int nbg = min(FS2_GRP, nb - bfirst);
        for (int bi = warp; bi < nbg; bi += FS2_GRP_T / 32) {
            const int b = bfirst + bi;
            const int e = fs2_exponent(bpre[b]);
            Fs2Par f;
            f.e = f.o = 0ull;
            int md;
            if (bsum[b] == 0.0) md = FS2_MODE_ZERO;               // non-negative weights summing to +0: nothing changes
            else if (e == INT_MIN) md = FS2_MODE_SERIAL;          // running sum still zero / subnormal here: walk it
            else {
                bool big;
                f = fs2_block_func(w, n, b, e, lane, &big);
                md = big ? FS2_MODE_SERIAL : FS2_MODE_PARITY;
            }
            if (lane == 0) {
                sc.A0[b] = f.e; sc.A1[b] = f.o; sc.eb[b] = e; sc.mode[b] = md;
                sF[bi] = f; sE[bi] = e; sM[bi] = md;
            }
        }
        __syncthreads();
        if (warp == 0) {
            const bool valid = lane < nbg;
            const int md = valid ? sM[lane] : FS2_MODE_ZERO;
            const int e = valid ? sE[lane] : INT_MIN;
            Fs2Par f;
            f.e = f.o = 0ull;
            if (valid && md == FS2_MODE_PARITY) f = sF[lane];
            const unsigned par = __ballot_sync(0xffffffffu, valid && md == FS2_MODE_PARITY);
            const unsigned ser = __ballot_sync(0xffffffffu, valid && md == FS2_MODE_SERIAL);
            int kind = FS2_KIND_ZERO, eref = INT_MIN;
            if (ser) kind = FS2_KIND_MIXED;
            else if (par) {
                eref = __shfl_sync(0xffffffffu, e, __ffs(par) - 1);
                const bool same = __all_sync(0xffffffffu, !(valid && md == FS2_MODE_PARITY) || e == eref);
                kind = same ? FS2_KIND_PARITY : FS2_KIND_MIXED;
            }
            Fs2Par inc = f;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                Fs2Par t;
                t.e = __shfl_up_sync(0xffffffffu, inc.e, o);
                t.o = __shfl_up_sync(0xffffffffu, inc.o, o);
                if (lane >= o) inc = fs2_par_compose(t, inc);
            }
            if (lane == 31) { sc.gA0[g] = inc.e; sc.gA1[g] = inc.o; sc.gE[g] = eref; sc.gK[g] = kind; }
        }
    }
    if (!fs2_last_cta(counter)) return;
    if (anomaly) {
        // negative / non-finite weights: the reference's loop, literally (fs2_resample_serial)
        if (threadIdx.x == 0) {
            const double inv = 1.0 / (double)n;
            double c = w[0];
            int64_t k = 0;
            for (int64_t i = 0; i < n; ++i) fu.cum[i] = (i == 0) ? w[0] : __dadd_rn(fu.cum[i - 1], w[i]);
            for (int64_t m = 0; m < n; ++m) {
                const double u = __dadd_rn(fu.u0, __dmul_rn((double)m, inv));
                while (u > c) {
                    if (k == n - 1 && !(w[k] > 0.0)) break;
                    k = (k + 1 < n - 1) ? k + 1 : n - 1;
                    c = __dadd_rn(c, w[k]);
                }
                fu.ancestor[m] = (int32_t)k;
            }
            fu.ctl[FS2_CTL_SERIAL] = 1;
        }
        return;
    }
    if (warp == 0) fs2_scan_walk(w, n, nb, sc, lane);
}

// exact c_k for every element; grid = number of groups, FS2_GRP_T threads
__global__ void __launch_bounds__(FS2_GRP_T)
fs2_scan_emit(const double *__restrict__ w, int64_t n, int nb, const Fs2Scan sc, double *cum, const int *ctl)
{
    // fused step: nothing to do unless this step resamples through the exact scan
    if (ctl && (!((volatile const int *)ctl)[FS2_CTL_RES] || ((volatile const int *)ctl)[FS2_CTL_SERIAL])) return;
    __shared__ double sC[FS2_GRP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x;
    const int bfirst = g * FS2_GRP;
    const int nbg = min(FS2_GRP, nb - bfirst);
    if (warp == 0) {
        const bool valid = lane < nbg;
        double start = 0.0;
        if (sc.gV[g]) {                       // accepted as a whole: block starts from the group's start and the block functions
            Fs2Par f;
            f.e = f.o = 0ull;
            if (valid && sc.mode[bfirst + lane] == FS2_MODE_PARITY) { f.e = sc.A0[bfirst + lane]; f.o = sc.A1[bfirst + lane]; }
            Fs2Par inc = f;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                Fs2Par t;
                t.e = __shfl_up_sync(0xffffffffu, inc.e, o);
                t.o = __shfl_up_sync(0xffffffffu, inc.o, o);
                if (lane >= o) inc = fs2_par_compose(t, inc);
            }
            Fs2Par ex;
            ex.e = __shfl_up_sync(0xffffffffu, inc.e, 1);
            ex.o = __shfl_up_sync(0xffffffffu, inc.o, 1);
            if (lane == 0) ex.e = ex.o = 0ull;
            const double cgv = sc.cg[g];
            const unsigned long long bits = (unsigned long long)__double_as_longlong(cgv);
            const unsigned long long C = (bits & 0x000fffffffffffffull) | 0x0010000000000000ull;
            start = (fs2_exponent(cgv) == INT_MIN) ? cgv : fs2_sum_from(bits, C + ((C & 1ull) ? ex.o : ex.e));
        } else if (valid) {
            start = sc.cstart[bfirst + lane];
        }
        sC[lane] = start;
    }
    __syncthreads();
    for (int bi = warp; bi < nbg; bi += FS2_GRP_T / 32) {
        const int b = bfirst + bi;
        const int md = sc.mode[b];
        const double c0 = sC[bi];
        const int64_t lo = (int64_t)b * FS2_SCAN_B;
        const int64_t hi = lo + FS2_SCAN_B < n ? lo + FS2_SCAN_B : n;
        const int64_t base = lo + 8 * lane;
        if (md == FS2_MODE_ZERO) {
#pragma unroll
            for (int j = 0; j < 8; ++j) if (base + j < hi) cum[base + j] = (base + j == 0) ? w[0] : c0;
        } else if (md == FS2_MODE_SERIAL) {
            if (lane == 0) {
                double c = c0;
                for (int64_t i = lo; i < hi; ++i) { c = (i == 0) ? w[0] : __dadd_rn(c, w[i]); cum[i] = c; }
            }
        } else {
            const int e = sc.eb[b];
            Fs2Par loc[8];
            Fs2Par f;
            f.e = f.o = 0ull;
            bool big = false;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (base + j < hi) f = fs2_par_compose(f, fs2_par_of(w[base + j], e, &big));
                loc[j] = f;                       // inclusive within the lane
            }
            Fs2Par inc = f;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                Fs2Par t;
                t.e = __shfl_up_sync(0xffffffffu, inc.e, o);
                t.o = __shfl_up_sync(0xffffffffu, inc.o, o);
                if (lane >= o) inc = fs2_par_compose(t, inc);
            }
            Fs2Par pre;
            pre.e = __shfl_up_sync(0xffffffffu, inc.e, 1);
            pre.o = __shfl_up_sync(0xffffffffu, inc.o, 1);
            if (lane == 0) pre.e = pre.o = 0ull;
            const unsigned long long bits0 = (unsigned long long)__double_as_longlong(c0);
            const unsigned long long C0 = (bits0 & 0x000fffffffffffffull) | 0x0010000000000000ull;
            const bool odd = (C0 & 1ull) != 0ull;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (base + j < hi) {
                    const Fs2Par t = fs2_par_compose(pre, loc[j]);
                    cum[base + j] = fs2_sum_from(bits0, C0 + (odd ? t.o : t.e));
                }
            }
        }
    }
}

// k(m) = min{k : c_k >= u_m}, clamped to n-1 (fast_slam_2.py:187-193)
__global__ void __launch_bounds__(256)
fs2_resample_search(const double *__restrict__ cum, int64_t n, double u0, int64_t m_begin, int64_t m_count,
                    int32_t *ancestor, int *stuck)
{
    const double inv = 1.0 / (double)n;                           // 1 / NUM_PARTICLES
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < m_count; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t m = m_begin + t;
        double u = __dadd_rn(u0, __dmul_rn((double)m, inv));      // u0 + m * (1/N)
        int64_t lo = 0, hi = n;                                    // first index with !(u > c)
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (u > cum[mid]) lo = mid + 1; else hi = mid;
        }
        if (lo >= n) { lo = n - 1; if (stuck) *stuck = 1; }       // beyond the total: the reference keeps adding w[n-1]
        ancestor[t] = (int32_t)lo;
    }
}

// literal single-thread restatement of fast_slam_2.py:183-196 for weights the scan does not accept
__global__ void fs2_resample_serial(const double *w, int64_t n, double u0, int64_t m_begin, int64_t m_count,
                                    int32_t *ancestor, double *cum)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double inv = 1.0 / (double)n;
    double c = w[0];
    int64_t k = 0;
    for (int64_t i = 0; i < n; ++i) cum[i] = (i == 0) ? w[0] : __dadd_rn(cum[i - 1], w[i]);
    for (int64_t m = 0; m < m_begin + m_count; ++m) {
        double u = __dadd_rn(u0, __dmul_rn((double)m, inv));
        while (u > c) {
            if (k == n - 1 && !(w[k] > 0.0)) break;
            k = (k + 1 < n - 1) ? k + 1 : n - 1;
            c = __dadd_rn(c, w[k]);
        }
        if (m >= m_begin) ancestor[m - m_begin] = (int32_t)k;
    }
}

// ------------------------------------------------------------------------------------------------
// copy part: new slot m := old particle ancestor[m]  (deepcopy incl. weight, fast_slam_2.py:196-199)
//
// Systematic resampling keeps order, so ancestor[] is non-decreasing and the offspring of one particle
// are consecutive.  The FIRST offspring inherits the ancestor's map slot (no bytes move); every further
// offspring gets a copy in the slot of a particle that died.  #copies = #dead, and only those maps
// cross HBM -- instead of all P of a double-buffered gather.
// ------------------------------------------------------------------------------------------------
// Where an ancestor lives.  Three encodings of ancestor[m], selected by the launch:
//   local     (peers == nullptr, rec == nullptr)   a < P
//   staged    (rec != nullptr)                     a < P local, a >= P record (a - P) received from another GPU
//   peer      (peers != nullptr)                   a is a GLOBAL index; rank a / P owns it as local index a % P and
//                                                  its store is mapped into this process (CUDA IPC over NVLink)
#define FS2_MAX_PEERS 16
struct Fs2Peers {
    const double *x[FS2_MAX_PEERS], *y[FS2_MAX_PEERS], *yaw[FS2_MAX_PEERS], *w[FS2_MAX_PEERS], *lm[FS2_MAX_PEERS];
    const int32_t *count[FS2_MAX_PEERS], *slot[FS2_MAX_PEERS];
    int rank, world;
};

__device__ __forceinline__ bool fs2_anc_local(int a, int64_t P, const Fs2Peers *peers, int *local_idx)
{
    if (peers) {
        const int r = (int)(a / P);
        *local_idx = (int)(a - (int64_t)r * P);
        return r == peers->rank;
    }
    *local_idx = a;
    return a < P;
}

// used[s] = 1 for every map slot that must survive this resample untouched; extra[m] = 1 for every new particle
// that needs a copy.  The FIRST local offspring of a local ancestor inherits the ancestor's slot.
__global__ void __launch_bounds__(256)
fs2_gather_mark(const int32_t *__restrict__ anc, int64_t P, const Fs2Peers *peers, const int32_t *__restrict__ slot,
                int32_t *used, int32_t *extra)
{
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < P; m += (int64_t)gridDim.x * blockDim.x) {
        const int a = anc[m];
        int li;
        const bool local = fs2_anc_local(a, P, peers, &li);
        const bool first = local && ((m == 0) || (anc[m - 1] != a));
        extra[m] = first ? 0 : 1;
        if (first) used[slot[li]] = 1;
    }
}

// peer mode: a local particle with offspring on ANOTHER GPU is read by that GPU during its gather, so its slot
// must not be recycled in this round even if it has no offspring here
__global__ void __launch_bounds__(256)
fs2_gather_mark_global(const int32_t *__restrict__ anc_all, int64_t N, int64_t P, int rank, const int32_t *__restrict__ slot,
                       int32_t *used)
{
    const int64_t lo = (int64_t)rank * P, hi = lo + P;
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < N; m += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = anc_all[m];
        if (a >= lo && a < hi && (m == 0 || anc_all[m - 1] != a)) used[slot[a - lo]] = 1;
    }
}

// exclusive scan of two int32 flag arrays at once (extra[P], and "slot s is free" over the S >= P slots of the pool),
// three small kernels (block sums, prefix, apply)
#define FS2_ISCAN_B 2048
__global__ void __launch_bounds__(256)
fs2_iscan_sums(const int32_t *__restrict__ extra, const int32_t *__restrict__ used, int64_t P, int64_t S, int2 *bs)
{
    __shared__ int2 ws[8];
    int64_t base = (int64_t)blockIdx.x * FS2_ISCAN_B;
    int sa = 0, sd = 0;
    for (int j = threadIdx.x; j < FS2_ISCAN_B; j += 256) {
        int64_t i = base + j;
        if (i < P) sa += extra[i];
        if (i < S) sd += (used[i] == 0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sd += __shfl_xor_sync(0xffffffffu, sd, o); }
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = make_int2(sa, sd);
    __syncthreads();
    if (threadIdx.x == 0) {
        int2 t = make_int2(0, 0);
        for (int k = 0; k < 8; ++k) { t.x += ws[k].x; t.y += ws[k].y; }
        bs[blockIdx.x] = t;
    }
}

// single block: exclusive prefix of the per-block (extra, free) counts
__global__ void __launch_bounds__(1024) fs2_iscan_prefix(int2 *bs, int nb, int32_t *ncopies)
{
    __shared__ int2 ws[32];
    __shared__ int2 carry_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = make_int2(0, 0);
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        int i = base + threadIdx.x;
        int2 v = (i < nb) ? bs[i] : make_int2(0, 0);
        int2 inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int tx = __shfl_up_sync(0xffffffffu, inc.x, o), ty = __shfl_up_sync(0xffffffffu, inc.y, o);
            if (lane >= o) { inc.x += tx; inc.y += ty; }
        }
        if (lane == 31) ws[wid] = inc;
        __syncthreads();
        int2 off = carry_s;
        for (int k = 0; k < wid; ++k) { off.x += ws[k].x; off.y += ws[k].y; }
        if (i < nb) bs[i] = make_int2(off.x + inc.x - v.x, off.y + inc.y - v.y);
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = make_int2(off.x + inc.x, off.y + inc.y);
        __syncthreads();
    }
    if (threadIdx.x == 0) { ncopies[0] = carry_s.x; ncopies[1] = carry_s.y; }   // copies needed, free slots available
}

// tasks[r] = destination particle m of the r-th extra offspring; freeslot[r] = the r-th free slot of the pool
__global__ void __launch_bounds__(256)
fs2_iscan_apply(const int32_t *__restrict__ extra, const int32_t *__restrict__ used, int64_t P, int64_t S, const int2 *bs,
                int32_t *tasks, int32_t *freeslot)
{
    __shared__ int2 wsum[8];
    __shared__ int2 carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = bs[blockIdx.x];
    __syncthreads();
    int64_t base = (int64_t)blockIdx.x * FS2_ISCAN_B;
    for (int j0 = 0; j0 < FS2_ISCAN_B; j0 += 256) {
        int64_t i = base + j0 + threadIdx.x;
        int fa = 0, fd = 0;
        if (i < P) fa = extra[i];
        if (i < S) fd = (used[i] == 0);
        int ia = fa, id = fd;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int ta = __shfl_up_sync(0xffffffffu, ia, o), td = __shfl_up_sync(0xffffffffu, id, o);
            if (lane >= o) { ia += ta; id += td; }
        }
        if (lane == 31) wsum[wid] = make_int2(ia, id);
        __syncthreads();
        int2 off = carry;
        for (int k = 0; k < wid; ++k) { off.x += wsum[k].x; off.y += wsum[k].y; }
        if (fa) tasks[off.x + ia - 1] = (int32_t)i;
        if (fd) freeslot[off.y + id - 1] = (int32_t)i;
        __syncthreads();
        if (threadIdx.x == 255) { carry.x = off.x + ia; carry.y = off.y + id; }
        __syncthreads();
    }
}

// Staged particles arrive as records of `rstride` doubles: { x, y, yaw, w, count, 0, 0, 0, map rows ... }
// (fs2_pack_records / fs2_pull_records).
// pose / weight / count of every new slot, and the inherited map slot of first offspring
__global__ void __launch_bounds__(256)
fs2_gather_pose(const int32_t *__restrict__ anc, const int32_t *__restrict__ extra, int64_t P,
                const double *x, const double *y, const double *yaw, const double *w, const int32_t *count,
                const int32_t *slot, const double *rec, int64_t rstride, const Fs2Peers *peers,
                double *x2, double *y2, double *yaw2, double *w2, int32_t *count2, int32_t *slot2)
{
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < P; m += (int64_t)gridDim.x * blockDim.x) {
        const int a = anc[m];
        int li;
        if (fs2_anc_local(a, P, peers, &li)) {
            x2[m] = x[li]; y2[m] = y[li]; yaw2[m] = yaw[li]; w2[m] = w[li]; count2[m] = count[li];
            if (!extra[m]) slot2[m] = slot[li];
        } else if (peers) {
            const int r = (int)(a / P);              // loads from the owning GPU's memory over NVLink
            x2[m] = peers->x[r][li]; y2[m] = peers->y[r][li]; yaw2[m] = peers->yaw[r][li]; w2[m] = peers->w[r][li];
            count2[m] = peers->count[r][li];
        } else {
            const double *r = rec + (size_t)(a - P) * (size_t)rstride;
            x2[m] = r[0]; y2[m] = r[1]; yaw2[m] = r[2]; w2[m] = r[3]; count2[m] = (int32_t)r[4];
        }
    }
}

// one warp per extra offspring: copy the ancestor's live map (count * 48 B, 16 B per lane per access) from the
// local store, from a received record, or straight out of the owning GPU's store (peer loads over NVLink: the
// migration of the survivors' maps and the copy-on-resample gather are then the same kernel, and local copies
// overlap the NVLink pulls)
__global__ void __launch_bounds__(256)
fs2_gather_copy(const int32_t *__restrict__ tasks, const int32_t *__restrict__ freeslot, const int32_t *ncopies,
                const int32_t *__restrict__ anc, int64_t P, const int32_t *__restrict__ slot_old,
                const int32_t *__restrict__ count_old, const double *rec, int64_t rstride, const Fs2Peers *peers,
                double *lm, int lcap, int32_t *slot2)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int n = min(ncopies[0], ncopies[1]);      // the host checks copies <= free slots before launching
    const size_t stride = 6 * (size_t)lcap;
    for (int64_t r = warp; r < n; r += nwarps) {
        const int m = tasks[r];
        const int a = anc[m];
        const int dst_slot = freeslot[r];
        const int4 *src;
        int ng;                                        // 16-byte granules
        int li;
        if (fs2_anc_local(a, P, peers, &li)) {
            src = reinterpret_cast<const int4 *>(lm + (size_t)slot_old[li] * stride);
            ng = count_old[li] * 3;
        } else if (peers) {
            const int pr = (int)(a / P);
            src = reinterpret_cast<const int4 *>(peers->lm[pr] + (size_t)peers->slot[pr][li] * stride);
            ng = peers->count[pr][li] * 3;
        } else {
            const double *rr = rec + (size_t)(a - P) * (size_t)rstride;
            src = reinterpret_cast<const int4 *>(rr + 8);
            ng = (int)rr[4] * 3;
        }
        int4 *dst = reinterpret_cast<int4 *>(lm + (size_t)dst_slot * stride);
        int g = lane;
        for (; g + 96 < ng; g += 128) {                // 4 independent 16 B loads in flight per lane
            int4 v0 = __ldcs(src + g), v1 = __ldcs(src + g + 32), v2 = __ldcs(src + g + 64), v3 = __ldcs(src + g + 96);
            __stcs(dst + g, v0); __stcs(dst + g + 32, v1); __stcs(dst + g + 64, v2); __stcs(dst + g + 96, v3);
        }
        for (; g < ng; g += 32) __stcs(dst + g, __ldcs(src + g));
        if (lane == 0) slot2[m] = dst_slot;
    }
}
