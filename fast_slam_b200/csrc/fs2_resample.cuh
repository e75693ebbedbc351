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
//   3. fs2_scan_blockfunc    the composed parity function (A0, A1) of each block in units of 2^(e_b-52)
//   4. fs2_scan_chain        one warp walks the blocks with the EXACT running sum, 32 block functions per step
//                            (staged through shared memory 1024 at a time, warp scan of the compositions):
//                            blocks whose assumed binade holds at entry and exit are applied at once; the few
//                            that straddle a power of two (<= ~60 per scan) are added element by element
//   5. fs2_scan_emit         exact c_k for every element (parity scan inside the block, or serial for
//                            the straddling blocks)
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

__global__ void __launch_bounds__(FS2_SCAN_T)
fs2_scan_blockfunc(const double *__restrict__ w, int64_t n, const double *bsum, const double *bpre,
                   unsigned long long *A0, unsigned long long *A1, int *eb, int *mode)
{
    __shared__ Fs2Par wp[FS2_SCAN_T / 32];
    __shared__ int s_big;
    const int b = blockIdx.x;
    const int e = fs2_exponent(bpre[b]);
    if (threadIdx.x == 0) s_big = 0;
    __syncthreads();
    if (bsum[b] == 0.0) {           // non-negative weights summing to +0: nothing changes
        if (threadIdx.x == 0) { A0[b] = 0; A1[b] = 0; eb[b] = e; mode[b] = FS2_MODE_ZERO; }
        return;
    }
    if (e == INT_MIN) {             // running sum still zero / subnormal here: walk it
        if (threadIdx.x == 0) { A0[b] = 0; A1[b] = 0; eb[b] = e; mode[b] = FS2_MODE_SERIAL; }
        return;
    }
    int64_t base = (int64_t)b * FS2_SCAN_B + FS2_SCAN_EPT * threadIdx.x;
    Fs2Par f;
    f.e = f.o = 0ull;
    bool big = false;
#pragma unroll
    for (int j = 0; j < FS2_SCAN_EPT; ++j) {
        int64_t i = base + j;
        if (i < n) f = fs2_par_compose(f, fs2_par_of(w[i], e, &big));
    }
    if (big) s_big = 1;
    // ordered warp reduction (lane order = element order)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Fs2Par g;
        g.e = __shfl_down_sync(0xffffffffu, f.e, o);
        g.o = __shfl_down_sync(0xffffffffu, f.o, o);
        if (((threadIdx.x & 31) % (2 * o)) == 0) f = fs2_par_compose(f, g);
    }
    if ((threadIdx.x & 31) == 0) wp[threadIdx.x >> 5] = f;
    __syncthreads();
    if (threadIdx.x == 0) {
        Fs2Par t = wp[0];
        for (int k = 1; k < FS2_SCAN_T / 32; ++k) t = fs2_par_compose(t, wp[k]);
        A0[b] = t.e; A1[b] = t.o; eb[b] = e;
        mode[b] = s_big ? FS2_MODE_SERIAL : FS2_MODE_PARITY;
    }
}

// exact sequential walk of one block's weights (the reference's "particle_weight += w[k]")
__device__ __forceinline__ double fs2_walk(const double *w, int64_t lo, int64_t hi, double c)
{
    for (int64_t i = lo; i < hi; ++i) c = (i == 0) ? w[0] : __dadd_rn(c, w[i]);
    return c;
}

// exact running sum at every scan-block boundary.  The per-block functions are staged 1024 at a time in shared
// memory by the whole thread block (one coalesced round trip instead of one dependent global load per step); warp 0
// then composes 32 of them per step with a warp scan and applies them to the exact running sum at once: the prefix
// of blocks whose assumed binade holds at entry and exit is accepted, the first block that does not (it straddles
// a power of two, or the sum is still zero) is added element by element -- its 256 weights sit in registers and
// are handed round by shuffles.
#define FS2_CHAIN_TILE 1024
__global__ void __launch_bounds__(FS2_CHAIN_TILE)
fs2_scan_chain(const double *w, int64_t n, int nb, const unsigned long long *A0, const unsigned long long *A1,
               const int *eb, int *mode, double *cstart, double *total)
{
    __shared__ unsigned long long sA0[FS2_CHAIN_TILE], sA1[FS2_CHAIN_TILE];
    __shared__ int sE[FS2_CHAIN_TILE], sM[FS2_CHAIN_TILE];
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    double c = 0.0;                                   // the exact running sum, identical in every lane of warp 0
    for (int t0 = 0; t0 < nb; t0 += FS2_CHAIN_TILE) {
        const int tn = min(FS2_CHAIN_TILE, nb - t0);
        __syncthreads();                              // warp 0 is done with the previous tile
        if ((int)threadIdx.x < tn) {
            const int b = t0 + threadIdx.x;
            const int md = mode[b];
            sM[threadIdx.x] = md;
            sE[threadIdx.x] = eb[b];
            sA0[threadIdx.x] = (md == FS2_MODE_PARITY) ? A0[b] : 0ull;
            sA1[threadIdx.x] = (md == FS2_MODE_PARITY) ? A1[b] : 0ull;
        }
        __syncthreads();
        if (threadIdx.x >= 32) continue;
        int b0 = 0;                                   // position inside the tile
        while (b0 < tn) {
            const int i = b0 + lane;
            const bool valid = i < tn;
            Fs2Par f;
            f.e = f.o = 0ull;
            int e = INT_MIN, md = FS2_MODE_SERIAL;
            if (valid) { md = sM[i]; e = sE[i]; f.e = sA0[i]; f.o = sA1[i]; }
            // inclusive scan of the compositions (lane order = block order)
            Fs2Par inc = f;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                Fs2Par g;
                g.e = __shfl_up_sync(full, inc.e, o);
                g.o = __shfl_up_sync(full, inc.o, o);
                if (lane >= o) inc = fs2_par_compose(g, inc);
            }
            Fs2Par ex;                                // composition of the lanes before me
            ex.e = __shfl_up_sync(full, inc.e, 1);
            ex.o = __shfl_up_sync(full, inc.o, 1);
            if (lane == 0) ex.e = ex.o = 0ull;
            const int ec = fs2_exponent(c);
            const unsigned long long bits = (unsigned long long)__double_as_longlong(c);
            const unsigned long long C = (bits & 0x000fffffffffffffull) | 0x0010000000000000ull;
            const bool odd = (C & 1ull) != 0ull;
            const unsigned long long Cin = C + (odd ? ex.o : ex.e), Cout = C + (odd ? inc.o : inc.e);
            // a block is fine if it changes nothing, or if it was composed for the binade the sum is in and stays there
            const bool ok = valid && (md == FS2_MODE_ZERO ||
                                      (md == FS2_MODE_PARITY && ec != INT_MIN && e == ec && Cout < 0x0020000000000000ull));
            const unsigned okm = __ballot_sync(full, ok);
            const int nacc = (okm == full) ? 32 : (__ffs(~okm) - 1);     // accepted prefix
            if (lane < nacc) {
                const unsigned long long vb = (bits & 0xfff0000000000000ull) | (Cin & 0x000fffffffffffffull);
                cstart[t0 + i] = (ec == INT_MIN) ? c : __longlong_as_double((long long)vb);   // all-zero blocks before the sum starts
            }
            if (nacc > 0) {
                const unsigned long long Cl = __shfl_sync(full, Cout, nacc - 1);
                if (ec != INT_MIN) {
                    const unsigned long long vb = (bits & 0xfff0000000000000ull) | (Cl & 0x000fffffffffffffull);
                    c = __longlong_as_double((long long)vb);
                }
            }
            b0 += nacc;
            if (nacc < 32 && b0 < tn) {
                // walk block t0 + b0 exactly (the reference's "particle_weight += w[k]")
                const int64_t lo = (int64_t)(t0 + b0) * FS2_SCAN_B;
                const int64_t hi = lo + FS2_SCAN_B < n ? lo + FS2_SCAN_B : n;
                if (lane == 0) { cstart[t0 + b0] = c; mode[t0 + b0] = FS2_MODE_SERIAL; }
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
    }
    if (threadIdx.x == 0) *total = c;
}

__global__ void __launch_bounds__(FS2_SCAN_T)
fs2_scan_emit(const double *__restrict__ w, int64_t n, const int *eb, const int *mode, const double *cstart, double *cum)
{
    __shared__ Fs2Par wp[FS2_SCAN_T / 32];
    const int b = blockIdx.x;
    const int md = mode[b];
    const int64_t lo = (int64_t)b * FS2_SCAN_B;
    const int64_t hi = lo + FS2_SCAN_B < n ? lo + FS2_SCAN_B : n;
    const double c0 = cstart[b];
    if (md == FS2_MODE_ZERO) {
        for (int64_t i = lo + threadIdx.x; i < hi; i += FS2_SCAN_T) cum[i] = (i == 0) ? w[0] : c0;
        return;
    }
    if (md == FS2_MODE_SERIAL) {
        if (threadIdx.x == 0) {
            double c = c0;
            for (int64_t i = lo; i < hi; ++i) { c = (i == 0) ? w[0] : __dadd_rn(c, w[i]); cum[i] = c; }
        }
        return;
    }
    const int e = eb[b];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int64_t base = lo + FS2_SCAN_EPT * threadIdx.x;
    Fs2Par loc[FS2_SCAN_EPT];
    Fs2Par f;
    f.e = f.o = 0ull;
    bool big = false;
#pragma unroll
    for (int j = 0; j < FS2_SCAN_EPT; ++j) {
        if (base + j < hi) f = fs2_par_compose(f, fs2_par_of(w[base + j], e, &big));
        loc[j] = f;                           // inclusive within the thread
    }
    // inclusive ordered scan across the warp
    Fs2Par inc = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Fs2Par g;
        g.e = __shfl_up_sync(0xffffffffu, inc.e, o);
        g.o = __shfl_up_sync(0xffffffffu, inc.o, o);
        if (lane >= o) inc = fs2_par_compose(g, inc);
    }
    if (lane == 31) wp[wid] = inc;
    __syncthreads();
    // exclusive prefix of this thread = (warps before) o (lanes before)
    Fs2Par pre;
    pre.e = pre.o = 0ull;
    for (int k = 0; k < wid; ++k) pre = fs2_par_compose(pre, wp[k]);
    Fs2Par lp;
    lp.e = __shfl_up_sync(0xffffffffu, inc.e, 1);
    lp.o = __shfl_up_sync(0xffffffffu, inc.o, 1);
    if (lane > 0) pre = fs2_par_compose(pre, lp);
    const unsigned long long bits0 = (unsigned long long)__double_as_longlong(c0);
    const unsigned long long C0 = (bits0 & 0x000fffffffffffffull) | 0x0010000000000000ull;
    const bool odd = (C0 & 1ull) != 0ull;
#pragma unroll
    for (int j = 0; j < FS2_SCAN_EPT; ++j) {
        if (base + j < hi) {
            Fs2Par t = fs2_par_compose(pre, loc[j]);
            unsigned long long C = C0 + (odd ? t.o : t.e);
            unsigned long long bits = (bits0 & 0xfff0000000000000ull) | (C & 0x000fffffffffffffull);
            cum[base + j] = __longlong_as_double((long long)bits);
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
