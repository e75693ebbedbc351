// fs2_known.cuh -- map clustering on the device (row N1 of SURVEY.md 8f):
//     LandmarkUtils.update_known_landmarks   fast_slam_2/utils/landmark_utils.py:120-144
//     GeometryUtils.cluster_points           fast_slam_2/utils/geometry_utils.py:26-62   (sklearn DBSCAN + means)
//
// The reference hands every landmark of every particle (N = sum of the map lengths; 2.7e8 points at 2^20
// particles x 256 landmarks) to DBSCAN(eps = 0.5, min_samples = int(0.7 N / P)) and averages each cluster.
// DBSCAN's result does not depend on its traversal (oracle/known_landmarks_oracle.py):
//     core point  : at least min_samples points p (itself included) with dx*dx + dy*dy <= eps*eps
//     cluster     : connected component of the core points, numbered by its lowest core point index
//     other points: lowest-numbered cluster among their core neighbours, else noise
// so it can be evaluated on a grid.  Cells have side h = eps/16 and live in 16 x 16 tiles (side eps) that are
// allocated through a hash on the tile coordinate.  One pass over the maps counts the points of every cell
// (count, lowest point index, fixed-point coordinate sums).  A cell pair is
//     inside  : every point of one is within eps of every point of the other  ((|dx|+1)^2 + (|dy|+1)^2 < 256)
//     outside : no two points are                                             ((|dx|-1)^2 + (|dy|-1)^2 > 256)
//     partial : anything else
// with 1/16 of a cell of margin either way, so the classification is safe against the rounding of the point
// coordinates.  Per cell, LB = points in inside cells, UB = LB + points in partial cells:
//     LB >= min_samples  -> every point of the cell is core        (KL_ALLCORE)
//     UB <  min_samples  -> none is                                (KL_NONE)
//     otherwise          -> decided point by point                 (KL_AMBIG)
// At the sizes this is built for almost every occupied cell is ALLCORE and the whole clustering happens on
// cells: union-find over cell pairs in range, cluster sums from the per-cell sums.  Whatever cannot be decided
// on cells -- AMBIG and NONE cells, their partial neighbours, and partial pairs of core cells that are not
// already connected -- is "involved": a second pass over the maps copies those points out (grouped by cell)
// and the exact float64 test of the reference is evaluated on them.  The result is exact for every input; the
// cost of the exact part grows with the number of involved points (capacity: fs2_kl workspace, FS2_ERR_NOMEM).
//
// Sums are kept in integers (cell index + 37-bit offset inside the cell), so the centroids do not depend on
// the order of the atomics; they differ from the reference's sequential float64 mean by its own rounding.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define KL_TS 16
#define KL_TC (KL_TS * KL_TS)
#define KL_WR 17
#define KL_WD (2 * KL_WR + 1)
#define KL_NB 5
#define KL_WIN (KL_NB * KL_TS)   // 80: side of the staged window of a tile
#define KL_FIX 37
#define KL_QCAP 640              // partial cells queued per point / per cell (at most ~330 of the 1225 are partial)

enum { KL_EMPTY = 0, KL_ALLCORE = 1, KL_AMBIG = 2, KL_NONE = 3 };
enum { KL_ERR_TILES = 1, KL_ERR_NONFINITE = 2, KL_ERR_RANGE = 4, KL_ERR_CELL_COUNT = 8, KL_ERR_CLUSTERS = 16, KL_ERR_WORK = 32 };
#define KL_KEY_EMPTY 0xffffffffffffffffull
#define KL_NOIDX 0xffffffffffffffffull
#define KL_NOCELL 0xffffffffu

typedef unsigned long long kl_u64;

__constant__ unsigned char kl_cls[KL_WD * KL_WD];   // 0 outside, 1 inside, 2 partial

struct KlGrid {
    kl_u64 *hkeys;          // [tcap] tile coordinate, or KL_KEY_EMPTY; the slot index is the tile id
    unsigned hmask;         // tcap - 1
    unsigned *nbr;          // [tcap][25] tile ids of the 5 x 5 neighbourhood
    // per cell, [tcap * 256]
    unsigned *cnt;
    kl_u64 *minidx;         // lowest point index of the cell
    kl_u64 *sx, *sy;        // sum of the points' fixed-point offsets inside the cell
    unsigned char *status, *inv;
    unsigned *parent;
    unsigned *ncore;
    kl_u64 *mincore;        // lowest core point index of the cell
    kl_u64 *rootmin;        // per root: lowest core point index of the cluster (= its label order)
    unsigned *off, *cursor; // segment of the cell's points in the compacted arrays
    unsigned *cid;          // per root: dense cluster id
    int *err;
    kl_u64 *work;           // exact distance tests charged so far by the point-level kernels ...
    kl_u64 work_limit;      // ... and the budget: past it they stop and the call fails (KL_ERR_WORK) instead of running on
    double h, inv_h;
    double eps2;
    long long min_samples;
    int pow2;               // h is a power of two
};

struct KlPts {              // compacted points of the involved cells, grouped by cell
    double *x, *y;
    kl_u64 *idx;
    unsigned *cell;
    unsigned char *flag;    // 0 noise, 1 core, 2 border
    unsigned *label;        // root cell of a border point
    unsigned cap;
};

struct KlAcc {              // per cluster
    kl_u64 *n;
    long long *ax, *ay;     // sum of cell indices (one per point)
    kl_u64 *bxl, *byl;      // 128-bit sums of the offsets: low words ...
    long long *bxh, *byh;   // ... and high words
    kl_u64 *minidx;
    unsigned *count;        // number of clusters
    unsigned cap;
};

struct KlSrcState {         // the filter's maps: particle p, landmark j -> point index base0 + base[p] + j
    const double *lm;
    const int32_t *slot, *count;
    const kl_u64 *base;
    kl_u64 base0;           // points of the shards before this one
    int64_t P;
    int32_t lcap;
};

struct KlSrcFlat {          // a plain [n][2] array
    const double *xy;
    int64_t n;
};

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ kl_u64 kl_key(int tx, int ty)
{
    return ((kl_u64)(unsigned)(tx + (1 << 30)) << 32) | (kl_u64)(unsigned)(ty + (1 << 30));
}

__device__ __forceinline__ void kl_key_decode(kl_u64 k, int &tx, int &ty)
{
    tx = (int)(unsigned)(k >> 32) - (1 << 30);
    ty = (int)(unsigned)(k & 0xffffffffu) - (1 << 30);
}

__device__ __forceinline__ unsigned kl_hash(kl_u64 k, unsigned mask)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned)k & mask;
}

__device__ __forceinline__ unsigned kl_tile_find(const KlGrid &g, int tx, int ty)
{
    const kl_u64 key = kl_key(tx, ty);
    unsigned s = kl_hash(key, g.hmask);
    for (unsigned n = 0; n <= g.hmask; ++n) {
        const kl_u64 k = g.hkeys[s];
        if (k == key) return s;
        if (k == KL_KEY_EMPTY) return KL_NOCELL;
        s = (s + 1) & g.hmask;
    }
    return KL_NOCELL;
}

__device__ __forceinline__ unsigned kl_tile_insert(const KlGrid &g, int tx, int ty)
{
    const kl_u64 key = kl_key(tx, ty);
    unsigned s = kl_hash(key, g.hmask);
    for (unsigned n = 0; n <= g.hmask; ++n) {
        kl_u64 k = g.hkeys[s];
        if (k == KL_KEY_EMPTY) k = atomicCAS(&g.hkeys[s], KL_KEY_EMPTY, key);
        if (k == key || k == KL_KEY_EMPTY) return s;
        s = (s + 1) & g.hmask;
    }
    atomicOr(g.err, KL_ERR_TILES);
    return KL_NOCELL;
}

// cell of a point: tile coordinate, cell inside the tile, fixed-point offsets inside the cell
__device__ __forceinline__ bool kl_locate(const KlGrid &g, double x, double y, int &tx, int &ty, int &lc, kl_u64 &qx, kl_u64 &qy)
{
    if (!isfinite(x) || !isfinite(y)) { atomicOr(g.err, KL_ERR_NONFINITE); return false; }
    // h = eps/16 is a power of two for the reference's eps = 0.5: scaling by 1/h is then exact and the same as
    // dividing, without the four double divisions per point
    double sxh, syh;
    if (g.pow2) { sxh = x * g.inv_h; syh = y * g.inv_h; } else { sxh = x / g.h; syh = y / g.h; }
    const double fx = floor(sxh), fy = floor(syh);
    if (fabs(fx) > 1.0e9 || fabs(fy) > 1.0e9) { atomicOr(g.err, KL_ERR_RANGE); return false; }
    const long long gx = (long long)fx, gy = (long long)fy;
    tx = (int)(gx >> 4); ty = (int)(gy >> 4);
    lc = (int)(gy & 15) * KL_TS + (int)(gx & 15);
    const double lim = (double)((1ull << KL_FIX) - 1ull);
    double ux, uy;
    if (g.pow2) { ux = sxh - fx; uy = syh - fy; } else { ux = (x - fx * g.h) / g.h; uy = (y - fy * g.h) / g.h; }
    ux = rint(ux * (double)(1ull << KL_FIX)); uy = rint(uy * (double)(1ull << KL_FIX));
    ux = fmin(fmax(ux, 0.0), lim); uy = fmin(fmax(uy, 0.0), lim);
    qx = (kl_u64)ux; qy = (kl_u64)uy;
    return true;
}

// cell at (X, Y) relative to the lower corner of `tile`, X and Y in [-32, 48)
__device__ __forceinline__ unsigned kl_cell_rel(const KlGrid &g, unsigned tile, int X, int Y)
{
    const int tdx = ((X + 2 * KL_TS) >> 4), tdy = ((Y + 2 * KL_TS) >> 4);      // 0..4
    const unsigned nb = g.nbr[tile * (KL_NB * KL_NB) + tdy * KL_NB + tdx];
    if (nb == KL_NOCELL) return KL_NOCELL;
    return nb * KL_TC + (unsigned)((Y & 15) * KL_TS + (X & 15));
}

__device__ __forceinline__ unsigned kl_find(unsigned *parent, unsigned a)
{
    while (true) {
        const unsigned p = *(volatile unsigned *)&parent[a];
        if (p == a) return a;
        const unsigned gp = *(volatile unsigned *)&parent[p];
        if (gp != p) atomicCAS(&parent[a], p, gp);      // path halving
        a = p;
    }
}

__device__ __forceinline__ void kl_union(unsigned *parent, unsigned a, unsigned b)
{
    while (true) {
        a = kl_find(parent, a);
        b = kl_find(parent, b);
        if (a == b) return;
        if (a < b) { const unsigned t = a; a = b; b = t; }
        if (atomicCAS(&parent[a], a, b) == a) return;   // the larger root hangs under the smaller
    }
}

// charge `tests` exact tests to the budget (one lane of the warp does it); false = budget exhausted
__device__ __forceinline__ bool kl_charge(const KlGrid &g, kl_u64 tests, int lane)
{
    kl_u64 before = 0;
    if (lane == 0) before = atomicAdd(g.work, tests);
    before = __shfl_sync(0xffffffffu, before, 0);
    if (before + tests > g.work_limit) {
        if (lane == 0) atomicOr(g.err, KL_ERR_WORK);
        return false;
    }
    return true;
}

__device__ __forceinline__ bool kl_near(double ax, double ay, double bx, double by, double eps2)
{
    const double dx = __dadd_rn(ax, -bx), dy = __dadd_rn(ay, -by);
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) <= eps2;    // sklearn's reduced distance, no fma
}

// ------------------------------------------------------------------------------------------------ exclusive scans
struct KlInCount { const int32_t *c; __device__ kl_u64 operator()(long long i) const { return (kl_u64)c[i]; } };
struct KlInInvolved {
    const unsigned *cnt; const unsigned char *inv; const kl_u64 *hkeys;
    __device__ kl_u64 operator()(long long i) const { return (hkeys[i >> 8] != KL_KEY_EMPTY && inv[i]) ? (kl_u64)cnt[i] : 0ull; }
};

template <class In>
__global__ void __launch_bounds__(256) kl_scan_sums(In in, long long n, kl_u64 *bsum)
{
    __shared__ kl_u64 ws[8];
    const long long base = (long long)blockIdx.x * 1024 + 4 * threadIdx.x;
    kl_u64 s = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) if (base + j < n) s += in(base + j);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { kl_u64 t = 0; for (int k = 0; k < 8; ++k) t += ws[k]; bsum[blockIdx.x] = t; }
}

__global__ void __launch_bounds__(1024) kl_scan_prefix(kl_u64 *bsum, int nb, kl_u64 *total)
{
    __shared__ kl_u64 ws[32];
    __shared__ kl_u64 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const kl_u64 v = (i < nb) ? bsum[i] : 0ull;
        kl_u64 s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const kl_u64 t = __shfl_up_sync(0xffffffffu, s, o); if ((threadIdx.x & 31) >= o) s += t; }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            kl_u64 t = ws[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const kl_u64 u = __shfl_up_sync(0xffffffffu, t, o); if (threadIdx.x >= o) t += u; }
            ws[threadIdx.x] = t;
        }
        __syncthreads();
        const kl_u64 before = carry + ((threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0ull) + s - v;
        if (i < nb) bsum[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

template <class In, class OutT>
__global__ void __launch_bounds__(256) kl_scan_apply(In in, long long n, const kl_u64 *bsum, OutT *out)
{
    __shared__ kl_u64 ws[8];
    const long long base = (long long)blockIdx.x * 1024 + 4 * threadIdx.x;
    kl_u64 v[4], s = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[j] = (base + j < n) ? in(base + j) : 0ull; s += v[j]; }
    kl_u64 inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const kl_u64 t = __shfl_up_sync(0xffffffffu, inc, o); if ((threadIdx.x & 31) >= o) inc += t; }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = inc;
    __syncthreads();
    kl_u64 pre = bsum[blockIdx.x] + inc - s;
    for (int k = 0; k < (int)(threadIdx.x >> 5); ++k) pre += ws[k];
#pragma unroll
    for (int j = 0; j < 4; ++j) { if (base + j < n) out[base + j] = (OutT)pre; pre += v[j]; }
}

// ------------------------------------------------------------------------------------------------ passes over the points
// Pass 1 goes through a per-block cache in shared memory: 8192 direct-mapped entries (cell id -> count and
// offset sums).  The first cell to hash to an entry owns it for the life of the block -- cells are met in
// proportion to their population, so the dense ones get there first -- and the points of the other cells fall
// through to global atomics.  4 global atomics per point (1.1e9 at 2^20 x 256: 8.3 ms, atomic-bound) become
// shared-memory atomics plus one flush per entry (3.5 ms; a cell may use either of two entries).  The lowest index needs no atomic once the cell has
// seen an earlier point: a read of the current minimum decides.  (Tried and slower, 4.2 ms: splitting the sums
// into 32-bit halves so that every shared atomic is a native one -- seven per point instead of three.)
#define KL_CE 8192
#define KL_CACHE_BYTES (KL_CE * 24)
#define KL_COUNT_THREADS 1024

struct KlCache { unsigned *key, *cnt; kl_u64 *sx, *sy; };

__device__ __forceinline__ KlCache kl_cache_init(unsigned char *smem)
{
    KlCache ch;
    ch.sx = reinterpret_cast<kl_u64 *>(smem);
    ch.sy = ch.sx + KL_CE;
    ch.key = reinterpret_cast<unsigned *>(ch.sy + KL_CE);
    ch.cnt = ch.key + KL_CE;
    for (int i = threadIdx.x; i < KL_CE; i += blockDim.x) { ch.key[i] = KL_NOCELL; ch.cnt[i] = 0u; ch.sx[i] = 0ull; ch.sy[i] = 0ull; }
    __syncthreads();
    return ch;
}

__device__ __forceinline__ void kl_count_point(const KlGrid &g, const KlCache &ch, double x, double y, kl_u64 idx)
{
    int tx, ty, lc; kl_u64 qx, qy;
    if (!kl_locate(g, x, y, tx, ty, lc, qx, qy)) return;
    const unsigned t = kl_tile_insert(g, tx, ty);
    if (t == KL_NOCELL) return;
    const unsigned c = t * KL_TC + (unsigned)lc;
    if (idx < *(volatile kl_u64 *)&g.minidx[c]) atomicMin(&g.minidx[c], idx);
    unsigned e = (c * 2654435761u) >> (32 - 13);                  // KL_CE = 2^13; two candidate entries per cell
    unsigned k = *(volatile unsigned *)&ch.key[e];
    if (k == KL_NOCELL) { k = atomicCAS(&ch.key[e], KL_NOCELL, c); if (k == KL_NOCELL) k = c; }
    if (k != c) {
        e = ((c ^ 0x9e3779b9u) * 2246822519u) >> (32 - 13);
        k = *(volatile unsigned *)&ch.key[e];
        if (k == KL_NOCELL) { k = atomicCAS(&ch.key[e], KL_NOCELL, c); if (k == KL_NOCELL) k = c; }
    }
    if (k == c) {
        atomicAdd(&ch.cnt[e], 1u);
        atomicAdd(&ch.sx[e], qx);
        atomicAdd(&ch.sy[e], qy);
    } else {
        atomicAdd(&g.cnt[c], 1u);
        atomicAdd(&g.sx[c], qx);
        atomicAdd(&g.sy[c], qy);
    }
}

__device__ __forceinline__ void kl_cache_flush(const KlGrid &g, const KlCache &ch)
{
    __syncthreads();
    for (int i = threadIdx.x; i < KL_CE; i += blockDim.x) {
        const unsigned c = ch.key[i];
        if (c == KL_NOCELL) continue;
        atomicAdd(&g.cnt[c], ch.cnt[i]);
        atomicAdd(&g.sx[c], ch.sx[i]);
        atomicAdd(&g.sy[c], ch.sy[i]);
    }
}

__global__ void __launch_bounds__(KL_COUNT_THREADS, 1) kl_count_state_kernel(KlSrcState s, KlGrid g)
{
    extern __shared__ __align__(16) unsigned char kl_smem[];
    const KlCache ch = kl_cache_init(kl_smem);
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < s.P; p += nw) {
        const int cnt = s.count[p];
        const kl_u64 base = s.base0 + s.base[p];
        const double *lm = s.lm + (size_t)s.slot[p] * 6 * (size_t)s.lcap;
        for (int j0 = 0; j0 < cnt; j0 += 128) {           // four loads in flight per lane
            double2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + 32 * u + lane;
                if (j < cnt) v[u] = *reinterpret_cast<const double2 *>(lm + 6 * (size_t)j);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + 32 * u + lane;
                if (j < cnt) kl_count_point(g, ch, v[u].x, v[u].y, base + (kl_u64)j);
            }
        }
    }
    kl_cache_flush(g, ch);
}

__global__ void __launch_bounds__(KL_COUNT_THREADS, 1) kl_count_flat_kernel(KlSrcFlat s, KlGrid g)
{
    extern __shared__ __align__(16) unsigned char kl_smem[];
    const KlCache ch = kl_cache_init(kl_smem);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < s.n; i += (long long)gridDim.x * blockDim.x) {
        const double2 v = *reinterpret_cast<const double2 *>(s.xy + 2 * i);
        kl_count_point(g, ch, v.x, v.y, (kl_u64)i);
    }
    kl_cache_flush(g, ch);
}

struct KlCompactOp {
    KlGrid g;
    KlPts pts;
    __device__ void operator()(double x, double y, kl_u64 idx) const
    {
        int tx, ty, lc; kl_u64 qx, qy;
        if (!kl_locate(g, x, y, tx, ty, lc, qx, qy)) return;
        const unsigned t = kl_tile_find(g, tx, ty);
        if (t == KL_NOCELL) return;
        const unsigned c = t * KL_TC + (unsigned)lc;
        if (!g.inv[c]) return;
        const unsigned pos = g.off[c] + atomicAdd(&g.cursor[c], 1u);
        if (pos >= pts.cap) return;
        pts.x[pos] = x; pts.y[pos] = y; pts.idx[pos] = idx; pts.cell[pos] = c;
    }
};

template <class Op>
__global__ void __launch_bounds__(256) kl_pass_state(KlSrcState s, Op op)
{
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < s.P; p += nw) {
        const int cnt = s.count[p];
        const kl_u64 base = s.base0 + s.base[p];
        const double *lm = s.lm + (size_t)s.slot[p] * 6 * (size_t)s.lcap;
        for (int j = lane; j < cnt; j += 32) {
            const double2 v = *reinterpret_cast<const double2 *>(lm + 6 * (size_t)j);
            op(v.x, v.y, base + (kl_u64)j);
        }
    }
}

template <class Op>
__global__ void __launch_bounds__(256) kl_pass_flat(KlSrcFlat s, Op op)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < s.n; i += (long long)gridDim.x * blockDim.x) {
        const double2 v = *reinterpret_cast<const double2 *>(s.xy + 2 * i);
        op(v.x, v.y, (kl_u64)i);
    }
}

// ------------------------------------------------------------------------------------------------ cell level
__global__ void __launch_bounds__(256) kl_nbr_kernel(KlGrid g)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned tile = i / (KL_NB * KL_NB), k = i % (KL_NB * KL_NB);
    if (tile > g.hmask) return;
    const kl_u64 key = g.hkeys[tile];
    if (key == KL_KEY_EMPTY) return;
    int tx, ty;
    kl_key_decode(key, tx, ty);
    g.nbr[i] = kl_tile_find(g, tx + (int)(k % KL_NB) - 2, ty + (int)(k / KL_NB) - 2);
    if (k == 0) atomicAdd(g.err + 1, 1);                 // occupied tiles (the host grows the grid past half full)
}

// one block per tile: LB / UB of every cell from the staged 80 x 80 window of counts
__global__ void __launch_bounds__(KL_TC) kl_classify_kernel(KlGrid g)
{
    __shared__ unsigned wc[KL_WIN * KL_WIN];
    const unsigned tile = blockIdx.x;
    if (g.hkeys[tile] == KL_KEY_EMPTY) return;
    for (int i = threadIdx.x; i < KL_WIN * KL_WIN; i += KL_TC) {
        const int wy = i / KL_WIN, wx = i % KL_WIN;
        const unsigned nb = g.nbr[tile * (KL_NB * KL_NB) + (wy >> 4) * KL_NB + (wx >> 4)];
        wc[i] = (nb == KL_NOCELL) ? 0u : g.cnt[nb * KL_TC + (unsigned)((wy & 15) * KL_TS + (wx & 15))];
    }
    __syncthreads();
    const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
    const unsigned c = tile * KL_TC + threadIdx.x;
    const unsigned own = wc[(2 * KL_TS + ly) * KL_WIN + 2 * KL_TS + lx];
    unsigned char st = KL_EMPTY;
    if (own) {
        kl_u64 lb = 0, ub = 0;
        for (int dy = -KL_WR; dy <= KL_WR; ++dy) {
            const unsigned *row = &wc[(2 * KL_TS + ly + dy) * KL_WIN + 2 * KL_TS + lx - KL_WR];
            const unsigned char *cl = &kl_cls[(dy + KL_WR) * KL_WD];
            for (int dx = 0; dx < KL_WD; ++dx) {
                const unsigned n = row[dx];
                const unsigned k = cl[dx];
                lb += (k == 1u) ? n : 0u;
                ub += (k != 0u) ? n : 0u;
            }
        }
        st = ((long long)lb >= g.min_samples) ? KL_ALLCORE : (((long long)ub < g.min_samples) ? KL_NONE : KL_AMBIG);
        if (own >= (1u << 26)) atomicOr(g.err, KL_ERR_CELL_COUNT);
    }
    g.status[c] = st;
    g.inv[c] = 0;
    g.parent[c] = c;
    g.ncore[c] = (st == KL_ALLCORE) ? own : 0u;
    g.mincore[c] = (st == KL_ALLCORE) ? g.minidx[c] : KL_NOIDX;
    g.rootmin[c] = KL_NOIDX;
    g.cursor[c] = 0;
    g.cid[c] = KL_NOCELL;
}

// neighbouring core cells are always within eps of each other: this alone connects a dense cloud
__global__ void __launch_bounds__(KL_TC) kl_union_adjacent_kernel(KlGrid g)
{
    const unsigned tile = blockIdx.x;
    if (g.hkeys[tile] == KL_KEY_EMPTY) return;
    const unsigned c = tile * KL_TC + threadIdx.x;
    if (g.status[c] != KL_ALLCORE) return;
    const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
    const int ddx[4] = {1, 0, 1, -1}, ddy[4] = {0, 1, 1, 1};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned c2 = kl_cell_rel(g, tile, lx + ddx[k], ly + ddy[k]);
        if (c2 != KL_NOCELL && g.status[c2] == KL_ALLCORE) kl_union(g.parent, c, c2);
    }
}

__global__ void __launch_bounds__(KL_TC) kl_flatten_kernel(KlGrid g)
{
    const unsigned tile = blockIdx.x;
    if (g.hkeys[tile] == KL_KEY_EMPTY) return;
    const unsigned c = tile * KL_TC + threadIdx.x;
    if (g.status[c] != KL_EMPTY) g.parent[c] = kl_find(g.parent, c);
}

// which cells need their points: see the header.  Also joins core cells in "inside" range that the adjacency
// pass left in different components.
__global__ void __launch_bounds__(KL_TC) kl_mark_kernel(KlGrid g)
{
    __shared__ unsigned char ws[KL_WIN * KL_WIN];
    const unsigned tile = blockIdx.x;
    if (g.hkeys[tile] == KL_KEY_EMPTY) return;
    for (int i = threadIdx.x; i < KL_WIN * KL_WIN; i += KL_TC) {
        const int wy = i / KL_WIN, wx = i % KL_WIN;
        const unsigned nb = g.nbr[tile * (KL_NB * KL_NB) + (wy >> 4) * KL_NB + (wx >> 4)];
        ws[i] = (nb == KL_NOCELL) ? (unsigned char)KL_EMPTY : g.status[nb * KL_TC + (unsigned)((wy & 15) * KL_TS + (wx & 15))];
    }
    __syncthreads();
    const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
    const unsigned c = tile * KL_TC + threadIdx.x;
    const unsigned char st = ws[(2 * KL_TS + ly) * KL_WIN + 2 * KL_TS + lx];
    if (st == KL_EMPTY) return;
    if (st == KL_ALLCORE) {
        const unsigned rc = g.parent[c];
        for (int dy = -KL_WR; dy <= KL_WR; ++dy)
            for (int dx = -KL_WR; dx <= KL_WR; ++dx) {
                if (ws[(2 * KL_TS + ly + dy) * KL_WIN + 2 * KL_TS + lx + dx] != KL_ALLCORE) continue;
                const unsigned k = kl_cls[(dy + KL_WR) * KL_WD + dx + KL_WR];
                if (k == 0u) continue;
                const unsigned c2 = kl_cell_rel(g, tile, lx + dx, ly + dy);
                if (c2 <= c) continue;                                    // each pair once
                if (*(volatile unsigned *)&g.parent[c2] == rc) continue;  // (a stale answer only costs work)
                if (k == 1u) kl_union(g.parent, c, c2);
                else { g.inv[c] = 1; g.inv[c2] = 1; }                     // to be settled on the points
            }
        return;
    }
    bool wanted = (st == KL_AMBIG);
    if (!wanted) {      // a NONE cell only matters if core points may be near: its points can be border points
        for (int dy = -KL_WR; dy <= KL_WR && !wanted; ++dy)
            for (int dx = -KL_WR; dx <= KL_WR; ++dx) {
                const unsigned char s2 = ws[(2 * KL_TS + ly + dy) * KL_WIN + 2 * KL_TS + lx + dx];
                if ((s2 == KL_ALLCORE || s2 == KL_AMBIG) && kl_cls[(dy + KL_WR) * KL_WD + dx + KL_WR] != 0u) { wanted = true; break; }
            }
    }
    if (!wanted) return;
    g.inv[c] = 1;
    for (int dy = -KL_WR; dy <= KL_WR; ++dy)
        for (int dx = -KL_WR; dx <= KL_WR; ++dx) {
            const unsigned char s2 = ws[(2 * KL_TS + ly + dy) * KL_WIN + 2 * KL_TS + lx + dx];
            if (s2 == KL_EMPTY || kl_cls[(dy + KL_WR) * KL_WD + dx + KL_WR] != 2u) continue;
            if (st == KL_NONE && s2 == KL_NONE) continue;                 // no core points there
            g.inv[kl_cell_rel(g, tile, lx + dx, ly + dy)] = 1;
        }
}

// ------------------------------------------------------------------------------------------------ point level
// window scan shared by the point-level kernels: every lane looks at the window cells lane, lane+32, ...;
// `inside(c2)` is called for occupied inside cells, partial ones are queued for the whole warp
template <class Inside>
__device__ __forceinline__ int kl_scan_window(const KlGrid &g, unsigned tile, int lx, int ly, int lane, unsigned *queue, Inside inside)
{
    int nq = 0;
    for (int w0 = 0; w0 < KL_WD * KL_WD; w0 += 32) {
        const int w = w0 + lane;
        unsigned c2 = KL_NOCELL;
        bool part = false;
        if (w < KL_WD * KL_WD) {
            const unsigned k = kl_cls[w];
            if (k != 0u) {
                c2 = kl_cell_rel(g, tile, lx + (w % KL_WD) - KL_WR, ly + (w / KL_WD) - KL_WR);
                if (c2 != KL_NOCELL && g.cnt[c2] != 0u) {
                    if (k == 1u) inside(c2); else part = true;
                }
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, part);
        if (part) {
            const int pos = nq + __popc(m & ((1u << lane) - 1u));
            if (pos < KL_QCAP) queue[pos] = c2;
        }
        nq += __popc(m);
    }
    __syncwarp();
    return nq < KL_QCAP ? nq : KL_QCAP;
}

// one warp per compacted point of an AMBIG cell: count its neighbours exactly
__global__ void __launch_bounds__(256) kl_exact_core_kernel(KlGrid g, KlPts pts, unsigned total)
{
    __shared__ unsigned queue[8][KL_QCAP];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned nw = gridDim.x * 8;
    for (unsigned i = blockIdx.x * 8 + wib; i < total; i += nw) {
        const unsigned c = pts.cell[i];
        const unsigned char st = g.status[c];
        if (st != KL_AMBIG) {
            if (lane == 0) { pts.flag[i] = (st == KL_ALLCORE) ? 1 : 0; pts.label[i] = KL_NOCELL; }
            continue;
        }
        const unsigned tile = c / KL_TC;
        const int lx = (int)(c % KL_TC) & 15, ly = (int)(c % KL_TC) >> 4;
        const double x = pts.x[i], y = pts.y[i];
        kl_u64 acc = 0;
        const int nq = kl_scan_window(g, tile, lx, ly, lane, queue[wib], [&](unsigned c2) { acc += g.cnt[c2]; });
        for (int q = 0; q < nq; ++q) {
            const unsigned c2 = queue[wib][q];
            const unsigned o = g.off[c2], n = g.cnt[c2];
            if (n > 4096u && !kl_charge(g, n, lane)) break;            // (small lists are not worth an atomic)
            for (unsigned j = lane; j < n; j += 32) acc += kl_near(x, y, pts.x[o + j], pts.y[o + j], g.eps2) ? 1ull : 0ull;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        const bool core = (long long)acc >= g.min_samples;
        if (lane == 0) {
            pts.flag[i] = core ? 1 : 0;
            pts.label[i] = KL_NOCELL;
            if (core) { atomicAdd(&g.ncore[c], 1u); atomicMin(&g.mincore[c], pts.idx[i]); }
        }
        __syncwarp();
    }
}

// one warp per cell with core points whose links are not all settled on cells: AMBIG cells, and core cells
// that take part in a partial pair
__global__ void __launch_bounds__(256) kl_union_exact_kernel(KlGrid g, KlPts pts)
{
    __shared__ unsigned queue[8][KL_QCAP];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (unsigned tile = blockIdx.x; tile <= g.hmask; tile += gridDim.x) {
      if (g.hkeys[tile] == KL_KEY_EMPTY) continue;
      for (unsigned lc = wib; lc < KL_TC; lc += 8) {
        const unsigned c = tile * KL_TC + lc;
        const unsigned char st = g.status[c];
        if (st == KL_EMPTY || g.ncore[c] == 0u) continue;
        if (!(st == KL_AMBIG || g.inv[c])) continue;
        const int lx = (int)(c % KL_TC) & 15, ly = (int)(c % KL_TC) >> 4;
        const int nq = kl_scan_window(g, tile, lx, ly, lane, queue[wib], [&](unsigned c2) {
            if (g.ncore[c2] != 0u && (st == KL_AMBIG || g.status[c2] == KL_AMBIG)) kl_union(g.parent, c, c2);
        });
        const unsigned oa = g.off[c], na = g.cnt[c];
        for (int q = 0; q < nq; ++q) {
            const unsigned c2 = queue[wib][q];
            if (c2 <= c || g.ncore[c2] == 0u || !g.inv[c2]) continue;          // warp-uniform
            if (kl_find(g.parent, c) == kl_find(g.parent, c2)) continue;
            const unsigned ob = g.off[c2], nb2 = g.cnt[c2];
            bool hit = false;
            // Along the direction u from this cell to the other one a pair can only be within eps if its projections
            // are: with amax = the farthest core point of A along u and bmin = the nearest of B, only a.u >= bmin - eps
            // and b.u <= amax + eps can take part (and nothing can if bmin - amax > eps).  Two dense cells a little more
            // than eps apart -- the expensive case -- are settled by this in one pass over their points.
            double ux, uy;
            {
                const double cax = pts.x[oa], cay = pts.y[oa], cbx = pts.x[ob], cby = pts.y[ob];
                ux = cbx - cax; uy = cby - cay;
                const double nrm = sqrt(ux * ux + uy * uy);
                if (nrm > 0.0) { ux /= nrm; uy /= nrm; } else { ux = 1.0; uy = 0.0; }
            }
            double amax = -1.0e300, bmin = 1.0e300;
            for (unsigned a = lane; a < na; a += 32) if (pts.flag[oa + a]) amax = fmax(amax, pts.x[oa + a] * ux + pts.y[oa + a] * uy);
            for (unsigned b = lane; b < nb2; b += 32) if (pts.flag[ob + b] == 1) bmin = fmin(bmin, pts.x[ob + b] * ux + pts.y[ob + b] * uy);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
                bmin = fmin(bmin, __shfl_xor_sync(0xffffffffu, bmin, o));
            }
            const double eps = sqrt(g.eps2), slack = 1.0e-9 * (fabs(amax) + fabs(bmin) + eps);   // projections are rounded
            if (bmin - amax > eps + slack) continue;                            // warp-uniform: no pair can be in range
            const double alo = bmin - eps - slack, bhi = amax + eps + slack;
            unsigned long long nb_in = 0, na_in = 0;
            for (unsigned b = lane; b < nb2; b += 32) nb_in += (pts.flag[ob + b] == 1 && pts.x[ob + b] * ux + pts.y[ob + b] * uy <= bhi) ? 1u : 0u;
            for (unsigned a = lane; a < na; a += 32) na_in += (pts.flag[oa + a] && pts.x[oa + a] * ux + pts.y[oa + a] * uy >= alo) ? 1u : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { nb_in += __shfl_xor_sync(0xffffffffu, nb_in, o); na_in += __shfl_xor_sync(0xffffffffu, na_in, o); }
            if (na_in * nb_in > 65536ull && !kl_charge(g, na_in * nb_in, lane)) continue;
            for (unsigned a = 0; a < na && !hit; ++a) {
                if (!pts.flag[oa + a]) continue;                               // core points only
                const double ax = pts.x[oa + a], ay = pts.y[oa + a];
                if (ax * ux + ay * uy < alo) continue;                         // warp-uniform
                bool h = false;
                for (unsigned b = lane; b < nb2; b += 32) {
                    const double bx = pts.x[ob + b], by = pts.y[ob + b];
                    h |= (pts.flag[ob + b] == 1) && (bx * ux + by * uy <= bhi) && kl_near(ax, ay, bx, by, g.eps2);
                }
                hit = __any_sync(0xffffffffu, h);
            }
            if (hit && lane == 0) kl_union(g.parent, c, c2);
            __syncwarp();
        }
      }
    }
}

__global__ void __launch_bounds__(KL_TC) kl_rootmin_kernel(KlGrid g)
{
    const unsigned tile = blockIdx.x;
    if (g.hkeys[tile] == KL_KEY_EMPTY) return;
    const unsigned c = tile * KL_TC + threadIdx.x;
    if (g.status[c] != KL_EMPTY && g.ncore[c] != 0u) atomicMin(&g.rootmin[g.parent[c]], g.mincore[c]);
}

// one warp per compacted non-core point: lowest-numbered cluster among its core neighbours
__global__ void __launch_bounds__(256) kl_border_kernel(KlGrid g, KlPts pts, unsigned total)
{
    __shared__ unsigned queue[8][KL_QCAP];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned nw = gridDim.x * 8;
    for (unsigned i = blockIdx.x * 8 + wib; i < total; i += nw) {
        if (pts.flag[i] != 0) continue;
        const unsigned c = pts.cell[i];
        const unsigned tile = c / KL_TC;
        const int lx = (int)(c % KL_TC) & 15, ly = (int)(c % KL_TC) >> 4;
        const double x = pts.x[i], y = pts.y[i];
        kl_u64 best = KL_NOIDX;
        unsigned broot = KL_NOCELL;
        const int nq = kl_scan_window(g, tile, lx, ly, lane, queue[wib], [&](unsigned c2) {
            if (g.ncore[c2] != 0u) {
                const unsigned r = g.parent[c2];
                const kl_u64 m = g.rootmin[r];
                if (m < best) { best = m; broot = r; }
            }
        });
        for (int q = 0; q < nq; ++q) {
            const unsigned c2 = queue[wib][q];
            if (g.ncore[c2] == 0u) continue;
            const unsigned r = g.parent[c2];
            const kl_u64 m = g.rootmin[r];
            if (__all_sync(0xffffffffu, m >= best)) continue;                  // cannot improve any lane
            const unsigned o = g.off[c2], n = g.cnt[c2];
            if (n > 4096u && !kl_charge(g, n, lane)) break;
            bool h = false;
            for (unsigned j = lane; j < n; j += 32)
                h |= (pts.flag[o + j] == 1) && kl_near(x, y, pts.x[o + j], pts.y[o + j], g.eps2);
            if (__any_sync(0xffffffffu, h) && m < best) { best = m; broot = r; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const kl_u64 b2 = __shfl_xor_sync(0xffffffffu, best, o);
            const unsigned r2 = __shfl_xor_sync(0xffffffffu, broot, o);
            if (b2 < best) { best = b2; broot = r2; }
        }
        if (lane == 0 && broot != KL_NOCELL) { pts.flag[i] = 2; pts.label[i] = broot; }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ cluster sums
__global__ void __launch_bounds__(KL_TC) kl_cluster_ids_kernel(KlGrid g, KlAcc acc)
{
    const unsigned tile = blockIdx.x;
    if (g.hkeys[tile] == KL_KEY_EMPTY) return;
    const unsigned c = tile * KL_TC + threadIdx.x;
    if (g.status[c] == KL_EMPTY || g.ncore[c] == 0u || g.parent[c] != c) return;
    const unsigned id = atomicAdd(acc.count, 1u);
    if (id >= acc.cap) { atomicOr(g.err, KL_ERR_CLUSTERS); return; }
    g.cid[c] = id;
    acc.minidx[id] = g.rootmin[c];
}

__device__ __forceinline__ void kl_add128(kl_u64 *lo, long long *hi, kl_u64 v)
{
    const kl_u64 old = atomicAdd(lo, v);
    if (old + v < old) atomicAdd(reinterpret_cast<kl_u64 *>(hi), 1ull);
}

__device__ __forceinline__ void kl_acc_add(const KlAcc &acc, unsigned id, kl_u64 n, long long gx, long long gy, kl_u64 qx, kl_u64 qy)
{
    atomicAdd(&acc.n[id], n);
    atomicAdd(reinterpret_cast<kl_u64 *>(&acc.ax[id]), (kl_u64)(gx * (long long)n));
    atomicAdd(reinterpret_cast<kl_u64 *>(&acc.ay[id]), (kl_u64)(gy * (long long)n));
    kl_add128(&acc.bxl[id], &acc.bxh[id], qx);
    kl_add128(&acc.byl[id], &acc.byh[id], qy);
}

// core cells enter with their sums ...
__global__ void __launch_bounds__(KL_TC) kl_acc_cells_kernel(KlGrid g, KlAcc acc)
{
    const unsigned tile = blockIdx.x;
    const kl_u64 key = g.hkeys[tile];
    if (key == KL_KEY_EMPTY) return;
    const unsigned c = tile * KL_TC + threadIdx.x;
    if (g.status[c] != KL_ALLCORE) return;
    const unsigned id = g.cid[g.parent[c]];
    if (id == KL_NOCELL) return;
    int tx, ty;
    kl_key_decode(key, tx, ty);
    kl_acc_add(acc, id, g.cnt[c], (long long)tx * KL_TS + (threadIdx.x & 15), (long long)ty * KL_TS + (threadIdx.x >> 4), g.sx[c], g.sy[c]);
}

// ... the points of the other cells one by one
__global__ void __launch_bounds__(256) kl_acc_points_kernel(KlGrid g, KlPts pts, KlAcc acc, unsigned total)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const unsigned c = pts.cell[i];
    const unsigned char f = pts.flag[i];
    if (f == 0 || g.status[c] == KL_ALLCORE) return;
    const unsigned root = (f == 1) ? g.parent[c] : pts.label[i];
    const unsigned id = g.cid[root];
    if (id == KL_NOCELL) return;
    int tx, ty, lc; kl_u64 qx, qy;
    if (!kl_locate(g, pts.x[i], pts.y[i], tx, ty, lc, qx, qy)) return;
    kl_acc_add(acc, id, 1ull, (long long)tx * KL_TS + (lc & 15), (long long)ty * KL_TS + (lc >> 4), qx, qy);
}

// ------------------------------------------------------------------------------------------------ shards
// A filter sharded over several GPUs clusters as one: every shard counts its own maps, the occupied tiles travel
// as records (key, 256 counts, lowest indices, offset sums), every shard merges all records into the same grid and
// runs the cell level on it; the points of involved cells are extracted per shard, gathered, and compacted from
// the gathered list, after which every shard holds the identical result.
#define KL_REC_WORDS (1 + KL_TC / 2 + 3 * KL_TC)      // 897 x 8 bytes

__global__ void __launch_bounds__(KL_TC) kl_export_tiles_kernel(KlGrid g, kl_u64 *out, unsigned cap, unsigned *n)
{
    __shared__ unsigned s_r;
    const unsigned tile = blockIdx.x;
    const kl_u64 key = g.hkeys[tile];
    if (key == KL_KEY_EMPTY) return;
    if (threadIdx.x == 0) s_r = atomicAdd(n, 1u);
    __syncthreads();
    if (s_r >= cap) return;
    kl_u64 *rec = out + (size_t)s_r * KL_REC_WORDS;
    const unsigned c = tile * KL_TC + threadIdx.x;
    if (threadIdx.x == 0) rec[0] = key;
    reinterpret_cast<unsigned *>(rec + 1)[threadIdx.x] = g.cnt[c];
    rec[1 + KL_TC / 2 + threadIdx.x] = g.minidx[c];
    rec[1 + KL_TC / 2 + KL_TC + threadIdx.x] = g.sx[c];
    rec[1 + KL_TC / 2 + 2 * KL_TC + threadIdx.x] = g.sy[c];
}

__global__ void __launch_bounds__(KL_TC) kl_merge_tiles_kernel(KlGrid g, const kl_u64 *recs, unsigned n)
{
    __shared__ unsigned s_tile;
    const kl_u64 *rec = recs + (size_t)blockIdx.x * KL_REC_WORDS;
    if (threadIdx.x == 0) {
        int tx, ty;
        kl_key_decode(rec[0], tx, ty);
        s_tile = kl_tile_insert(g, tx, ty);
    }
    __syncthreads();
    if (s_tile == KL_NOCELL) return;
    const unsigned cnt = reinterpret_cast<const unsigned *>(rec + 1)[threadIdx.x];
    if (!cnt) return;
    const unsigned c = s_tile * KL_TC + threadIdx.x;
    atomicAdd(&g.cnt[c], cnt);
    atomicMin(&g.minidx[c], rec[1 + KL_TC / 2 + threadIdx.x]);
    atomicAdd(&g.sx[c], rec[1 + KL_TC / 2 + KL_TC + threadIdx.x]);
    atomicAdd(&g.sy[c], rec[1 + KL_TC / 2 + 2 * KL_TC + threadIdx.x]);
}

struct KlExtractOp {        // (x, y, index) of every local point that sits in an involved cell
    KlGrid g;
    double *out;
    unsigned cap;
    unsigned *n;
    __device__ void operator()(double x, double y, kl_u64 idx) const
    {
        int tx, ty, lc; kl_u64 qx, qy;
        if (!kl_locate(g, x, y, tx, ty, lc, qx, qy)) return;
        const unsigned t = kl_tile_find(g, tx, ty);
        if (t == KL_NOCELL || !g.inv[t * KL_TC + (unsigned)lc]) return;
        const unsigned pos = atomicAdd(n, 1u);
        if (pos >= cap) return;
        out[3 * (size_t)pos] = x; out[3 * (size_t)pos + 1] = y; out[3 * (size_t)pos + 2] = (double)idx;   // idx < 2^53
    }
};

template <class Op>
__global__ void __launch_bounds__(256) kl_pass_flat3(const double *p, long long n, Op op)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        op(p[3 * i], p[3 * i + 1], (kl_u64)p[3 * i + 2]);
}
