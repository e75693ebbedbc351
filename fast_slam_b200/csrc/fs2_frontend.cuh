// fs2_frontend.cuh -- the scan front-end on the device, batched over scans (rows A11-A15 of SURVEY.md 8a):
//   LineFilter.filter                       line_filter.py:12-21            -> fe_filter_geometry
//   HoughTransformation image + HoughLines  hough_transformation.py:44-73,24 -> fe_raster_list, fe_vote_peaks, fe_peaks_rank
//                                           (global-accumulator form, FS2_FE_LEGACY=1: fe_raster_vote, fe_peaks_find)
//   line intersections, back to metres      hough_transformation.py:76-145  -> fe_intersect_cluster
//   GeometryUtils.cluster_points (DBSCAN eps 0.5, min_samples 1 = connected components)  geometry_utils.py:26-62
//   LandmarkUtils.__get_corners             landmark_utils.py:66-89
//   GeometryUtils.calculate_distance_and_angle  geometry_utils.py:65-74
// The integer parts follow OpenCV's HoughLinesStandard exactly (float32 trig tables accumulated in float32,
// round-half-even of the float32 sum, 4-neighbour maxima above the threshold, votes-descending / index-
// ascending order), so lines are bit-identical to cv2's; intersections use cosf/sinf where the reference goes
// through numpy's float32 cos/sin, hence real-valued results agree to float32 rounding.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FE_SCALE 100.0
#define FE_PAD 20
#define FE_NUMANGLE 180
#define FE_MAX_LINES 128
#define FE_MAX_INTER 2048
#define FE_MAX_K 64
#define FE_THREADS 256
#define FE_ST_LINES_OVERFLOW 1
#define FE_ST_INTER_OVERFLOW 2
#define FE_ST_K_OVERFLOW 4
#define FE_ST_EMPTY 8            // no beam inside the range limits (the reference raises on an empty scan)
#define FE_ST_NONFINITE 16       // a beam or point that is not finite

__constant__ float fe_tab_sin[FE_NUMANGLE];
__constant__ float fe_tab_cos[FE_NUMANGLE];
__constant__ double fe_kernel[65];   // gaussian weights, radius <= 32

struct FeGeo {             // per scan
    int off_x, off_y, width, height;
    int numrho;
    int npts;              // points of this scan (<= N: beams outside the range limits are dropped)
    long long bitmap_off;  // in 32-bit words
    long long acc_off;     // in ints
};

// ---- Robot.scan_environment (models/robot.py:32-58), batched: beams outside [min_range, max_range] are dropped
// (order kept), the rest become x = dist * cos(angle), y = dist * sin(angle); the cos/sin of the beam angles
// come from the host's libm, as the reference's math.cos / math.sin do.
__global__ void __launch_bounds__(FE_THREADS)
fe_polar_points(const double *ranges, const double *tab_cos, const double *tab_sin, int N, double min_range,
                double max_range, double *scans, int *nvalid, int *status)
{
    __shared__ int s_w[FE_THREADS / 32];
    __shared__ int s_base;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double *r = ranges + (size_t)b * N;
    double *dst = scans + (size_t)b * N * 2;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < N; i0 += FE_THREADS) {
        const int i = i0 + tid;
        double d = 0.0;
        bool ok = false;
        if (i < N) {
            d = r[i];
            ok = !(d < min_range || d > max_range);        // robot.py:48 (a NaN passes, as it does there)
            if (ok && !isfinite(d)) { atomicOr(&status[b], FE_ST_NONFINITE); ok = false; }
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_w[wid] = __popc(m);
        __syncthreads();
        int before = s_base + __popc(m & ((1u << lane) - 1u));
        for (int k = 0; k < wid; ++k) before += s_w[k];
        if (ok) {
            dst[2 * before] = __dmul_rn(d, tab_cos[i]);    // robot.py:55-56
            dst[2 * before + 1] = __dmul_rn(d, tab_sin[i]);
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int k = 0; k < FE_THREADS / 32; ++k) t += s_w[k]; s_base += t; }
        __syncthreads();
    }
    if (tid == 0) { nvalid[b] = s_base; if (s_base == 0) atomicOr(&status[b], FE_ST_EMPTY); }
}

// ---- A11 + image geometry ---------------------------------------------------------------------------
__global__ void __launch_bounds__(FE_THREADS)
fe_filter_geometry(const double *scans, int N, const int *nvalid, int radius, double *filtered, FeGeo *geo)
{
    __shared__ double smin[2][FE_THREADS / 32], smax[2][FE_THREADS / 32];
    const int b = blockIdx.x;
    const double *src = scans + (size_t)b * N * 2;
    double *dst = filtered + (size_t)b * N * 2;
    double mn[2] = {1e300, 1e300}, mx[2] = {-1e300, -1e300};
    const int stride = N;
    (void)stride;
    N = nvalid ? nvalid[b] : N;                              // the filter's "reflect" boundary is the scan's own end
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        for (int c = 0; c < 2; ++c) {
            double acc;
            if (radius == 0) {
                acc = src[2 * i + c];                 // identity at the default sigma (quirk Q17)
            } else {
                acc = 0.0;
                for (int k = -radius; k <= radius; ++k) {
                    int j = i + k;
                    while (j < 0 || j >= N) j = (j < 0) ? -j - 1 : 2 * N - 1 - j;   // scipy 'reflect'
                    acc = __dadd_rn(acc, __dmul_rn(fe_kernel[k + radius], src[2 * j + c]));
                }
            }
            dst[2 * i + c] = acc;
            const double s = __dmul_rn(acc, FE_SCALE);
            mn[c] = fmin(mn[c], s);
            mx[c] = fmax(mx[c], s);
        }
    }
    for (int c = 0; c < 2; ++c) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fmin(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = fmax(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
        if ((threadIdx.x & 31) == 0) { smin[c][threadIdx.x >> 5] = mn[c]; smax[c][threadIdx.x >> 5] = mx[c]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a[2] = {1e300, 1e300}, z[2] = {-1e300, -1e300};
        for (int c = 0; c < 2; ++c)
            for (int k = 0; k < FE_THREADS / 32; ++k) { a[c] = fmin(a[c], smin[c][k]); z[c] = fmax(z[c], smax[c][k]); }
        const int min_x = (int)a[0], min_y = (int)a[1], max_x = (int)z[0], max_y = (int)z[1];   // int(): toward zero
        FeGeo g;
        g.off_x = (min_x < 0 ? -min_x : 0) + FE_PAD;
        g.off_y = (min_y < 0 ? -min_y : 0) + FE_PAD;
        g.width = max_x + g.off_x + FE_PAD;
        g.height = max_y + g.off_y + FE_PAD;
        g.numrho = 2 * (g.width + g.height) + 1;
        g.npts = N; g.bitmap_off = 0; g.acc_off = 0;
        if (N == 0) { g.off_x = g.off_y = FE_PAD; g.width = g.height = 2 * FE_PAD; g.numrho = 2 * (g.width + g.height) + 1; }
        geo[b] = g;
    }
}

// ---- A12: rasterise (13-pixel discs, de-duplicated through a bitmap) and vote ---------------------------
__global__ void __launch_bounds__(FE_THREADS)
fe_raster_vote(const double *filtered, int N, const FeGeo *geo, unsigned *bitmap, int *acc)
{
    const int b = blockIdx.y;
    const FeGeo g = geo[b];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= g.npts * 13) return;
    const int i = t / 13, d = t % 13;
    // the 13 pixels with |dx| + |dy| <= 2 (cv2.circle radius 2, filled)
    const int ddx[13] = {0, -1, 0, 1, -2, -1, 0, 1, 2, -1, 0, 1, 0};
    const int ddy[13] = {-2, -1, -1, -1, 0, 0, 0, 0, 0, 1, 1, 1, 2};
    const double *p = filtered + ((size_t)b * N + i) * 2;
    const int x = (int)__dmul_rn(p[0], FE_SCALE) + g.off_x + ddx[d];
    const int y = (int)__dmul_rn(p[1], FE_SCALE) + g.off_y + ddy[d];
    if (x < 0 || x >= g.width || y < 0 || y >= g.height) return;
    const long long pix = (long long)y * g.width + x;
    unsigned *word = bitmap + g.bitmap_off + (pix >> 5);
    const unsigned bit = 1u << (pix & 31);
    if (atomicOr(word, bit) & bit) return;            // another disc already set this pixel
    int *a = acc + g.acc_off;
    const float xf = (float)x, yf = (float)y;
    const int half = (g.numrho - 1) / 2;
    for (int n = 0; n < FE_NUMANGLE; ++n) {
        const int r = __float2int_rn(__fadd_rn(__fmul_rn(xf, fe_tab_cos[n]), __fmul_rn(yf, fe_tab_sin[n]))) + half;
        atomicAdd(a + (size_t)(n + 1) * (g.numrho + 2) + r + 1, 1);
    }
}

// ---- A12, fused form (the default): the accumulator never leaves the SM ------------------------------------
// fe_raster_vote + fe_peaks_find keep one (180 + 2) x (numrho + 2) int accumulator per scan in global memory -- 2.2 MB
// for a 1081-beam scan of a room, 2.2 GB for config 5's batch of 1024 scans: cleared, voted into with 1.6e9 global
// atomics and read back once, and sized on the host in the middle of the pipeline (a device -> host -> device round trip).
// Here a scan's set pixels are first written as a LIST (fe_raster_list: one block per scan, de-duplicated through a
// shared-memory hash set), then fe_vote_peaks gives every block a BAND of 10 angle rows (+ 1 halo row either side, the
// peak test looks at the rows n - 1 and n + 1) of one scan: it votes the scan's pixels into 16-bit cells in shared memory
// (a cell holds at most one vote per set pixel: <= 13 N < 65536) and runs OpenCV's local-maximum test right there.  Rows
// longer than the tile are done in several column ranges with a 1-cell halo.  Votes commute, so the lines are the same
// bit for bit as the global-accumulator path (FS2_FE_LEGACY=1, and any scan of more than FE_FUSED_MAX_POINTS points).
// the 13 pixels with |dx| + |dy| <= 2 (cv2.circle radius 2, filled)
__constant__ int fe_ddx[13] = {0, -1, 0, 1, -2, -1, 0, 1, 2, -1, 0, 1, 0};
__constant__ int fe_ddy[13] = {-2, -1, -1, -1, 0, 0, 0, 0, 0, 1, 1, 1, 2};
#define FE_HASH_BITS 15
#define FE_HASH_SIZE (1 << FE_HASH_BITS)
#define FE_LIST_THREADS 512
#define FE_FUSED_MAX_POINTS 2016            // 13 N <= 0.8 x FE_HASH_SIZE (and < 65536, the 16-bit vote cells)
#define FE_ST_TOO_LARGE 32                  // image of more than 2^28 pixels (the call fails with FS2_ERR_INVALID)

__global__ void __launch_bounds__(FE_LIST_THREADS)
fe_raster_list(const double *filtered, int N, const FeGeo *geo, int2 *pix, int2 *pix_t, int *npix, int *status)
{
    extern __shared__ int s_hash[];          // FE_HASH_SIZE pixel indices (y * width + x), -1 = empty
    __shared__ int s_n;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const FeGeo g = geo[b];
    if (g.width <= 0 || g.height <= 0 || (long long)g.width * g.height > (1ll << 28)) {
        if (tid == 0) { atomicOr(&status[b], FE_ST_TOO_LARGE); npix[b] = 0; }
        return;
    }
    for (int i = tid; i < FE_HASH_SIZE; i += FE_LIST_THREADS) s_hash[i] = -1;
    if (tid == 0) s_n = 0;
    __syncthreads();
    int2 *out = pix + (size_t)b * N * 13;
    const int total = g.npts * 13;
    for (int t0 = 0; t0 < total; t0 += FE_LIST_THREADS) {          // whole warps stay in the loop (ballot below)
        const int t = t0 + tid;
        bool fresh = false;
        int x = 0, y = 0;
        if (t < total) {
            const int i = t / 13, d = t - i * 13;
            const double *p = filtered + ((size_t)b * N + i) * 2;
            x = (int)__dmul_rn(p[0], FE_SCALE) + g.off_x + fe_ddx[d];
            y = (int)__dmul_rn(p[1], FE_SCALE) + g.off_y + fe_ddy[d];
            if (x >= 0 && x < g.width && y >= 0 && y < g.height) {
                const int key = y * g.width + x;
                unsigned h = ((unsigned)key * 2654435761u) >> (32 - FE_HASH_BITS);
                while (true) {                                     // linear probing; the table is never more than 0.8 full
                    const int old = atomicCAS(&s_hash[h], -1, key);
                    if (old == -1) { fresh = true; break; }
                    if (old == key) break;                         // another disc already set this pixel
                    h = (h + 1) & (FE_HASH_SIZE - 1);
                }
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, fresh);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_n, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (fresh) out[base + __popc(m & ((1u << lane) - 1u))] = make_int2(x, y);
        }
    }
    __syncthreads();
    const int np = s_n;
    if (tid == 0) npix[b] = np;
    // The list is in order of insertion: neighbours in it are neighbouring pixels of one disc or of consecutive beams, and
    // for most angles they vote for the same or the next rho cell -- 32 lanes of a voting warp on one shared-memory word
    // (measured: 8.4 bank-conflict replays per vote instruction).  The voters therefore read a TRANSPOSED copy: the list
    // cut into 32 segments, entry 32 j + l = element j of segment l, so that the lanes of a warp hold pixels a 32nd of the
    // scan apart (padded with x = -1 where the last segment is short).
    const int seg = (np + 31) >> 5;
    int2 *out_t = pix_t + (size_t)b * (N * 13 + 32);
    for (int q = tid; q < seg * 32; q += FE_LIST_THREADS) {
        const int src = (q & 31) * seg + (q >> 5);
        out_t[q] = src < np ? out[src] : make_int2(-1, -1);
    }
}

#define FE_VB_ROWS 10                                       // angle rows per band: 18 bands cover the 180 angles exactly
#define FE_VBANDS (FE_NUMANGLE / FE_VB_ROWS)
#define FE_VP_THREADS 512
#define FE_VP_ROWW 2304                                     // 32-bit words per tile row = 4608 cells, 2 of them halo
#define FE_VP_SMEM ((FE_VB_ROWS + 2) * FE_VP_ROWW * 4)      // 108 KB: two blocks per SM
static_assert(FE_VBANDS * FE_VB_ROWS == FE_NUMANGLE, "bands must tile the angles");

__device__ __forceinline__ int fe_cell(const unsigned *row, int c) { return (int)((row[c >> 1] >> ((c & 1) << 4)) & 0xffffu); }

__global__ void __launch_bounds__(FE_VP_THREADS, 2)
fe_vote_peaks(const FeGeo *geo, const int2 *pix, const int *npix, int N, int threshold, int2 *cand, int *ncand)
{
    extern __shared__ unsigned s_acc[];      // [FE_VB_ROWS + 2][FE_VP_ROWW] words, two 16-bit cells each
    const int b = blockIdx.y, tid = threadIdx.x;
    const FeGeo g = geo[b];
    const int np = npix[b];
    if (np == 0) return;                     // no set pixel, no line (also: images fe_raster_list refused)
    const int2 *px = pix + (size_t)b * (N * 13 + 32);        // the transposed list of fe_raster_list
    const int np_t = ((np + 31) >> 5) << 5;
    const int numrho = g.numrho, half = (numrho - 1) / 2, stride = numrho + 2;
    const int n0 = (int)blockIdx.x * FE_VB_ROWS - 1;         // angle of tile row 0 (-1: OpenCV's zero border row)
    const int W = 2 * FE_VP_ROWW - 2;                        // interior cells per column range
    // The kernel is bound by instruction issue (ncu: 80 % issue-active, 29 instructions per warp-wide vote before this
    // form), so the vote is kept to its arithmetic: the band's trig values sit in registers; rows outside 0 .. 179 (the
    // first and the last band's halo) vote with cos = sin = 0 and are cleared again before the peak test; a scan whose
    // rho axis fits one column range -- the usual case -- needs no range test (every rho index is inside by construction).
    float cs[FE_VB_ROWS + 2], sn[FE_VB_ROWS + 2];
#pragma unroll
    for (int k = 0; k < FE_VB_ROWS + 2; ++k) {
        const int n = n0 + k;
        const bool valid = n >= 0 && n < FE_NUMANGLE;
        cs[k] = valid ? fe_tab_cos[valid ? n : 0] : 0.f;
        sn[k] = valid ? fe_tab_sin[valid ? n : 0] : 0.f;
    }
    const bool lo_border = n0 < 0, hi_border = n0 + FE_VB_ROWS + 1 >= FE_NUMANGLE;
    for (int r0 = 0; r0 < numrho; r0 += W) {
        const int wt = min(W, numrho - r0);                  // this range holds rho indices [r0, r0 + wt) in cells 1 .. wt
        const int words = (wt + 3) >> 1;                     // cells 0 .. wt + 1
        for (int k = 0; k < FE_VB_ROWS + 2; ++k)
            for (int i = tid; i < words; i += FE_VP_THREADS) s_acc[k * FE_VP_ROWW + i] = 0u;
        __syncthreads();
        const int off = half - r0 + 1;
        if (numrho <= W) {
            for (int p = tid; p < np_t; p += FE_VP_THREADS) {
                const int2 q = px[p];
                if (q.x < 0) continue;                       // padding
                const float xf = (float)q.x, yf = (float)q.y;
#pragma unroll
                for (int k = 0; k < FE_VB_ROWS + 2; ++k) {
                    const int c = __float2int_rn(__fadd_rn(__fmul_rn(xf, cs[k]), __fmul_rn(yf, sn[k]))) + off;
                    atomicAdd(&s_acc[k * FE_VP_ROWW + (c >> 1)], 1u + (unsigned)(c & 1) * 0xffffu);
                }
            }
        } else {
            for (int p = tid; p < np_t; p += FE_VP_THREADS) {
                const int2 q = px[p];
                if (q.x < 0) continue;
                const float xf = (float)q.x, yf = (float)q.y;
#pragma unroll
                for (int k = 0; k < FE_VB_ROWS + 2; ++k) {
                    const int c = __float2int_rn(__fadd_rn(__fmul_rn(xf, cs[k]), __fmul_rn(yf, sn[k]))) + off;
                    if ((unsigned)c <= (unsigned)(wt + 1)) atomicAdd(&s_acc[k * FE_VP_ROWW + (c >> 1)], 1u + (unsigned)(c & 1) * 0xffffu);
                }
            }
        }
        __syncthreads();
        if (lo_border || hi_border) {                        // the same for the whole block
            if (lo_border) for (int i = tid; i < words; i += FE_VP_THREADS) s_acc[i] = 0u;
            if (hi_border) for (int i = tid; i < words; i += FE_VP_THREADS) s_acc[(FE_VB_ROWS + 1) * FE_VP_ROWW + i] = 0u;
            __syncthreads();
        }
        for (int k = 1; k <= FE_VB_ROWS; ++k) {
            const unsigned *row = s_acc + k * FE_VP_ROWW;
            for (int i = tid; i < words; i += FE_VP_THREADS) {       // two cells per word; nearly all are below the threshold
                const unsigned w2 = row[i];
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int v = hf ? (int)(w2 >> 16) : (int)(w2 & 0xffffu);
                    const int c = 2 * i + hf;
                    if (v <= threshold || c < 1 || c > wt) continue;
                    if (v > fe_cell(row, c - 1) && v >= fe_cell(row, c + 1) && v > fe_cell(row - FE_VP_ROWW, c) &&
                        v >= fe_cell(row + FE_VP_ROWW, c)) {
                        const int base = (n0 + k + 1) * stride + (r0 + c - 1) + 1;      // index in OpenCV's padded accumulator
                        const int slot = atomicAdd(&ncand[b], 1);
                        if (slot < FE_MAX_LINES) cand[(size_t)b * FE_MAX_LINES + slot] = make_int2(base, v);
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ---- A12: local maxima above the threshold, strongest first ----------------------------------------------
// Two kernels: FE_PSPLIT blocks per scan each search a band of angle rows of the accumulator (one block per scan left
// 90 % of the GPU idle for a single scan and was latency-bound for a batch) and append what they find to the scan's
// candidate list; one block per scan then ranks the candidates.  The order of discovery does not matter: the rank
// (votes descending, accumulator index ascending -- OpenCV's hough_cmp_gt) is a total order.
#define FE_PSPLIT 12
__global__ void __launch_bounds__(FE_THREADS)
fe_peaks_find(const FeGeo *geo, const int *acc, int threshold, int2 *cand, int *ncand)
{
    const int b = blockIdx.y;
    const FeGeo g = geo[b];
    const int *a = acc + g.acc_off;
    const int stride = g.numrho + 2;
    const int rows = (FE_NUMANGLE + FE_PSPLIT - 1) / FE_PSPLIT;
    const int n0 = blockIdx.x * rows, n1 = min(FE_NUMANGLE, n0 + rows);
    const long long cells = (long long)(n1 - n0) * g.numrho;
    for (long long c = threadIdx.x; c < cells; c += blockDim.x) {
        const int n = n0 + (int)(c / g.numrho), r = (int)(c % g.numrho);
        const int base = (n + 1) * stride + r + 1;
        const int v = a[base];
        if (v > threshold && v > a[base - 1] && v >= a[base + 1] && v > a[base - stride] && v >= a[base + stride]) {
            const int k = atomicAdd(&ncand[b], 1);
            if (k < FE_MAX_LINES) cand[(size_t)b * FE_MAX_LINES + k] = make_int2(base, v);
        }
    }
}

__global__ void __launch_bounds__(FE_MAX_LINES)
fe_peaks_rank(const FeGeo *geo, const int2 *cand, const int *ncand, float2 *lines, int *nlines, int *status)
{
    __shared__ int s_base[FE_MAX_LINES], s_votes[FE_MAX_LINES];
    const int b = blockIdx.x;
    const FeGeo g = geo[b];
    const int stride = g.numrho + 2;
    int n = ncand[b];
    if (n > FE_MAX_LINES) { n = FE_MAX_LINES; if (threadIdx.x == 0) atomicOr(&status[b], FE_ST_LINES_OVERFLOW); }
    const int k = threadIdx.x;
    if (k < n) { const int2 c = cand[(size_t)b * FE_MAX_LINES + k]; s_base[k] = c.x; s_votes[k] = c.y; }
    __syncthreads();
    if (k < n) {
        // rank sort: votes descending, accumulator index ascending (hough_cmp_gt)
        int rank = 0;
        for (int j = 0; j < n; ++j)
            rank += (s_votes[j] > s_votes[k]) || (s_votes[j] == s_votes[k] && s_base[j] < s_base[k]);
        const int base = s_base[k];
        const int an = base / stride - 1, r = base - (an + 1) * stride - 1;
        const float thetaf = (float)(3.14159265358979323846 / 180.0);
        float2 l;
        l.x = __fmul_rn(__fadd_rn((float)r, -__fmul_rn((float)(g.numrho - 1), 0.5f)), 1.0f);
        l.y = __fmul_rn((float)an, thetaf);
        lines[(size_t)b * FE_MAX_LINES + rank] = l;
    }
    if (threadIdx.x == 0) nlines[b] = n;
}

// union-find over indices in shared memory: lab[i] <= i always, roots have lab[i] == i
__device__ __forceinline__ int fe_find(volatile int *lab, int i)
{
    int p;
    while ((p = lab[i]) != i) i = p;
    return i;
}

// link the sets of two root-level entries a > b: the larger root is hooked under the smaller one; if `a` had been hooked
// meanwhile (to `old`), its entry keeps the smaller of the two parents and the other one is linked to it in turn, so no
// link is ever lost
__device__ __forceinline__ void fe_merge_roots(int *lab, int a, int b)
{
    while (a != b) {
        if (a < b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(&lab[a], b);
        if (old == a) return;
        a = old;
    }
}

// ---- A12 (intersections) + A13 + A14 + A15 -------------------------------------------------------------
// 1024 threads per scan: the pair loops below are chains of dependent fp64 instructions, and the slowest scan of a batch
// (the one with the most intersections) sets the time of the launch -- a block needs its warps to hide that latency
#define FE_IC_THREADS 1024
__global__ void __launch_bounds__(FE_IC_THREADS, 2)
fe_intersect_cluster(const double *filtered, int N, const FeGeo *geo, const float2 *lines, const int *nlines,
                     double eps_sq, double corner_sq, double *meas, int *kcount, int *status, float2 *inter_out, int *ninter)
{
    __shared__ float s_rho[FE_MAX_LINES], s_th[FE_MAX_LINES], s_cos[FE_MAX_LINES], s_sin[FE_MAX_LINES];
    __shared__ float2 s_pt[FE_MAX_INTER];
    __shared__ int s_lab[FE_MAX_INTER];
    __shared__ int s_scan[FE_IC_THREADS / 32];
    __shared__ int s_cnt, s_k;
    __shared__ float2 s_cent[FE_MAX_K];
    __shared__ int s_keep[FE_MAX_K];
    const int b = blockIdx.x, tid = threadIdx.x;
    const FeGeo g = geo[b];
    const int n = nlines[b];
    for (int k = tid; k < n; k += blockDim.x) {
        const float2 l = lines[(size_t)b * FE_MAX_LINES + k];
        s_rho[k] = l.x; s_th[k] = l.y; s_cos[k] = cosf(l.y); s_sin[k] = sinf(l.y);
    }
    if (tid == 0) { s_cnt = 0; s_k = 0; }
    __syncthreads();
    // pairs in the reference's nested-loop order (i outer, j inner), compacted in that order
    const int npairs = n * (n - 1) / 2;
    const float lim = 0.78539816339744830962f;   // np.deg2rad(45) compared against a float32 difference
    for (int base = 0; base < npairs; base += blockDim.x) {
        const int t = base + tid;
        bool ok = false;
        float x = 0.f, y = 0.f;
        if (t < npairs) {
            // invert t = i*n - i*(i+1)/2 + (j - i - 1)
            int i = (int)floorf(((2.f * n - 1.f) - sqrtf((2.f * n - 1.f) * (2.f * n - 1.f) - 8.f * t)) * 0.5f);
            while (i > 0 && (long long)i * n - (long long)i * (i + 1) / 2 > t) --i;
            while ((long long)(i + 1) * n - (long long)(i + 1) * (i + 2) / 2 <= t) ++i;
            const int j = t - (i * n - i * (i + 1) / 2) + i + 1;
            float d = fabsf(__fadd_rn(s_th[i], -s_th[j]));
            d = fminf(d, __fadd_rn(3.14159265358979323846f, -d));
            if (!((double)d < 0.78539816339744830962)) {
                const float a1 = s_cos[i], b1 = s_sin[i], a2 = s_cos[j], b2 = s_sin[j];
                const float det = __fadd_rn(__fmul_rn(a1, b2), -__fmul_rn(a2, b1));
                if (fabsf(det) > 1e-10f) {
                    x = __fdiv_rn(__fadd_rn(__fmul_rn(b2, s_rho[i]), -__fmul_rn(b1, s_rho[j])), det);
                    y = __fdiv_rn(__fadd_rn(__fmul_rn(a1, s_rho[j]), -__fmul_rn(a2, s_rho[i])), det);
                    ok = (x >= 0.f && x < (float)g.width && y >= 0.f && y < (float)g.height);
                }
            }
        }
        (void)lim;
        // ordered compaction: ballot inside the warp, warp totals through shared memory
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        const int lanep = __popc(bal & ((1u << (tid & 31)) - 1u));
        if ((tid & 31) == 0) s_scan[tid >> 5] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < FE_IC_THREADS / 32; ++w) { const int c = s_scan[w]; if (w < (tid >> 5)) woff += c; tot += c; }
        const int pos = s_cnt + woff + lanep;
        if (ok && pos < FE_MAX_INTER) {
            // back to metres in float32 (hough_transformation.py:142-145)
            s_pt[pos] = make_float2(__fdiv_rn(__fadd_rn(x, -(float)g.off_x), 100.0f), __fdiv_rn(__fadd_rn(y, -(float)g.off_y), 100.0f));
        }
        __syncthreads();
        if (tid == 0) s_cnt += tot;
        __syncthreads();
    }
    int C = s_cnt;
    if (C > FE_MAX_INTER) { C = FE_MAX_INTER; if (tid == 0) atomicOr(&status[b], FE_ST_INTER_OVERFLOW); }
    if (inter_out) {       // HoughTransformation.detect_line_intersections returns these (hough_transformation.py:14-41)
        for (int i = tid; i < C; i += blockDim.x) inter_out[(size_t)b * FE_MAX_INTER + i] = s_pt[i];
        if (tid == 0) ninter[b] = C;
    }
    // connected components at eps (DBSCAN, min_samples = 1), labelled by their minimum index.  "distance <= eps" is tested
    // as "squared distance <= eps_sq", where the host passes the largest double whose correctly rounded square root is
    // <= eps (fe_sq_threshold): the same decision as np.sqrt(dx**2 + dy**2) <= eps for every input, without the root.
    // The intersections sit in a few clusters of hundreds of points around the corners, nearly all pairs of a cluster
    // within eps: merging per pair is what costs (measured: 85 % of the kernel), not the distance tests.  So:
    //  (1) every point takes its FIRST neighbour j < i as parent (no atomics, the loop stops at the hit): trees whose
    //      root is their smallest index, one or very few per cluster; flatten;
    //  (2) one sweep over all pairs j < i with the labels frozen: a pair within eps whose labels differ merges the two
    //      ROOTS (lock-free, only root entries are written; a read of the larger root's entry skips what is merged
    //      already).  The merges see every edge of the eps-graph, so one sweep completes the components; flatten.
    // Both loops run warp-uniform (every lane of a warp walks j up to the warp's largest i, lanes past their own i idle),
    // four pairs per iteration so that four chains of dependent fp64 instructions overlap, and reconverge explicitly:
    // left to themselves the lanes of a warp drifted apart after the first divergent merge and ran the pair loop 6 lanes
    // at a time (ncu source view: 22 M warp iterations instead of 3.4 M).
    volatile int *lab = s_lab;
    const int lane = tid & 31;
    for (int i0 = tid & ~31; i0 < C; i0 += blockDim.x) {
        const int i = i0 + lane;
        const bool act = i < C;
        const int jend = min(i0 + 31, C - 1);                   // the warp's largest i
        const double xi = act ? (double)s_pt[i].x : 0.0, yi = act ? (double)s_pt[i].y : 0.0;
        int parent = i;
        bool found = !act;
        for (int j = 0; j < jend; j += 4) {
            double d2[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float2 q = s_pt[min(j + u, C - 1)];
                const double dx = (double)q.x - xi, dy = (double)q.y - yi;
                d2[u] = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (!found && j + u < i && d2[u] <= eps_sq) { parent = j + u; found = true; }
            if (__all_sync(0xffffffffu, found || j + 4 >= i)) break;
        }
        if (act) lab[i] = parent;
    }
    __syncthreads();
    for (int i = tid; i < C; i += blockDim.x) lab[i] = fe_find(s_lab, i);        // a parent is always an ancestor: safe
    __syncthreads();
    for (int i0 = tid & ~31; i0 < C; i0 += blockDim.x) {
        const int i = i0 + lane;
        const bool act = i < C;
        const int jend = min(i0 + 31, C - 1);
        const double xi = act ? (double)s_pt[i].x : 0.0, yi = act ? (double)s_pt[i].y : 0.0;
        const int la = act ? lab[i] : -1;
        for (int j = 0; j < jend; j += 4) {
            int lb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float2 q = s_pt[min(j + u, C - 1)];
                const double dx = (double)q.x - xi, dy = (double)q.y - yi;
                const bool near = act && j + u < i && __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) <= eps_sq;
                lb[u] = near ? lab[j + u] : la;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (lb[u] != la) {                               // rare: a pair within eps across two trees
                    const int hi = max(la, lb[u]), lo = min(la, lb[u]);
                    if (lab[hi] != lo) fe_merge_roots(s_lab, hi, lo);
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    for (int i = tid; i < C; i += blockDim.x) lab[i] = fe_find(s_lab, i);
    __syncthreads();
    // labels in order of first appearance = rank of the component's minimum index; centroid = sequential fp32 mean
    for (int i = tid; i < C; i += blockDim.x) {
        if (s_lab[i] != i) continue;                 // i is its component's first point
        int label = 0;
        for (int j = 0; j < i; ++j) label += (s_lab[j] == j);
        if (label >= FE_MAX_K) { atomicOr(&status[b], FE_ST_K_OVERFLOW); continue; }
        float sx = 0.f, sy = 0.f;
        int cnt = 0;
        for (int j = i; j < C; ++j)
            if (s_lab[j] == i) { sx = __fadd_rn(sx, s_pt[j].x); sy = __fadd_rn(sy, s_pt[j].y); ++cnt; }
        s_cent[label] = make_float2(__fdiv_rn(sx, (float)cnt), __fdiv_rn(sy, (float)cnt));
        atomicMax(&s_k, label + 1);
    }
    __syncthreads();
    const int K = min(s_k, FE_MAX_K);
    for (int k = tid; k < K; k += blockDim.x) s_keep[k] = 0;
    __syncthreads();
    // corners: any filtered scan point within the threshold (landmark_utils.py:78-87)
    const double *f = filtered + (size_t)b * N * 2;
    const int np = g.npts;
    for (int t = tid; t < K * np; t += blockDim.x) {
        const int k = t / np, i = t % np;
        const double dx = (double)s_cent[k].x - f[2 * i], dy = (double)s_cent[k].y - f[2 * i + 1];
        if (__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) <= corner_sq) s_keep[k] = 1;
    }
    __syncthreads();
    if (tid == 0) {
        int out = 0;
        for (int k = 0; k < K; ++k) {
            if (!s_keep[k]) continue;
            const float cx = s_cent[k].x, cy = s_cent[k].y;
            const float q = __fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy));      // x**2 + y**2 on np.float32
            meas[((size_t)b * FE_MAX_K + out) * 2 + 0] = sqrt((double)q);
            meas[((size_t)b * FE_MAX_K + out) * 2 + 1] = atan2((double)cy, (double)cx);
            ++out;
        }
        kcount[b] = out;
    }
}
