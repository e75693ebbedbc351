// fs2_step.cuh -- the weight half of FastSLAM2.iterate (rows A7-A10, A16 of SURVEY.md 8a; reference
// fast_slam_2.py:56-67, 161-223) as ONE asynchronous chain of launches with the resampling decision taken on the
// device: the host enqueues everything, reads the 128-byte stats block once and synchronises once per step.
//
//   fs2_normalize_scan_kernel   A7 + A8 + A10: normalise (Q8), sum w^2, first arg-max; the plain block sums of the
//                               normalised weights (first stage of the exact scan, fs2_resample.cuh) come out of the
//                               same pass.  The last block writes Neff, the estimate, the DECISION ctl[RES] =
//                               (Neff < N/2, fast_slam_2.py:62) and the exclusive prefix of the block sums.
//   -- everything below returns at once when ctl[RES] == 0 --
//   fs2_scan_groupfunc          block and group parity functions; the last CTA to finish walks them with the exact
//                               running sum (fs2_resample.cuh).  In the fused step it also clears the slot-use flags
//                               and the scan status words, and weights the exact scan does not accept (negative, NaN)
//                               take the literal serial loop there.
//   fs2_scan_emit               exact c_k (fs2_resample.cuh)
//   fs2_search_mark_kernel      ancestor of every slot by binary search (Q10) + first-offspring / extra-offspring marks
//   fs2_iscan_kernel            ONE-pass scan (decoupled look-back) of the two flag arrays -> copy tasks and free slots
//   fs2_gather_kernel           pose / weight / count of every new particle, then one warp per copied map
//   fs2_commit_estimate_kernel  publishes the new poses and redoes the arg-max over the copied weights (Q11)
#pragma once
#include "fs2_resample.cuh"
#include "fs2_weights.cuh"

// ---- A7/A8/A10 + block sums ------------------------------------------------------------------------------
__global__ void __launch_bounds__(FS2_SCAN_T)
fs2_normalize_scan_kernel(double *w, const double *x, const double *y, const double *yaw, int64_t P, int64_t Pglobal,
                          const double *total_dev, double *partial_sq, Fs2MaxIdx *partial_best, unsigned int *counter,
                          double *stats, double *bsum, double *bpre, int nb, int *ctl, int decide)
{
    static_assert(FS2_SCAN_B == FS2_SCAN_T, "one weight per thread and tile");
    __shared__ double ws[FS2_SCAN_T / 32];
    __shared__ Fs2MaxIdx wb[FS2_SCAN_T / 32];
    __shared__ double wt[FS2_SCAN_T / 32];
    const double total = *total_dev;
    const bool reset = total < 1e-5;            // fast_slam_2.py:168-170
    const double uni = 1.0 / (double)Pglobal;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double sq = 0.0;
    Fs2MaxIdx best;
    best.v = 0.0; best.i = -1;
    bool bad = false;
    for (int t = blockIdx.x; t < nb; t += gridDim.x) {
        const int64_t i = (int64_t)t * FS2_SCAN_B + threadIdx.x;
        double v = 0.0;
        if (i < P) {
            v = w[i];
            if (reset) v = uni;
            else if (!(v < 1e-5)) v = v / total;  // :173 (weights below 1e-5 are left as they are)
            w[i] = v;
            bad |= !(v >= 0.0) || !(v < 1.797e308);
            sq = fma(v, v, sq);
            Fs2MaxIdx c;
            c.v = v; c.i = i;
            best = fs2_better(best, c);
        }
        const double s = fs2_warp_sum((i < P) ? v : 0.0);
        if (lane == 0) wt[wid] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tsum = 0.0;
            for (int k = 0; k < FS2_SCAN_T / 32; ++k) tsum += wt[k];
            bsum[t] = tsum;
        }
        __syncthreads();
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(&ctl[FS2_CTL_ANOMALY], 1);
    sq = fs2_warp_sum(sq);
    best = fs2_warp_best(best);
    if (lane == 0) { ws[wid] = sq; wb[wid] = best; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        Fs2MaxIdx b = wb[0];
        for (int k = 0; k < FS2_SCAN_T / 32; ++k) { t += ws[k]; if (k) b = fs2_better(b, wb[k]); }
        partial_sq[blockIdx.x] = t;
        partial_best[blockIdx.x] = b;
    }
    if (!fs2_last_block(counter)) return;
    // ---- last block: global statistics, the decision, and the (approximate) exclusive prefix of the block sums ----
    {
        double t = 0.0;
        Fs2MaxIdx b;
        b.v = 0.0; b.i = -1;
        for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) {
            t += ((volatile double *)partial_sq)[k];
            Fs2MaxIdx c;
            c.v = ((volatile double *)&partial_best[k].v)[0];
            c.i = ((volatile long long *)&partial_best[k].i)[0];
            b = fs2_better(b, c);
        }
        t = fs2_warp_sum(t);
        b = fs2_warp_best(b);
        if (lane == 0) { ws[wid] = t; wb[wid] = b; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            Fs2MaxIdx bb = wb[0];
            for (int k = 0; k < FS2_SCAN_T / 32; ++k) { s += ws[k]; if (k) bb = fs2_better(bb, wb[k]); }
            const double neff = (s < 1.0 / (double)Pglobal) ? (double)Pglobal : 1.0 / s;  // :220-223
            stats[1] = s;               // FS2_STAT_SUMSQ
            stats[2] = neff;            // FS2_STAT_NEFF
            stats[3] = bb.v;            // FS2_STAT_WMAX
            stats[4] = (double)bb.i;    // FS2_STAT_ARGMAX
            if (bb.i >= 0) { stats[5] = x[bb.i]; stats[6] = y[bb.i]; stats[7] = yaw[bb.i]; }
            const int res = (decide && neff < (double)Pglobal / 2.0) ? 1 : 0;             // fast_slam_2.py:62
            ctl[FS2_CTL_RES] = res;
            stats[8] = (double)res;     // FS2_STAT_RESAMPLED
            stats[9] = 0.0;             // FS2_STAT_COPIES
            stats[13] = 0.0;            // FS2_STAT_DEFERRED
            stats[10] = (double)((volatile int *)ctl)[FS2_CTL_ANOMALY];
            stats[11] = 0.0;            // FS2_STAT_STUCK
        }
    }
    // exclusive prefix of bsum[0..nb): thread-contiguous runs, warp scan, serial over the 8 warp totals
    __shared__ double carry_s;
    if (threadIdx.x == 0) carry_s = 0.0;
    __syncthreads();
    const int per = 8;                               // block sums per thread and round
    for (int base = 0; base < nb; base += FS2_SCAN_T * per) {
        const int i0 = base + threadIdx.x * per;
        double loc[per];
        double run = 0.0;
#pragma unroll
        for (int j = 0; j < per; ++j) {
            const int i = i0 + j;
            const double v = (i < nb) ? ((volatile double *)bsum)[i] : 0.0;
            loc[j] = run;
            run += v;
        }
        double inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double tt = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += tt;
        }
        if (lane == 31) wt[wid] = inc;
        __syncthreads();
        double off = carry_s;
        for (int k = 0; k < wid; ++k) off += wt[k];
        off += inc - run;
#pragma unroll
        for (int j = 0; j < per; ++j) {
            const int i = i0 + j;
            if (i < nb) bpre[i] = off + loc[j];
        }
        __syncthreads();
        if (threadIdx.x == FS2_SCAN_T - 1) carry_s = off + run;
        __syncthreads();
    }
}

// ancestor of every slot + the marks of the copy-on-resample gather (fs2_gather_mark), one kernel.
// extra[m]: 0 = first offspring of its ancestor (keeps the ancestor's map slot), otherwise an extra offspring that needs
// a map of its own.  With `defer` (deferred map copies, fs2_update_ws.cuh) the offspring of an ancestor -- consecutive
// particles m0 .. m1 -- are cut into groups of 1 + FS2_SIBMAX: rank j = m - m0 with j % 8 == 0 is a LEADER (extra[m] = 2
// for j > 0: the gather copies its map now), the others are FOLLOWERS (extra[m] = 1: a free slot now, the copy from the
// next update kernel).  Leaders go on the `leaders` list (atomic append: any order), nfol[m] = followers of leader m.
// m0 and the group's extent come from the sample points themselves: u_q = u0 + q/n is monotone in q and
// ancestor(q) == a  <=>  cum[a-1] < u_q <= cum[a]  (the search below; a == n-1 also takes everything beyond the total).
__global__ void __launch_bounds__(256)
fs2_search_mark_kernel(const double *__restrict__ cum, int64_t n, double u0, int32_t *ancestor, const int32_t *__restrict__ slot,
                       int32_t *used, int32_t *extra, int *ctl, int defer, int32_t *leaders, int32_t *nfol, int32_t *dctl)
{
    if (!((volatile int *)ctl)[FS2_CTL_RES]) return;
    const bool serial = ((volatile int *)ctl)[FS2_CTL_SERIAL] != 0;
    const bool dfr = defer && !serial;          // (the literal serial path hands over ancestors the rule above does not describe)
    const double inv = 1.0 / (double)n;
    const int lane = threadIdx.x & 31;
    auto uof = [&](int64_t m) -> double { return __dadd_rn(u0, __dmul_rn((double)m, inv)); };
    auto search = [&](int64_t m) -> int {
        const double u = uof(m);
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (u > cum[mid]) lo = mid + 1; else hi = mid;
        }
        if (lo >= n) { lo = n - 1; ctl[FS2_CTL_STUCK] = 1; }
        return (int)lo;
    };
    if (dfr && blockIdx.x == 0 && threadIdx.x == 0) dctl[0] = 1;
    const int64_t nround = (n + 31) & ~(int64_t)31;      // whole warps stay in the loop (shuffles below)
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < nround; m += (int64_t)gridDim.x * blockDim.x) {
        int a = -1;
        if (m < n) {
            a = serial ? ancestor[m] : search(m);
            if (!serial) ancestor[m] = a;
        }
        int prev = __shfl_up_sync(0xffffffffu, a, 1);
        if (lane == 0 && m > 0 && m < n) prev = serial ? ancestor[m - 1] : search(m - 1);
        bool leader = false;
        if (m < n) {
            const bool first = (m == 0) || (prev != a);
            int code = first ? 0 : 1;
            if (dfr) {
                int64_t m0 = m;
                if (!first) {                                  // first offspring of a: the lowest q with u_q > cum[a-1]
                    int64_t lo = 0, hi = m;
                    if (a > 0) {
                        const double cl = cum[a - 1];
                        while (lo < hi) {
                            const int64_t mid = (lo + hi) >> 1;
                            if (uof(mid) > cl) hi = mid; else lo = mid + 1;
                        }
                    } else hi = 0;
                    m0 = (a > 0) ? lo : 0;
                }
                leader = (((m - m0) & (int64_t)FS2_SIBMAX) == 0);
                int nf = 0;
                if (leader) {
                    const bool last = (a == (int)(n - 1));
                    const double ch = cum[a];
#pragma unroll
                    for (int k = 1; k <= FS2_SIBMAX; ++k)
                        if (m + k < n && (last || uof(m + k) <= ch)) ++nf;      // monotone: the true ones come first
                    if (!first) code = 2;
                }
                nfol[m] = nf;
            }
            extra[m] = code;
            if (first) used[slot[a]] = 1;
        }
        if (dfr) {
            const unsigned lb = __ballot_sync(0xffffffffu, leader);
            int base = 0;
            if (lane == 0 && lb) base = atomicAdd(&dctl[1], __popc(lb));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (leader) leaders[base + __popc(lb & ((1u << lane) - 1u))] = (int32_t)m;
        }
    }
}

// ---- one-pass exclusive scan of two flag arrays (decoupled look-back) ------------------------------------------
// flags: extra[i] (i < P) and "slot i is free" = !used[i] (i < S).  Outputs tasks[r] = the r-th extra offspring's
// particle, freeslot[r] = the r-th free slot, ncopies = (extras, free slots).  State word per tile:
// [63:62] 0 = nothing yet, 1 = tile aggregate, 2 = inclusive prefix; [61:31] extras, [30:0] free (P, S < 2^31).
#define FS2_IS_TILE 2048
#define FS2_IS_T 256
#define FS2_IS_PER (FS2_IS_TILE / FS2_IS_T)

__device__ __forceinline__ unsigned long long fs2_is_pack(unsigned st, unsigned a, unsigned d)
{
    return ((unsigned long long)st << 62) | ((unsigned long long)a << 31) | (unsigned long long)d;
}

__global__ void __launch_bounds__(FS2_IS_T)
fs2_iscan_kernel(const int32_t *__restrict__ extra, const int32_t *__restrict__ used, int64_t P, int64_t S,
                 unsigned long long *state, unsigned int *ticket, int32_t *tasks, int32_t *freeslot, int32_t *ncopies,
                 int ntiles, const int *ctl, double *stats)
{
    if (!((volatile int *)ctl)[FS2_CTL_RES]) return;
    __shared__ unsigned s_tile;
    __shared__ uint2 wsum[FS2_IS_T / 32];
    __shared__ uint2 s_prefix;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);       // tiles are handed out in order: a tile only ever waits for earlier ones
    __syncthreads();
    const unsigned tile = s_tile;
    if ((int)tile >= ntiles) return;
    const int64_t base = (int64_t)tile * FS2_IS_TILE + (int64_t)threadIdx.x * FS2_IS_PER;
    unsigned fa = 0, fd = 0;                 // bit j: flag of element base + j
    unsigned ca = 0, cd = 0;
    static_assert(FS2_IS_PER == 8, "two 16-byte loads per flag array");
    int ve[FS2_IS_PER], vu[FS2_IS_PER];
    if (base + FS2_IS_PER <= P) {
        *reinterpret_cast<int4 *>(ve) = *reinterpret_cast<const int4 *>(extra + base);
        *reinterpret_cast<int4 *>(ve + 4) = *reinterpret_cast<const int4 *>(extra + base + 4);
    } else {
#pragma unroll
        for (int j = 0; j < FS2_IS_PER; ++j) ve[j] = (base + j < P) ? extra[base + j] : 0;
    }
    if (base + FS2_IS_PER <= S) {
        *reinterpret_cast<int4 *>(vu) = *reinterpret_cast<const int4 *>(used + base);
        *reinterpret_cast<int4 *>(vu + 4) = *reinterpret_cast<const int4 *>(used + base + 4);
    } else {
#pragma unroll
        for (int j = 0; j < FS2_IS_PER; ++j) vu[j] = (base + j < S) ? used[base + j] : 1;
    }
#pragma unroll
    for (int j = 0; j < FS2_IS_PER; ++j) {
        if (ve[j]) { fa |= 1u << j; ++ca; }
        if (vu[j] == 0) { fd |= 1u << j; ++cd; }
    }
    unsigned ia = ca, id = cd;               // inclusive over the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned ta = __shfl_up_sync(0xffffffffu, ia, o), td = __shfl_up_sync(0xffffffffu, id, o);
        if (lane >= o) { ia += ta; id += td; }
    }
    if (lane == 31) wsum[wid] = make_uint2(ia, id);
    __syncthreads();
    unsigned oa = 0, od = 0, ta = 0, td = 0;
    for (int k = 0; k < FS2_IS_T / 32; ++k) {
        if (k < wid) { oa += wsum[k].x; od += wsum[k].y; }
        ta += wsum[k].x; td += wsum[k].y;
    }
    // publish the aggregate, then look back for the exclusive prefix of this tile
    if (threadIdx.x == 0) {
        volatile unsigned long long *vs = state;
        if (tile == 0) {
            vs[0] = fs2_is_pack(2u, ta, td);
            s_prefix = make_uint2(0u, 0u);
        } else {
            vs[tile] = fs2_is_pack(1u, ta, td);
            __threadfence();
            unsigned pa = 0, pd = 0;
            int t = (int)tile - 1;
            for (;;) {
                unsigned long long v;
                do { v = vs[t]; } while ((v >> 62) == 0ull);
                pa += (unsigned)((v >> 31) & 0x7fffffffull);
                pd += (unsigned)(v & 0x7fffffffull);
                if ((v >> 62) == 2ull) break;
                --t;
            }
            vs[tile] = fs2_is_pack(2u, pa + ta, pd + td);
            s_prefix = make_uint2(pa, pd);
        }
        if ((int)tile == ntiles - 1) {
            const uint2 pf = s_prefix;
            ncopies[0] = (int32_t)(pf.x + ta);      // copies needed
            ncopies[1] = (int32_t)(pf.y + td);      // free slots
            stats[9] = (double)(pf.x + ta);         // FS2_STAT_COPIES
        }
    }
    __syncthreads();
    unsigned ra = s_prefix.x + oa + ia - ca, rd = s_prefix.y + od + id - cd;    // exclusive rank of this thread's first flag
#pragma unroll
    for (int j = 0; j < FS2_IS_PER; ++j) {
        const int64_t i = base + j;
        if ((fa >> j) & 1u) tasks[ra++] = (int32_t)i;
        if ((fd >> j) & 1u) freeslot[rd++] = (int32_t)i;
    }
}

// ---- pose + map copies of the local gather, one kernel (fs2_gather_pose then fs2_gather_copy) -------------------
// When the marks were made for deferred copies (dctl[0] set by fs2_search_mark_kernel) a follower (extra == 1) only gets
// its free slot here.
__global__ void __launch_bounds__(256)
fs2_gather_kernel(const int32_t *__restrict__ anc, const int32_t *__restrict__ extra, int64_t P, const double *x,
                  const double *y, const double *yaw, const double *w, const int32_t *count, const int32_t *slot,
                  double *x2, double *y2, double *yaw2, double *w2, int32_t *count2, int32_t *slot2,
                  const int32_t *__restrict__ tasks, const int32_t *__restrict__ freeslot, const int32_t *ncopies, double *lm,
                  int lcap, const int *ctl, const int32_t *dctl)
{
    if (!((volatile int *)ctl)[FS2_CTL_RES]) return;
    const bool dfr = dctl && ((volatile const int32_t *)dctl)[0] != 0;
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < P; m += (int64_t)gridDim.x * blockDim.x) {
        const int a = anc[m];
        x2[m] = x[a]; y2[m] = y[a]; yaw2[m] = yaw[a]; w2[m] = w[a]; count2[m] = count[a];
        if (!extra[m]) slot2[m] = slot[a];
    }
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int n = min(ncopies[0], ncopies[1]);      // without spare slots: copies == free slots
    const size_t stride = 6 * (size_t)lcap;
    if (dfr) {                                      // the followers' slots, a thread each
        for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
            const int m = tasks[r];
            if (extra[m] == 1) slot2[m] = freeslot[r];
        }
    }
    for (int64_t r = warp; r < n; r += nwarps) {
        const int m = tasks[r];
        if (dfr && extra[m] == 1) continue;
        const int a = anc[m];
        const int dst_slot = freeslot[r];
        const int4 *src = reinterpret_cast<const int4 *>(lm + (size_t)slot[a] * stride);
        const int ng = count[a] * 3;                   // 16-byte granules
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

// the copies a deferred-copy resample left to the next update kernel, made here instead: whoever touches the maps
// before an update has run (a download, the map clustering, another resample, a motion-only step) calls this first.
// One warp per leader; its map is untouched since the resample.
__global__ void __launch_bounds__(256)
fs2_materialize_kernel(const int32_t *dctl, const int32_t *__restrict__ leaders, const int32_t *__restrict__ nfol,
                       const int32_t *__restrict__ slot, const int32_t *__restrict__ count, double *lm, int lcap)
{
    if (!((volatile const int32_t *)dctl)[0]) return;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int n = dctl[1];
    const size_t stride = 6 * (size_t)lcap;
    for (int64_t r = warp; r < n; r += nwarps) {
        const int p = leaders[r];
        const int nf = nfol[p];
        if (nf == 0) continue;
        const int4 *src = reinterpret_cast<const int4 *>(lm + (size_t)slot[p] * stride);
        const int ng = count[p] * 3;
        for (int k = 1; k <= nf; ++k) {
            int4 *dst = reinterpret_cast<int4 *>(lm + (size_t)slot[p + k] * stride);
            for (int g = lane; g < ng; g += 32) dst[g] = src[g];
        }
    }
}

// ---- publish the gathered poses and redo the arg-max over the copied weights (Q11) ------------------------------
__global__ void __launch_bounds__(FS2_RED_THREADS)
fs2_commit_estimate_kernel(double *x, double *y, double *yaw, double *w, int32_t *count, int32_t *slot, const double *x2,
                           const double *y2, const double *yaw2, const double *w2, const int32_t *count2,
                           const int32_t *slot2, int64_t P, Fs2MaxIdx *partial_best, unsigned int *counter, double *stats,
                           int *ctl, const int32_t *dctl)
{
    if (!((volatile int *)ctl)[FS2_CTL_RES]) return;
    __shared__ Fs2MaxIdx wb[FS2_RED_THREADS / 32];
    Fs2MaxIdx best;
    best.v = 0.0; best.i = -1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = w2[i];
        x[i] = x2[i]; y[i] = y2[i]; yaw[i] = yaw2[i]; w[i] = v; count[i] = count2[i]; slot[i] = slot2[i];
        Fs2MaxIdx c;
        c.v = v; c.i = i;
        best = fs2_better(best, c);
    }
    best = fs2_warp_best(best);
    if ((threadIdx.x & 31) == 0) wb[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        Fs2MaxIdx b = wb[0];
        for (int k = 1; k < FS2_RED_THREADS / 32; ++k) b = fs2_better(b, wb[k]);
        partial_best[blockIdx.x] = b;
    }
    if (!fs2_last_block(counter)) return;
    Fs2MaxIdx b;
    b.v = 0.0; b.i = -1;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) {
        Fs2MaxIdx c;
        c.v = ((volatile double *)&partial_best[k].v)[0];
        c.i = ((volatile long long *)&partial_best[k].i)[0];
        b = fs2_better(b, c);
    }
    b = fs2_warp_best(b);
    if ((threadIdx.x & 31) == 0) wb[threadIdx.x >> 5] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        Fs2MaxIdx bb = wb[0];
        for (int k = 1; k < FS2_RED_THREADS / 32; ++k) bb = fs2_better(bb, wb[k]);
        stats[3] = bb.v;                       // FS2_STAT_WMAX / ARGMAX / estimate after the resample
        stats[4] = (double)bb.i;
        if (bb.i >= 0) { stats[5] = x2[bb.i]; stats[6] = y2[bb.i]; stats[7] = yaw2[bb.i]; }
        stats[11] = (double)((volatile int *)ctl)[FS2_CTL_STUCK];
        // FS2_STAT_DEFERRED: followers, whose map copy the next update kernel writes
        stats[13] = (dctl && ((volatile const int32_t *)dctl)[0]) ? (double)(P - (int64_t)((volatile const int32_t *)dctl)[1]) : 0.0;
    }
}
