// fs2_weights.cuh -- weight normalisation, effective sample size and pose estimate
// (rows A7, A8, A10 of SURVEY.md 8a; reference fast_slam_2.py:161-175, 212-223, 201-210).
//
// Two passes over w[P] (8 B per particle each, HBM-bound and tiny next to the map stream):
//   fs2_weight_total_kernel : per-block tree sums of w, the last block to finish adds the block results
//                             in block order (deterministic) -> stats[FS2_STAT_TOTAL].
//   fs2_normalize_kernel    : applies the reference's rule (Q8) with the total, and reduces sum w^2 and
//                             the first arg-max of the normalised weights the same way; the last block
//                             writes Neff (Q9) and the pose of the arg-max particle (Q11).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FS2_RED_THREADS 256
#define FS2_RED_MAX_BLOCKS 1184  // 148 SMs x 8

struct Fs2MaxIdx {
    double v;
    long long i;
};

// larger weight wins, ties go to the lower index (max(..., key=) keeps the first, fast_slam_2.py:208)
__device__ __forceinline__ Fs2MaxIdx fs2_better(Fs2MaxIdx a, Fs2MaxIdx b)
{
    if (b.i < 0) return a;
    if (a.i < 0) return b;
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}

__device__ __forceinline__ double fs2_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ Fs2MaxIdx fs2_warp_best(Fs2MaxIdx m)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Fs2MaxIdx t;
        t.v = __shfl_xor_sync(0xffffffffu, m.v, o);
        t.i = __shfl_xor_sync(0xffffffffu, m.i, o);
        m = fs2_better(m, t);
    }
    return m;
}

// returns true in exactly one thread (thread 0 of the last block to arrive)
__device__ __forceinline__ bool fs2_last_block(unsigned int *counter)
{
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(counter, 1u);
        last = (t == gridDim.x - 1);
        if (last) *counter = 0;  // re-arm for the next launch
    }
    __syncthreads();
    return last;
}

__global__ void __launch_bounds__(FS2_RED_THREADS)
fs2_weight_total_kernel(const double *__restrict__ w, int64_t P, double *partial, unsigned int *counter, double *stats,
                        int *ctl_clear = nullptr)
{
    __shared__ double ws[FS2_RED_THREADS / 32];
    // fused step (fs2_step.cuh): a fresh set of control words for the kernels behind this one
    if (ctl_clear && blockIdx.x == 0 && threadIdx.x < 8) ctl_clear[threadIdx.x] = 0;
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) s += w[i];
    s = fs2_warp_sum(s);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < FS2_RED_THREADS / 32; ++k) t += ws[k];
        partial[blockIdx.x] = t;
    }
    if (fs2_last_block(counter)) {
        // thread-strided, then a fixed-order tree: deterministic for a given grid
        double t = 0.0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) t += ((volatile double *)partial)[b];
        t = fs2_warp_sum(t);
        if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = t;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int k = 0; k < FS2_RED_THREADS / 32; ++k) tot += ws[k];
            stats[0] = tot;  // FS2_STAT_TOTAL
        }
    }
}

__global__ void __launch_bounds__(FS2_RED_THREADS)
fs2_normalize_kernel(double *w, const double *x, const double *y, const double *yaw, int64_t P, int64_t Pglobal,
                     const double *total_dev, int apply, double *partial_sq, Fs2MaxIdx *partial_best,
                     unsigned int *counter, double *stats, const int32_t *ids = nullptr, int64_t goff = 0)
{
    __shared__ double ws[FS2_RED_THREADS / 32];
    __shared__ Fs2MaxIdx wb[FS2_RED_THREADS / 32];
    const double total = *total_dev;
    const bool reset = total < 1e-5;            // fast_slam_2.py:168-170
    const double uni = 1.0 / (double)Pglobal;
    double sq = 0.0;
    Fs2MaxIdx best;
    best.v = 0.0; best.i = -1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
        double v = w[i];
        if (apply) {
            if (reset) v = uni;
            else if (!(v < 1e-5)) v = v / total;  // :173 (weights below 1e-5 are left as they are)
            w[i] = v;
        }
        sq = fma(v, v, sq);
        Fs2MaxIdx c;
        c.v = v; c.i = i;
        best = fs2_better(best, c);
    }
    sq = fs2_warp_sum(sq);
    best = fs2_warp_best(best);
    if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5] = sq; wb[threadIdx.x >> 5] = best; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        Fs2MaxIdx b = wb[0];
        for (int k = 0; k < FS2_RED_THREADS / 32; ++k) { t += ws[k]; if (k) b = fs2_better(b, wb[k]); }
        partial_sq[blockIdx.x] = t;
        partial_best[blockIdx.x] = b;
    }
    if (fs2_last_block(counter)) {
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
        if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5] = t; wb[threadIdx.x >> 5] = b; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            Fs2MaxIdx bb = wb[0];
            for (int k = 0; k < FS2_RED_THREADS / 32; ++k) { s += ws[k]; if (k) bb = fs2_better(bb, wb[k]); }
            stats[1] = s;                                                        // FS2_STAT_SUMSQ
            stats[2] = (s < 1.0 / (double)Pglobal) ? (double)Pglobal : 1.0 / s;  // FS2_STAT_NEFF (:220-223)
            stats[3] = bb.v;                                                     // FS2_STAT_WMAX
            stats[4] = (double)bb.i;                                             // FS2_STAT_ARGMAX
            if (bb.i >= 0) { stats[5] = x[bb.i]; stats[6] = y[bb.i]; stats[7] = yaw[bb.i]; }
            // its LOGICAL (global) index: the local order of a shard is the logical order of its particles, so the
            // first local arg-max is the lowest logical one; ties between shards go by this id (fast_slam_2.py:208)
            stats[12] = (bb.i >= 0) ? (double)(ids ? (int64_t)ids[bb.i] : goff + bb.i) : -1.0;   // FS2_STAT_ARGMAX_ID
        }
    }
}
