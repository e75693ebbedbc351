// fs2_icp.cuh -- ICP.get_transformation (fast_slam_2/algorithms/icp.py:13-90), batched: one thread block per pair
// of point sets.  Per iteration, as in the reference: nearest target point of every source point (scipy KDTree.query
// there, exhaustive search here -- the same neighbour, lowest index on an exact tie), best_fit_transform of the
// matched pairs (icp.py:60-90), apply it to the source points, compose the totals, stop when the mean distance
// changes by less than the threshold.  The 2 x 2 SVD of best_fit_transform is replaced by its closed form: the
// proper rotation that maximises trace(R H), H = sum (s - cs)(t - ct)^T, is the rotation by
// atan2(H01 - H10, H00 + H11), which is also what "V U^T with the reflection fix" (icp.py:76-85) yields.  Not on the
// reference's shipped loop (icp.py:8-9: unused); row N4 of SURVEY.md 8f.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define ICP_THREADS 256
#define ICP_MAX_POINTS 4096

__device__ __forceinline__ double icp_block_sum(double v, double *ws)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int k = 0; k < ICP_THREADS / 32; ++k) t += ws[k];
    return t;
}

__global__ void __launch_bounds__(ICP_THREADS)
icp_kernel(const double *source, const double *target, int ns, int nt, int max_iterations, double threshold, double *rot,
           double *trans, int *iters)
{
    extern __shared__ __align__(16) unsigned char icp_smem[];
    __shared__ double ws[ICP_THREADS / 32];
    double2 *S = reinterpret_cast<double2 *>(icp_smem);       // current source points
    double2 *T = S + ns;                                      // target points
    int *nn = reinterpret_cast<int *>(T + nt);                // nearest target of each source point
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < ns; i += ICP_THREADS) S[i] = reinterpret_cast<const double2 *>(source + (size_t)b * ns * 2)[i];
    for (int i = tid; i < nt; i += ICP_THREADS) T[i] = reinterpret_cast<const double2 *>(target + (size_t)b * nt * 2)[i];
    __syncthreads();
    double r00 = 1.0, r01 = 0.0, r10 = 0.0, r11 = 1.0, tx = 0.0, ty = 0.0;      // totals (icp.py:31-32)
    double prev = __longlong_as_double(0x7ff0000000000000ll);                   // +inf
    int it = 0;
    for (; it < max_iterations;) {
        // nearest neighbours and their distances
        double dsum = 0.0, ssx = 0.0, ssy = 0.0, stx = 0.0, sty = 0.0;
        for (int i = tid; i < ns; i += ICP_THREADS) {
            const double2 s = S[i];
            double best = __longlong_as_double(0x7ff0000000000000ll);
            int bj = 0;
            for (int j = 0; j < nt; ++j) {
                const double dx = s.x - T[j].x, dy = s.y - T[j].y;
                const double d2 = dx * dx + dy * dy;
                if (d2 < best) { best = d2; bj = j; }
            }
            nn[i] = bj;
            dsum += sqrt(best);
            ssx += s.x; ssy += s.y; stx += T[bj].x; sty += T[bj].y;
        }
        const double mean_error = icp_block_sum(dsum, ws) / (double)ns;
        const double csx = icp_block_sum(ssx, ws) / (double)ns, csy = icp_block_sum(ssy, ws) / (double)ns;
        const double ctx = icp_block_sum(stx, ws) / (double)ns, cty = icp_block_sum(sty, ws) / (double)ns;
        // H = centered_source^T centered_target (icp.py:74)
        double h00 = 0.0, h01 = 0.0, h10 = 0.0, h11 = 0.0;
        for (int i = tid; i < ns; i += ICP_THREADS) {
            const double ax = S[i].x - csx, ay = S[i].y - csy;
            const double2 t = T[nn[i]];
            const double bx = t.x - ctx, by = t.y - cty;
            h00 += ax * bx; h01 += ax * by; h10 += ay * bx; h11 += ay * by;
        }
        h00 = icp_block_sum(h00, ws); h01 = icp_block_sum(h01, ws); h10 = icp_block_sum(h10, ws); h11 = icp_block_sum(h11, ws);
        const double th = atan2(h01 - h10, h00 + h11);
        double sn, cs;
        sincos(th, &sn, &cs);
        const double q00 = cs, q01 = -sn, q10 = sn, q11 = cs;
        const double ux = ctx - (q00 * csx + q01 * csy), uy = cty - (q10 * csx + q11 * csy);      // icp.py:88
        __syncthreads();
        for (int i = tid; i < ns; i += ICP_THREADS) {                                            // icp.py:44
            const double2 s = S[i];
            S[i] = make_double2(q00 * s.x + q01 * s.y + ux, q10 * s.x + q11 * s.y + uy);
        }
        // totals (icp.py:47-48)
        const double n00 = q00 * r00 + q01 * r10, n01 = q00 * r01 + q01 * r11;
        const double n10 = q10 * r00 + q11 * r10, n11 = q10 * r01 + q11 * r11;
        const double ntx = q00 * tx + q01 * ty + ux, nty = q10 * tx + q11 * ty + uy;
        r00 = n00; r01 = n01; r10 = n10; r11 = n11; tx = ntx; ty = nty;
        ++it;
        __syncthreads();
        if (fabs(prev - mean_error) < threshold) break;                                          // icp.py:51-53
        prev = mean_error;
    }
    if (tid == 0) {
        rot[4 * b + 0] = r00; rot[4 * b + 1] = r01; rot[4 * b + 2] = r10; rot[4 * b + 3] = r11;
        trans[2 * b + 0] = tx; trans[2 * b + 1] = ty;
        if (iters) iters[b] = it;
    }
}
