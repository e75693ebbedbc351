// fs2_place.cuh -- where the offspring of a global resample go when the particles are sharded over several GPUs
// (reference fast_slam_2.py:177-199).
//
// The reference's particle list has an ORDER (the running sum of the resampler walks it, ties of the arg-max go to
// the lowest index, the serializer writes it out), but nothing says where particle m has to live.  Keeping logical
// slot m on GPU m / P makes every offspring whose ancestor sat on another GPU a map transfer over NVLink -- and the
// reference never resets weights after a resample, so whole shards descend from one GPU's particles (round 1: up to
// 640 k maps of 12.6 KB pulled out of ONE GPU per resample while the other links idled).  Here an offspring stays on
// its ancestor's GPU; only what exceeds a GPU's P particle slots moves, whole fat lineages first (one transfer, the
// duplicates are local copies at the destination), into the GPUs that fall short.
//
// Every rank runs the same plan on the same replicated inputs (the logical ancestors anc[m] of all N new particles,
// the table place[m] = rank * P + local index of logical particle m), so all ranks agree without exchanging a word:
//
//   pl_home        home[m] = rank of place[anc[m]]; offspring per rank; offspring per ancestor
//   pl_hist        per exporting rank, offspring by size of their lineage (sizes >= 255 share the last bin)
//   pl_decide      per rank: surplus over P, the lineage-size threshold T (largest T whose lineages of size >= T hold
//                  at least the surplus), the intervals of the export sequence each importing rank takes
//   pl_mcscan_*    exclusive prefix per class over m (class = home rank of the eligible offspring; then destination)
//   pl_dest        exported = eligible and among the first `surplus` of its home rank; destination by interval
//   pl_finish      place_new[m] = dest * P + (offspring with the same destination before m); for this rank's new
//                  particles: logical id, physical source, logical ancestor; marks for ancestors other ranks will read
//
// New local particles are numbered in logical order, so the offspring of one ancestor are consecutive on every GPU
// and the local order of a shard is the logical order of its particles (ties of the arg-max, serialisation).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PL_MAXR 16            // ranks
#define PL_HB 256             // lineage-size bins
#define PL_TILE 1024          // elements per tile of the multi-class scan
#define PL_T 256

struct PlPlan {               // device-resident scratch of the plan, sized for N = global particles
    int64_t N, P;
    int world, rank;
    uint8_t *home, *dest, *cls;      // [N]
    int32_t *ocnt;                   // [N] offspring per (logical) ancestor
    int32_t *prefix;                 // [N] exclusive prefix of the last multi-class scan
    unsigned *tilecnt;               // [ntiles][PL_MAXR]
    unsigned *cnt;                   // [PL_MAXR] offspring born on each rank
    unsigned *hist;                  // [PL_MAXR][PL_HB]
    int *surplus;                    // [PL_MAXR] > 0: exports that many, < 0: imports
    int *thresh;                     // [PL_MAXR]
    unsigned *expoff;                // [PL_MAXR] start of each exporting rank in the export sequence
    unsigned *impend;                // [PL_MAXR] end of each rank's interval of the export sequence (cumulative deficits)
    unsigned long long *info;        // [4] exported offspring, (lineage, destination) pairs = maps pulled over NVLink, ...
};

__global__ void __launch_bounds__(256)
pl_home_kernel(PlPlan pl, const int32_t *__restrict__ anc, const int32_t *__restrict__ place)
{
    __shared__ unsigned sc[PL_MAXR];
    if (threadIdx.x < PL_MAXR) sc[threadIdx.x] = 0u;
    __syncthreads();
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < pl.N; m += (int64_t)gridDim.x * blockDim.x) {
        const int a = anc[m];
        const int r = (int)(place[a] / pl.P);
        pl.home[m] = (uint8_t)r;
        atomicAdd(&sc[r], 1u);
        atomicAdd(&pl.ocnt[a], 1);
    }
    __syncthreads();
    if (threadIdx.x < PL_MAXR && sc[threadIdx.x]) atomicAdd(&pl.cnt[threadIdx.x], sc[threadIdx.x]);
}

__global__ void __launch_bounds__(256)
pl_hist_kernel(PlPlan pl, const int32_t *__restrict__ anc)
{
    __shared__ unsigned sh[PL_MAXR * PL_HB];
    for (int i = threadIdx.x; i < PL_MAXR * PL_HB; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < pl.N; m += (int64_t)gridDim.x * blockDim.x) {
        const int r = pl.home[m];
        if ((int64_t)pl.cnt[r] > pl.P) {
            const int sz = min(pl.ocnt[anc[m]], PL_HB - 1);
            atomicAdd(&sh[r * PL_HB + sz], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PL_MAXR * PL_HB; i += blockDim.x)
        if (sh[i]) atomicAdd(&pl.hist[i], sh[i]);
}

// one block: surplus, threshold, export offsets, import intervals
__global__ void pl_decide_kernel(PlPlan pl)
{
    if (threadIdx.x != 0) return;
    unsigned off = 0, acc = 0;
    for (int r = 0; r < pl.world; ++r) {
        const int s = (int)((int64_t)pl.cnt[r] - pl.P);
        pl.surplus[r] = s;
        pl.expoff[r] = off;
        int T = 1;
        if (s > 0) {
            // largest T such that the lineages of size >= T (born on r) hold at least s offspring
            unsigned long long held = 0;
            T = 1;
            for (int b = PL_HB - 1; b >= 1; --b) {
                held += pl.hist[r * PL_HB + b];
                if (held >= (unsigned long long)s) { T = b; break; }
            }
            off += (unsigned)s;
        }
        pl.thresh[r] = T;
    }
    for (int r = 0; r < pl.world; ++r) {
        if (pl.surplus[r] < 0) acc += (unsigned)(-pl.surplus[r]);
        pl.impend[r] = acc;          // importer r takes export positions [impend[r-1], impend[r]); exporters take none
    }
    pl.info[0] = off;
}

// class of every offspring for the first scan: its home rank if it may be exported, else none (255)
__global__ void __launch_bounds__(256)
pl_eligible_kernel(PlPlan pl, const int32_t *__restrict__ anc)
{
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < pl.N; m += (int64_t)gridDim.x * blockDim.x) {
        const int r = pl.home[m];
        const bool el = pl.surplus[r] > 0 && min(pl.ocnt[anc[m]], PL_HB - 1) >= pl.thresh[r];
        pl.cls[m] = el ? (uint8_t)r : (uint8_t)255;
    }
}

// ---- exclusive prefix per class (classes 0 .. PL_MAXR-1, 255 = none) over cls[0..N), three small kernels ----------
__global__ void __launch_bounds__(PL_T)
pl_mcscan_count(const uint8_t *__restrict__ cls, int64_t N, unsigned *tilecnt)
{
    __shared__ unsigned sc[PL_MAXR];
    if (threadIdx.x < PL_MAXR) sc[threadIdx.x] = 0u;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * PL_TILE;
    for (int j = threadIdx.x; j < PL_TILE; j += PL_T) {
        const int64_t i = base + j;
        if (i < N) { const unsigned c = cls[i]; if (c < PL_MAXR) atomicAdd(&sc[c], 1u); }
    }
    __syncthreads();
    if (threadIdx.x < PL_MAXR) tilecnt[(size_t)blockIdx.x * PL_MAXR + threadIdx.x] = sc[threadIdx.x];
}

// one block of 1024 threads: per class, exclusive prefix of the tile counts (in place)
__global__ void __launch_bounds__(1024)
pl_mcscan_prefix(unsigned *tilecnt, int ntiles)
{
    __shared__ unsigned ws[32];
    __shared__ unsigned carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int c = 0; c < PL_MAXR; ++c) {
        if (threadIdx.x == 0) carry = 0u;
        __syncthreads();
        for (int base = 0; base < ntiles; base += 1024) {
            const int t = base + threadIdx.x;
            const unsigned v = (t < ntiles) ? tilecnt[(size_t)t * PL_MAXR + c] : 0u;
            unsigned inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += u;
            }
            if (lane == 31) ws[wid] = inc;
            __syncthreads();
            unsigned off = carry;
            for (int k = 0; k < wid; ++k) off += ws[k];
            if (t < ntiles) tilecnt[(size_t)t * PL_MAXR + c] = off + inc - v;
            __syncthreads();
            if (threadIdx.x == 1023) carry = off + inc;
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(PL_T)
pl_mcscan_apply(const uint8_t *__restrict__ cls, int64_t N, const unsigned *__restrict__ tilecnt, int32_t *prefix)
{
    __shared__ unsigned wtot[PL_T / 32][PL_MAXR];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int PER = PL_TILE / PL_T;                       // consecutive elements per thread
    const int64_t base = (int64_t)blockIdx.x * PL_TILE + (int64_t)threadIdx.x * PER;
    unsigned c[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) c[j] = (base + j < N) ? cls[base + j] : 255u;
    unsigned mine[PER];                                       // exclusive rank of my elements within the warp, per their class
    for (int k = 0; k < PL_MAXR; ++k) {
        unsigned n = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) n += (c[j] == (unsigned)k);
        unsigned inc = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) wtot[wid][k] = inc;
        unsigned run = inc - n;
#pragma unroll
        for (int j = 0; j < PER; ++j)
            if (c[j] == (unsigned)k) mine[j] = run++;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        if (base + j < N && c[j] < PL_MAXR) {
            unsigned off = tilecnt[(size_t)blockIdx.x * PL_MAXR + c[j]];
            for (int w = 0; w < wid; ++w) off += wtot[w][c[j]];
            prefix[base + j] = (int32_t)(off + mine[j]);
        }
    }
}

// destination of every offspring
__global__ void __launch_bounds__(256)
pl_dest_kernel(PlPlan pl)
{
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < pl.N; m += (int64_t)gridDim.x * blockDim.x) {
        const int r = pl.home[m];
        int d = r;
        if (pl.cls[m] != 255 && pl.prefix[m] < pl.surplus[r]) {
            const unsigned e = pl.expoff[r] + (unsigned)pl.prefix[m];      // position in the export sequence
            d = 0;
            while (d < pl.world - 1 && e >= pl.impend[d]) ++d;             // first rank whose interval ends beyond e
        }
        pl.dest[m] = (uint8_t)d;
    }
}

// place_new, this rank's local lists, and the marks for ancestors that other ranks will read out of this store
__global__ void __launch_bounds__(256)
pl_finish_kernel(PlPlan pl, const int32_t *__restrict__ anc, const int32_t *__restrict__ place, int32_t *place_new,
                 int32_t *logi_new, int32_t *src, int32_t *ancl, const int32_t *__restrict__ slot, int32_t *used)
{
    unsigned long long pulled = 0;
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < pl.N; m += (int64_t)gridDim.x * blockDim.x) {
        const int d = pl.dest[m];
        const int a = anc[m];
        const int q = place[a];
        const int pn = (int)((int64_t)d * pl.P + pl.prefix[m]);
        place_new[m] = pn;
        if (d == pl.rank) {
            const int i = pl.prefix[m];
            logi_new[i] = (int32_t)m;
            src[i] = q;
            ancl[i] = a;
            // a map pulled over NVLink: first offspring of a lineage on this rank whose ancestor lives elsewhere
            if (pl.home[m] != d && (m == 0 || anc[m - 1] != a || pl.dest[m - 1] != d)) ++pulled;
        } else if (pl.home[m] == pl.rank) {
            used[slot[q - (int)((int64_t)pl.rank * pl.P)]] = 1;     // read by rank d during its gather: keep the slot this round
        }
    }
    for (int o = 16; o > 0; o >>= 1) pulled += __shfl_xor_sync(0xffffffffu, pulled, o);
    if ((threadIdx.x & 31) == 0 && pulled) atomicAdd(&pl.info[1], pulled);
}

// ---- the gather of a placed shard -------------------------------------------------------------------------------
// src[i]  : physical source (rank * P + local index) of new local particle i;  ancl[i] : its logical ancestor
// (non-decreasing in i).  first = first local offspring of its lineage.  A first offspring with a LOCAL source inherits
// the source's map slot; every other particle needs a free slot: extra[i] = 1.
// With `defer` (deferred map copies, fs2_update_ws.cuh DEFER) the local offspring of a lineage -- consecutive particles --
// are cut into groups of a leader and up to 7 followers, as fs2_search_mark_kernel does on one GPU: rank j in the local
// lineage with j % 8 == 0 is a leader (it owns a map after this gather: inherited, pulled or copied), the others are
// followers (nfol[i] = -1: a free slot now, the copy from the next update kernel).  nfol[i] >= 0: followers of leader i.
__global__ void __launch_bounds__(256)
pl_mark_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ ancl, int64_t P, int rank,
               const int32_t *__restrict__ slot, int32_t *used, int32_t *extra, int defer, int32_t *nfol, int32_t *leaders,
               int32_t *dctl)
{
    const int64_t lo = (int64_t)rank * P;
    const int lane = threadIdx.x & 31;
    if (defer && blockIdx.x == 0 && threadIdx.x == 0) dctl[0] = 1;
    const int64_t Pr = (P + 31) & ~(int64_t)31;              // whole warps stay in the loop (the vote below)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < Pr; i += (int64_t)gridDim.x * blockDim.x) {
        bool leader = false;
        if (i < P) {
            const int q = src[i];
            const int a = ancl[i];
            const bool first = (i == 0) || (ancl[i - 1] != a);
            const bool local = q >= lo && q < lo + P;
            const bool inherit = first && local;
            extra[i] = inherit ? 0 : 1;
            if (inherit) used[slot[q - lo]] = 1;
            if (defer) {
                int64_t b0 = i;
                if (!first) {                                 // first local offspring of the lineage (ancl is non-decreasing)
                    int64_t l0 = 0, l1 = i;
                    while (l0 < l1) { const int64_t mid = (l0 + l1) >> 1; if (ancl[mid] < a) l0 = mid + 1; else l1 = mid; }
                    b0 = l0;
                }
                leader = (((i - b0) & 7) == 0);
                int nf = -1;
                if (leader) {
                    nf = 0;
#pragma unroll
                    for (int k = 1; k <= 7; ++k) if (i + k < P && ancl[i + k] == a) ++nf;
                }
                nfol[i] = nf;
            }
        }
        if (defer) {
            const unsigned lb = __ballot_sync(0xffffffffu, leader);
            int base = 0;
            if (lane == 0 && lb) base = atomicAdd(&dctl[1], __popc(lb));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (leader) leaders[base + __popc(lb & ((1u << lane) - 1u))] = (int32_t)i;
        }
    }
}

__global__ void __launch_bounds__(256)
pl_pose_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ extra, int64_t P, const Fs2Peers *peers,
               double *x2, double *y2, double *yaw2, double *w2, int32_t *count2, int32_t *slot2)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
        const int q = src[i];
        const int r = (int)(q / P);
        const int li = (int)(q - (int64_t)r * P);
        x2[i] = peers->x[r][li]; y2[i] = peers->y[r][li]; yaw2[i] = peers->yaw[r][li]; w2[i] = peers->w[r][li];
        count2[i] = peers->count[r][li];
        if (!extra[i]) slot2[i] = peers->slot[r][li];            // inherits (local source)
    }
}

// one warp per particle that needs a free slot.  phase 0: sources that are valid right now -- the ancestor's slot,
// local or in another GPU's store (peer loads over NVLink) -- for the first offspring of a lineage here and for every
// offspring of a LOCAL ancestor.  phase 1: further offspring of a REMOTE ancestor copy the map their lineage's first
// offspring pulled in phase 0, locally.
__global__ void __launch_bounds__(256)
pl_copy_kernel(int phase, const int32_t *__restrict__ tasks, const int32_t *__restrict__ freeslot, const int32_t *ncopies,
               const int32_t *__restrict__ src, const int32_t *__restrict__ ancl, int64_t P, int rank, const Fs2Peers *peers,
               double *lm, int lcap, int32_t *slot2, const int32_t *__restrict__ count2, const int32_t *dctl,
               const int32_t *__restrict__ nfol)
{
    const bool dfr = dctl && ((volatile const int32_t *)dctl)[0] != 0;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int n = min(ncopies[0], ncopies[1]);
    const size_t stride = 6 * (size_t)lcap;
    const int64_t lo = (int64_t)rank * P;
    for (int64_t t = warp; t < n; t += nwarps) {
        const int i = tasks[t];
        if (dfr && nfol[i] < 0) {                      // a follower: its slot now, its map from the next update kernel
            if (phase == 0 && lane == 0) slot2[i] = freeslot[t];
            continue;
        }
        const int q = src[i];
        const bool local = q >= lo && q < lo + P;
        const bool first = (i == 0) || (ancl[i - 1] != ancl[i]);
        const bool now = local || first;
        if ((phase == 0) != now) continue;
        const int dst_slot = freeslot[t];
        const int4 *s4;
        if (now) {
            const int r = (int)(q / P);
            const int li = (int)(q - (int64_t)r * P);
            s4 = reinterpret_cast<const int4 *>(peers->lm[r] + (size_t)peers->slot[r][li] * stride);
        } else {
            // first local offspring of my lineage: lowest index with the same logical ancestor
            const int a = ancl[i];
            int b0 = 0, b1 = i;
            while (b0 < b1) { const int mid = (b0 + b1) >> 1; if (ancl[mid] < a) b0 = mid + 1; else b1 = mid; }
            s4 = reinterpret_cast<const int4 *>(lm + (size_t)slot2[b0] * stride);
        }
        const int ng = count2[i] * 3;                  // 16-byte granules
        int4 *d4 = reinterpret_cast<int4 *>(lm + (size_t)dst_slot * stride);
        int g = lane;
        for (; g + 96 < ng; g += 128) {
            int4 v0 = __ldcs(s4 + g), v1 = __ldcs(s4 + g + 32), v2 = __ldcs(s4 + g + 64), v3 = __ldcs(s4 + g + 96);
            __stcs(d4 + g, v0); __stcs(d4 + g + 32, v1); __stcs(d4 + g + 64, v2); __stcs(d4 + g + 96, v3);
        }
        for (; g < ng; g += 32) __stcs(d4 + g, __ldcs(s4 + g));
        if (lane == 0) slot2[i] = dst_slot;
    }
}
