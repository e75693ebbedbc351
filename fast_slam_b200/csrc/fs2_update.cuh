// fs2_update.cuh -- the fused motion + association + EKF + weighting kernel (rows A2-A6 of SURVEY.md 8a).
//
// One warp owns one particle at a time (persistent grid, warps stride over particles).
//
//   phase 1  stream the particle's map ONCE: 32 landmarks per chunk (1536 contiguous bytes) are staged
//            into a per-warp shared-memory ring by the TMA (cp.async.bulk, one instruction per chunk issued
//            by lane 0, completion on a per-stage mbarrier); the ring runs ahead across particle
//            boundaries, so the next particle's first chunks load while this one's EKF math runs.  Lane j
//            reads landmark j back with three conflict-free LDS.128, builds its conservative fp32 gate box
//            (fs2_box) and looks the box centre up in a small CELL TABLE of the step's observations
//            (built once per step on the host, Fs2ObsBatch::tab1/tab2): the table returns the bit mask
//            of observations that can possibly lie inside a box of that size centred in that cell, so
//            the per-landmark cost does not grow with the number of observations.  The few candidate
//            bits are then checked against the actual box.
//   phase 2  (landmark, observation) pairs that pass are queued in shared memory and re-tested with the
//            exact fp64 gate (fs2_gate_test, bit-identical to the oracle); the up-to-4 lowest matching
//            landmark indices of every observation are kept in a sorted shared-memory list
//            (atomicMin cascade).
//   phase 3  observations are applied in list order (quirk Q7) by speculation: every remaining
//            observation (lane k = observation k) picks its first match (pre-step list patched with the
//            landmarks touched so far this step), all EKF updates / new landmarks are computed in
//            parallel, then dependencies on EARLIER observations of the same round are detected:
//            same landmark (match.any), or an earlier observation's POST-state landmark with a lower
//            index now stops this observation's scan (table lookup of the post box + exact test).
//            Everything before the first dependency is committed; the rest is re-speculated against the
//            updated state.  Independent observations (the normal case) finish in one round.
//   fallback observation-by-observation exact scan of the map in global memory -- literally the
//            reference's loop (landmark_utils.py:103-117) -- used when an observation's match list
//            overflowed and was exhausted, or when FS2_FLAG_FORCE_SEQUENTIAL is set (cross-check).
#pragma once
#include "fs2_math.cuh"

#define FS2_WPB 8          // warps per block
#ifndef FS2_NST
#define FS2_NST 2          // TMA ring stages per warp (2 x 3 KB; a third stage costs more L1 than it hides latency)
#endif
#define FS2_CHUNK 64       // landmarks per stage: two per lane, processed as two independent instruction streams
#define FS2_CHUNK_BYTES (FS2_CHUNK * 48)
#ifndef FS2_QCAP
#define FS2_QCAP 160       // candidate queue entries per warp (drained when fewer than 96 are free: 3 x 32 can arrive per round)
#endif
#define FS2_NONE 0x7fffffff
#define FS2_FULL 0xffffffffu
#ifndef FS2_G1
#define FS2_G1 48          // fine observation cell table: G1 x G1 cells (+ a border ring), margin e1
#endif
#define FS2_G2 8           // coarse table for big boxes (fresh landmarks: gate * sqrt(0.1) = 2.53 m), margin e2
#define FS2_G1P (FS2_G1 + 2)
#define FS2_G2P (FS2_G2 + 2)

struct Fs2State {
    double *x, *y, *yaw, *w;
    int32_t *count;
    int32_t *slot;   // map slot of each particle (copy-on-resample indirection)
    double *lm;      // [slots][lcap][6]
    int32_t *status;
    int64_t P;
    int32_t lcap;
    int32_t pad;
};

struct Fs2ObsBatch {  // <= 32 observations of one step, host-prepared (robot frame: fast_slam_2.py:100-103)
    double zd[32], za[32], ox[32], oy[32];
    double sza[32], cza[32];          // sin / cos of the bearing za (host libm)
    float oxf[32], oyf[32];           // padded with +inf beyond M
    unsigned tab1[FS2_G1P * FS2_G1P]; // observations within e1 (Chebyshev) of each fine cell
    unsigned tab2[FS2_G2P * FS2_G2P]; // ... within e2 of each coarse cell
    float gx0, gy0;                   // lower corner of the observations' bounding square
    float inv_s1, inv_s2;             // 1 / cell size
    float e1, e2;                     // a box with rx, ry <= e may use the table of that level
    // the same two levels without building the box (warp-specialised kernel's in-loop screen): a SAFE landmark
    // (fs2_box's conditions) with max(c00, c11) <= amaxN and |x|, |y| < xymax has box half-widths <= eN;
    // a safe landmark with max(c00, c11) <= amax2 at or beyond xymax cannot hold any observation at all
    float amax1, amax2, xymax;
    float cx1, cy1, cx2, cy2;         // table column = floor(clamp(fma(x, inv_sN, cxN), 0, GN + 1)), likewise rows
    float slack;                      // 2.4e-7 * max(|ox|,|oy|) + tiny
    int32_t M;
    int32_t k0;                       // index of the batch's first observation in the step's list
    unsigned all_mask;                // bits of the real observations
};

struct Fs2UpdateArgs {
    double r00, r01, r10, r11;  // MEASUREMENT_NOISE (config.py:15)
    double gate;                // MAXIMUM_LANDMARK_DISTANCE (config.py:18)
    double rotation, translation;
    const double *noise;        // double[P] or nullptr
    int32_t *assoc;             // int32[Mtotal][P] or nullptr
    float gate_f;
    int32_t do_motion;
    int32_t force_seq;
    int32_t pad;
    // deferred map copies of the last resample (warp-specialised kernel only, see fs2_update_ws.cuh)
    const int32_t *dctl;        // device: [0] copies are pending -- this launch runs the leaders list; [1] its length
    const int32_t *leaders;     // particles that own their map already: first offspring and every 8th sibling
    const int32_t *nfol;        // per leader: the next nfol particles are its followers (<= FS2_SIBMAX)
};
#define FS2_SIBMAX 7

struct Fs2UpdateSmem {
    double ox[32], oy[32], zd[32], za[32];
    alignas(8) float2 of[33];                         // (oxf, oyf); [32] = (+inf, +inf) sentinel
    unsigned tab1[FS2_G1P * FS2_G1P];
    unsigned tab2[FS2_G2P * FS2_G2P];
    alignas(128) unsigned char ring[FS2_WPB][FS2_NST][FS2_CHUNK_BYTES];
    alignas(8) unsigned long long bar[FS2_WPB][FS2_NST];   // one mbarrier per ring stage
    int qidx[FS2_WPB][FS2_QCAP];
    unsigned qmask[FS2_WPB][FS2_QCAP];
    alignas(16) int4 ml[FS2_WPB][32];                  // per observation: its <= 4 lowest exact matches on the pre-step map
    unsigned ovf[FS2_WPB];                 // observations with more than 4
    unsigned conf[FS2_WPB];                // observations that depend on an earlier one of the round
    int bound[FS2_WPB][32];                // association index of observation k (FS2_NONE: appends)
    alignas(16) Fs2Lm tlm[FS2_WPB][32];                // landmarks touched so far this step (current state)
    alignas(16) float4 tbox[FS2_WPB][32];
    int tidx[FS2_WPB][32];
};

__device__ __forceinline__ void fs2_mbar_init(unsigned long long *bar, unsigned count)
{
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void fs2_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }

// one elected lane: arm the stage's barrier with the byte count and start the bulk copy global -> shared
__device__ __forceinline__ void fs2_tma_load(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(d), "l"(src), "r"(bytes), "r"(b) : "memory");
}

__device__ __forceinline__ void fs2_mbar_wait(unsigned bar_saddr, unsigned parity)
{
    unsigned ok;
    do {
        // the suspend-time hint lets the hardware park the warp until the phase completes instead of polling
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(bar_saddr), "r"(parity), "r"(0x989680u) : "memory");
    } while (!ok);
}

__device__ __forceinline__ Fs2Lm fs2_load_lm(const double *lm, int i)
{
    const double2 *p = reinterpret_cast<const double2 *>(lm + 6 * (size_t)i);
    double2 a = p[0], b = p[1], c = p[2];
    Fs2Lm l;
    l.x = a.x; l.y = a.y; l.c00 = b.x; l.c01 = b.y; l.c10 = c.x; l.c11 = c.y;
    return l;
}
__device__ __forceinline__ void fs2_store_lm(double *lm, int i, const Fs2Lm &l)
{
    double2 *p = reinterpret_cast<double2 *>(lm + 6 * (size_t)i);
    p[0] = make_double2(l.x, l.y);
    p[1] = make_double2(l.c00, l.c01);
    p[2] = make_double2(l.c10, l.c11);
}

// exact gate incl. the "singular covariance raises" rule: true = the reference's loop stops here
__device__ __forceinline__ bool fs2_stops_here(const Fs2Lm &l, double ox, double oy, double gate)
{
    Fs2Gate g = fs2_gate_prepare(l.c00, l.c01, l.c10, l.c11);
    return g.singular || fs2_gate_test(g, l.x, l.y, ox, oy, gate);
}

// index of the table cell that holds (x, y): column/row 0 and G + 1 are the half-infinite border ring.  The clamp
// runs in fp32 before the conversion (NaN lands in cell 0; whatever the mask there, the exact test rejects NaN)
template <int G>
__device__ __forceinline__ int fs2_cell(float x, float y, float inv_s, float offx, float offy)
{
    const float fx = fminf(fmaxf(fmaf(x, inv_s, offx), 0.f), (float)(G + 1));
    const float fy = fminf(fmaxf(fmaf(y, inv_s, offy), 0.f), (float)(G + 1));
    return __float2int_rd(fy) * (G + 2) + __float2int_rd(fx);
}

// observations that can lie inside box b (superset), from the cell tables
template <class SM>
__device__ __forceinline__ unsigned fs2_candidates(const SM &sm, const Fs2ObsBatch &ob, const Fs2Box &b)
{
    const float r = fmaxf(b.rx, b.ry);
    if (r <= ob.e1) return sm.tab1[fs2_cell<FS2_G1>(b.mx, b.my, ob.inv_s1, ob.cx1, ob.cy1)];
    if (r <= ob.e2) return sm.tab2[fs2_cell<FS2_G2>(b.mx, b.my, ob.inv_s2, ob.cx2, ob.cy2)];
    return (r >= 0.f) ? ob.all_mask : 0u;   // r < 0 marks "no landmark"; NaN radius cannot happen (fs2_box)
}

// keep the candidate bits whose observation really lies inside the box.  The first two candidates are tested
// in straight-line code (the usual count is 0-2; an empty slot reads the +inf sentinel of[32]); whatever is left
// is returned in *rest for fs2_box_filter_rest, so that callers with several landmarks per lane can interleave
// the loop-free parts and run ONE leftover loop.
template <class SM>
__device__ __forceinline__ unsigned fs2_box_filter2(const SM &sm, const Fs2Box &b, unsigned cand, unsigned *rest)
{
    const unsigned low0 = cand & (0u - cand);
    const unsigned c1 = cand ^ low0;
    const unsigned low1 = c1 & (0u - c1);
    *rest = c1 ^ low1;
    const float2 o0 = sm.of[__clz(__brev(cand))];     // index 32 when empty
    const float2 o1 = sm.of[__clz(__brev(c1))];
    unsigned keep = 0;
    if (fabsf(o0.x - b.mx) < b.rx && fabsf(o0.y - b.my) < b.ry) keep = low0;
    if (fabsf(o1.x - b.mx) < b.rx && fabsf(o1.y - b.my) < b.ry) keep |= low1;
    return keep;
}

template <class SM>
__device__ __forceinline__ unsigned fs2_box_filter_rest(const SM &sm, const Fs2Box &b, unsigned rest)
{
    unsigned keep = 0;
    while (rest) {
        const int k = __ffs(rest) - 1;
        rest &= rest - 1;
        const float2 o = sm.of[k];
        if (fabsf(o.x - b.mx) < b.rx && fabsf(o.y - b.my) < b.ry) keep |= (1u << k);
    }
    return keep;
}

template <class SM>
__device__ __forceinline__ unsigned fs2_box_filter(const SM &sm, const Fs2Box &b, unsigned cand)
{
    unsigned rest;
    unsigned keep = fs2_box_filter2(sm, b, cand, &rest);
    if (rest) keep |= fs2_box_filter_rest(sm, b, rest);
    return keep;
}

// sorted insert of landmark index idx into observation k's list of lowest matches (ml: int4[32] in shared memory)
__device__ __forceinline__ void fs2_ml_insert(int4 *ml, unsigned *ovf, int k, int idx)
{
    int *slot = reinterpret_cast<int *>(ml + k);
    int v = idx;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        int old = atomicMin(slot + s, v);
        v = max(old, v);
        if (v == FS2_NONE) return;
    }
    atomicOr(ovf, 1u << k);
}

// phase 2: exact re-test of the queued candidates
__device__ __forceinline__ void fs2_drain(const int *qidx, const unsigned *qmask, int4 *ml, unsigned *ovf,
                                          const double *s_ox, const double *s_oy, int lane, const double *lm, int qn,
                                          double gate)
{
    for (int e = lane; e < qn; e += 32) {
        const int idx = qidx[e];
        unsigned m = qmask[e];
        const Fs2Lm l = fs2_load_lm(lm, idx);
        const Fs2Gate g = fs2_gate_prepare(l.c00, l.c01, l.c10, l.c11);
        while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            if (g.singular || fs2_gate_test(g, l.x, l.y, s_ox[k], s_oy[k], gate)) fs2_ml_insert(ml, ovf, k, idx);
        }
    }
    __syncwarp();
}

#ifndef FS2_MIN_BLOCKS
#define FS2_MIN_BLOCKS 2
#endif

__global__ void __launch_bounds__(FS2_WPB * 32, FS2_MIN_BLOCKS)
fs2_update_kernel(const Fs2State st, const __grid_constant__ Fs2ObsBatch ob, const Fs2UpdateArgs ua)
{
    extern __shared__ __align__(128) unsigned char fs2_smem_raw[];
    Fs2UpdateSmem &sm = *reinterpret_cast<Fs2UpdateSmem *>(fs2_smem_raw);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    if (threadIdx.x < 32) {
        sm.ox[lane] = ob.ox[lane]; sm.oy[lane] = ob.oy[lane];
        sm.zd[lane] = ob.zd[lane]; sm.za[lane] = ob.za[lane];
        sm.of[lane] = make_float2(ob.oxf[lane], ob.oyf[lane]);
        if (lane == 0) sm.of[32] = make_float2(__int_as_float(0x7f800000), __int_as_float(0x7f800000));
    }
    for (int i = threadIdx.x; i < FS2_G1P * FS2_G1P; i += blockDim.x) sm.tab1[i] = ob.tab1[i];
    for (int i = threadIdx.x; i < FS2_G2P * FS2_G2P; i += blockDim.x) sm.tab2[i] = ob.tab2[i];
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < FS2_NST; ++s) fs2_mbar_init(&sm.bar[wib][s], 1);
        fs2_fence_mbar_init();
    }
    __syncthreads();
    const int M = ob.M;
    const int lcap = st.lcap;
    const int64_t nwarps = (int64_t)gridDim.x * FS2_WPB;
    const bool is_obs = lane < M;
    const double zd = sm.zd[lane], za = sm.za[lane];
    const double oxd = sm.ox[lane], oyd = sm.oy[lane];
    const float2 myof = sm.of[lane];
    const bool streaming = (ua.force_seq == 0) && (M > 0);
    unsigned char *ring = &sm.ring[wib][0][0];
    unsigned long long *bars = &sm.bar[wib][0];

    // ring bookkeeping: chunks are numbered consecutively over the warp's whole life; chunk g lives in
    // stage g % NST and completes phase (g / NST) & 1 of that stage's barrier
    unsigned gc = 0;            // next chunk to consume
    unsigned gp = 0;            // next chunk to issue
    const unsigned bar_s0 = (unsigned)__cvta_generic_to_shared(bars);
    int64_t p = (int64_t)blockIdx.x * FS2_WPB + wib;
    // the particle header (pose, weight, noise, map size and slot) is fetched one particle ahead
    int cnt_cur = 0, slot_cur = 0;
    double px_n = 0.0, py_n = 0.0, pyaw_n = 0.0, pw_n = 0.0, nz_n = 0.0;
    if (p < st.P) {
        cnt_cur = st.count[p]; slot_cur = st.slot[p];
        px_n = st.x[p]; py_n = st.y[p]; pyaw_n = st.yaw[p]; pw_n = st.w[p];
        if (ua.do_motion) nz_n = ua.noise[p];
    }
    // issue chunk c of a map (lane 0 only)
    auto issue = [&](const double *map, int cnt_map, int c) {
        const unsigned bytes = (unsigned)min(FS2_CHUNK, cnt_map - FS2_CHUNK * c) * 48u;
        fs2_tma_load(ring + (gp % FS2_NST) * FS2_CHUNK_BYTES,
                     reinterpret_cast<const unsigned char *>(map) + (size_t)c * FS2_CHUNK_BYTES, bytes, bars + (gp % FS2_NST));
    };
    if (streaming && p < st.P) {          // prime the ring with the first particle's first chunks
        const double *map = st.lm + (size_t)slot_cur * 6 * (size_t)lcap;
        const int nch = (cnt_cur + FS2_CHUNK - 1) / FS2_CHUNK;
        const int pre = min(FS2_NST - 1, nch);
        for (int c = 0; c < pre; ++c) { if (lane == 0) issue(map, cnt_cur, c); ++gp; }
    }

    for (; p < st.P; p += nwarps) {
        double px = px_n, py = py_n, pyaw = pyaw_n, pw = pw_n;
        const double nz = nz_n;
        int cnt = cnt_cur;
        double *lm = st.lm + (size_t)slot_cur * 6 * (size_t)lcap;
        const int64_t pn = p + nwarps;
        int cnt_next = 0, slot_next = 0;
        if (pn < st.P) {
            cnt_next = st.count[pn]; slot_next = st.slot[pn];
            px_n = st.x[pn]; py_n = st.y[pn]; pyaw_n = st.yaw[pn]; pw_n = st.w[pn];
            if (ua.do_motion) nz_n = ua.noise[pn];
        }
        int stat = 0;
        if (ua.do_motion) fs2_move(px, py, pyaw, ua.rotation, ua.translation, nz);

        int ks = 0;            // first observation not yet applied
        int nt = 0;            // touched landmarks in sm.tlm / tidx / tbox
        bool seq = (ua.force_seq != 0);
        int my_assoc = -3;     // result of observation `lane`

        if (!seq && M > 0) {
            // ---------------- phase 1: stream + screen ----------------
            sm.ml[wib][lane] = make_int4(FS2_NONE, FS2_NONE, FS2_NONE, FS2_NONE);
            if (lane == 0) sm.ovf[wib] = 0u;
            const int nchunks = (cnt + FS2_CHUNK - 1) / FS2_CHUNK;
            int issued = min(FS2_NST - 1, nchunks);     // chunks of this particle already in flight
            int qn = 0;
            __syncwarp();
            for (int c = 0; c < nchunks; ++c) {
                if (issued < nchunks) {                 // keep NST-1 chunks in flight
                    if (lane == 0) issue(lm, cnt, issued);
                    ++gp; ++issued;
                }
                const unsigned stage = gc % FS2_NST;
                fs2_mbar_wait(bar_s0 + 8u * stage, (gc / FS2_NST) & 1u);
                ++gc;
                // two landmarks per lane, written as two independent streams for instruction-level parallelism
                const int iA = c * FS2_CHUNK + lane, iB = iA + 32;
                const double2 *srcA = reinterpret_cast<const double2 *>(ring + stage * FS2_CHUNK_BYTES + 48 * lane);
                const double2 *srcB = srcA + 96;    // 32 landmarks * 48 B / 16 B
                Fs2Box bA, bB;
                bA.mx = bA.my = 0.f; bA.rx = bA.ry = -1.f;
                bB = bA;
                if (iA < cnt) {
                    const double2 a0 = srcA[0], a1 = srcA[1], a2 = srcA[2];
                    bA = fs2_box(a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, ua.gate_f, ob.slack);
                }
                if (iB < cnt) {
                    const double2 b0 = srcB[0], b1 = srcB[1], b2 = srcB[2];
                    bB = fs2_box(b0.x, b0.y, b1.x, b1.y, b2.x, b2.y, ua.gate_f, ob.slack);
                }
                const unsigned candA = fs2_candidates(sm, ob, bA), candB = fs2_candidates(sm, ob, bB);
                const unsigned maskA = fs2_box_filter(sm, bA, candA);
                const unsigned maskB = fs2_box_filter(sm, bB, candB);
                const unsigned hasA = __ballot_sync(FS2_FULL, maskA != 0);
                const unsigned hasB = __ballot_sync(FS2_FULL, maskB != 0);
                if (hasA | hasB) {
                    // queue order must stay ascending in landmark index: all of A (iA < iB) first
                    if (maskA) {
                        const int pos = qn + __popc(hasA & lt_mask);
                        sm.qidx[wib][pos] = iA;
                        sm.qmask[wib][pos] = maskA;
                    }
                    qn += __popc(hasA);
                    if (maskB) {
                        const int pos = qn + __popc(hasB & lt_mask);
                        sm.qidx[wib][pos] = iB;
                        sm.qmask[wib][pos] = maskB;
                    }
                    qn += __popc(hasB);
                    if (qn > FS2_QCAP - 96) {
                        __syncwarp();
                        fs2_drain(sm.qidx[wib], sm.qmask[wib], sm.ml[wib], &sm.ovf[wib], sm.ox, sm.oy, lane, lm, qn, ua.gate);
                        qn = 0;
                    }
                }
                __syncwarp();  // every lane is done with this stage before lane 0 hands it back to the TMA
            }
            // the ring is empty: start on the next particle's map while this one's math runs
            if (pn < st.P) {
                const double *mapn = st.lm + (size_t)slot_next * 6 * (size_t)lcap;
                const int nchn = (cnt_next + FS2_CHUNK - 1) / FS2_CHUNK;
                const int pre = min(FS2_NST - 1, nchn);
                for (int c = 0; c < pre; ++c) { if (lane == 0) issue(mapn, cnt_next, c); ++gp; }
            }
            // ---------------- phase 2: exact re-test of what is left in the queue ----------------
            fs2_drain(sm.qidx[wib], sm.qmask[wib], sm.ml[wib], &sm.ovf[wib], sm.ox, sm.oy, lane, lm, qn, ua.gate);
            const int4 ml = sm.ml[wib][lane];
            const bool ml_overflow = (sm.ovf[wib] >> lane) & 1u;

            // ---------------- phase 3: speculative, order-preserving application ----------------
            while (ks < M && !seq) {
                const bool active = is_obs && lane >= ks;
                // (a) association of every remaining observation against the current state
                int a_un = FS2_NONE;        // first pre-step match that has not been touched this step
                bool exhausted = false;
                if (active) {
                    if (nt == 0) {
                        a_un = ml.x;
                    } else {
                        const int cands[4] = {ml.x, ml.y, ml.z, ml.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int cand = cands[i];
                            if (a_un == FS2_NONE && cand != FS2_NONE) {
                                bool touched = false;
                                for (int t = 0; t < nt; ++t) touched |= (sm.tidx[wib][t] == cand);
                                if (!touched) a_un = cand;
                            }
                        }
                        if (a_un == FS2_NONE && ml_overflow) exhausted = true;  // more pre-step matches exist, unknown
                    }
                }
                int a_t = FS2_NONE;         // lowest touched landmark whose CURRENT state stops the scan
                int a_t_pos = -1;
                if (active) {
                    for (int t = 0; t < nt; ++t) {
                        const float4 tb = sm.tbox[wib][t];
                        const int ti = sm.tidx[wib][t];
                        if (ti < a_t && fabsf(myof.x - tb.x) < tb.z && fabsf(myof.y - tb.y) < tb.w) {
                            if (fs2_stops_here(sm.tlm[wib][t], oxd, oyd, ua.gate)) { a_t = ti; a_t_pos = t; }
                        }
                    }
                }
                // an exhausted list only matters if no touched landmark below it decides; be conservative
                if (__any_sync(FS2_FULL, exhausted)) { seq = true; break; }
                const bool from_t = a_t < a_un;
                const int a = from_t ? a_t : a_un;          // FS2_NONE: no landmark stops the scan -> append
                const bool matched = active && a != FS2_NONE;
                const unsigned unm = __ballot_sync(FS2_FULL, active && !matched);
                const int app_rank = __popc(unm & lt_mask);
                // (b) speculative result
                Fs2Lm post;
                post.x = post.y = post.c00 = post.c01 = post.c10 = post.c11 = 0.0;
                double like = 1.0;
                int st_k = 0;
                int widx = FS2_NONE;        // landmark index this observation writes
                int res = -3;
                if (matched) {
                    const Fs2Lm in = from_t ? sm.tlm[wib][a_t_pos] : fs2_load_lm(lm, a);
                    const double det = __dadd_rn(__dmul_rn(in.c00, in.c11), -__dmul_rn(in.c01, in.c10));
                    if (det == 0.0) {       // np.linalg.inv raises inside associate: update skipped
                        st_k = 1;
                        res = -2;
                    } else {
                        st_k = fs2_ekf(px, py, pyaw, zd, za, ua.r00, ua.r01, ua.r10, ua.r11, in, &post, &like);
                        if (st_k == 2) res = -2; else { res = a; widx = a; }
                    }
                } else if (active) {
                    res = -1;
                    if (cnt + app_rank < lcap) {
                        post = fs2_new_landmark(px, py, pyaw, zd, za);
                        widx = cnt + app_rank;
                    } else {
                        st_k = 8;           // FS2_ST_MAP_FULL
                    }
                }
                Fs2Box pb;
                pb.mx = 0.f; pb.my = 0.f; pb.rx = -1.f; pb.ry = -1.f;
                if (widx != FS2_NONE) pb = fs2_box(post.x, post.y, post.c00, post.c01, post.c10, post.c11, ua.gate_f, ob.slack);
                // (c) does an observation depend on an EARLIER, not yet committed one?
                //   same landmark: the later one must see the earlier one's result
                const int key = matched ? a : (widx != FS2_NONE ? widx : (0x40000000 | lane));
                const unsigned act_mask = __ballot_sync(FS2_FULL, active);
                const unsigned same = __match_any_sync(FS2_FULL, active ? key : (0x50000000 | lane));
                sm.bound[wib][lane] = matched ? a : FS2_NONE;
                if (lane == 0) sm.conf[wib] = 0u;
                __syncwarp();
                if (matched && (same & lt_mask & act_mask)) atomicOr(&sm.conf[wib], 1u << lane);
                //   my post-state landmark, at a lower index than a LATER observation's choice, stops its scan
                if (widx != FS2_NONE) {
                    unsigned later = fs2_candidates(sm, ob, pb) & act_mask & ~(lt_mask | (1u << lane));
                    later = fs2_box_filter(sm, pb, later);
                    while (later) {
                        const int k2 = __ffs(later) - 1;
                        later &= later - 1;
                        if (widx < sm.bound[wib][k2] && fs2_stops_here(post, sm.ox[k2], sm.oy[k2], ua.gate))
                            atomicOr(&sm.conf[wib], 1u << k2);
                    }
                }
                __syncwarp();
                const unsigned cf = sm.conf[wib];
                const int kc = cf ? (__ffs(cf) - 1) : M;    // observations [ks, kc) are final
                // (d) commit
                const bool commit = active && lane < kc;
                if (commit) {
                    my_assoc = res;
                    stat |= st_k;
                    if (widx != FS2_NONE) fs2_store_lm(lm, widx, post);
                }
                if (kc < M) {
                    // touched set (only needed when another round follows): replace an entry or append, in lane order
                    int tpos = -1;
                    if (commit && widx != FS2_NONE) {
                        for (int t = 0; t < nt; ++t) if (sm.tidx[wib][t] == widx) tpos = t;
                    }
                    const unsigned newt = __ballot_sync(FS2_FULL, commit && widx != FS2_NONE && tpos < 0);
                    if (commit && widx != FS2_NONE) {
                        if (tpos < 0) tpos = nt + __popc(newt & lt_mask);
                        sm.tidx[wib][tpos] = widx;
                        sm.tlm[wib][tpos] = post;
                        sm.tbox[wib][tpos] = make_float4(pb.mx, pb.my, pb.rx, pb.ry);
                    }
                    nt += __popc(newt);
                }
                cnt += __popc(__ballot_sync(FS2_FULL, commit && !matched && widx != FS2_NONE));
                {   // weight *= likelihood of every committed observation (product tree; fp64 rounding only)
                    double lk = commit ? like : 1.0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) lk *= __shfl_xor_sync(FS2_FULL, lk, o);
                    pw *= lk;
                }
                ks = kc;
                __syncwarp();
            }
        }

        // ---------------- fallback: the reference's loop, observation by observation ----------------
        if (M > 0 && ks < M && (seq || ua.force_seq)) {
            __syncwarp();
            for (int k = ks; k < M; ++k) {
                const double kox = sm.ox[k], koy = sm.oy[k];
                int found = FS2_NONE;
                for (int base = 0; base < cnt && found == FS2_NONE; base += 32) {
                    int i = base + lane;
                    bool stop = false;
                    if (i < cnt) stop = fs2_stops_here(fs2_load_lm(lm, i), kox, koy, ua.gate);
                    unsigned b = __ballot_sync(FS2_FULL, stop);
                    if (b) found = base + __ffs(b) - 1;
                }
                int res;
                int st_k = 0;
                if (found != FS2_NONE) {
                    Fs2Lm in = fs2_load_lm(lm, found);   // uniform across the warp
                    double det = __dadd_rn(__dmul_rn(in.c00, in.c11), -__dmul_rn(in.c01, in.c10));
                    if (det == 0.0) {
                        st_k = 1; res = -2;
                    } else {
                        Fs2Lm post; double like;
                        st_k = fs2_ekf(px, py, pyaw, sm.zd[k], sm.za[k], ua.r00, ua.r01, ua.r10, ua.r11, in, &post, &like);
                        if (st_k == 2) res = -2;
                        else {
                            res = found;
                            if (lane == 0) fs2_store_lm(lm, found, post);
                            pw = __dmul_rn(pw, like);
                        }
                    }
                } else {
                    res = -1;
                    if (cnt < lcap) {
                        Fs2Lm post = fs2_new_landmark(px, py, pyaw, sm.zd[k], sm.za[k]);
                        if (lane == 0) fs2_store_lm(lm, cnt, post);
                        cnt += 1;
                    } else st_k = 8;
                }
                stat |= st_k;
                if (lane == k) my_assoc = res;
                __syncwarp();
            }
        }

        // ---------------- epilogue ----------------
        stat = __reduce_or_sync(FS2_FULL, stat);
        if (lane == 0) {
            if (ua.do_motion) { st.x[p] = px; st.y[p] = py; st.yaw[p] = pyaw; }
            if (M > 0) { st.w[p] = pw; st.count[p] = cnt; }
            if (stat) st.status[p] |= stat;
        }
        if (ua.assoc && is_obs) ua.assoc[(size_t)(ob.k0 + lane) * (size_t)st.P + p] = my_assoc;
        cnt_cur = cnt_next;
        slot_cur = slot_next;
        __syncwarp();
    }
}
