// fs2_update.cuh -- the fused motion + association + EKF + weighting kernel (rows A2-A6 of SURVEY.md 8a).
//
// One warp owns one particle at a time (persistent grid, warps stride over particles).
//
//   phase 1  stream the particle's map ONCE: 32 landmarks per chunk (1536 contiguous bytes) are staged
//            into a per-warp shared-memory ring with cp.async (3 x 16 B per lane, coalesced), lane j reads
//            landmark j back with three conflict-free LDS.128 and screens it against all observations of
//            the step with the conservative fp32 box (fs2_box).  The observations sit in the kernel's
//            constant parameter block, so one screen is 2 FADD + 2 FSETP + 1 predicated LOP3.
//   phase 2  the few (landmark, observation) pairs that pass are queued in shared memory and re-tested
//            with the exact fp64 gate (fs2_gate_test, bit-identical to the oracle); each observation
//            (lane k = observation k) collects, in ascending landmark order, up to 4 exact matches on the
//            pre-step map.
//   phase 3  observations are applied in list order (quirk Q7) by speculation: every remaining
//            observation picks its first match (pre-step list patched with the landmarks touched so far
//            this step), all EKF updates / new landmarks are computed in parallel (lane k = observation
//            k), then each observation k' is checked against the POST state of every earlier one j:
//            same landmark, or k' would now match j's landmark at a lower index than its own choice.
//            Everything before the first such dependency is committed; the rest is re-speculated against
//            the updated state.  Independent observations (the normal case) finish in one round.
//   fallback observation-by-observation exact scan of the map in global memory -- literally the
//            reference's loop (landmark_utils.py:103-117) -- used when an observation's match list
//            overflowed and was exhausted, or when FS2_FLAG_FORCE_SEQUENTIAL is set (cross-check).
#pragma once
#include "fs2_math.cuh"

#define FS2_WPB 8          // warps per block
#define FS2_NST 3          // cp.async ring stages per warp
#define FS2_CHUNK_BYTES 1536
#define FS2_QCAP 96        // candidate queue entries per warp
#define FS2_NONE 0x7fffffff
#define FS2_FULL 0xffffffffu

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
    float oxf[32], oyf[32];  // padded with +inf beyond M
    float slack;             // 2.4e-7 * max(|ox|,|oy|) + tiny
    int32_t M;
    int32_t k0;              // index of the batch's first observation in the step's list
    int32_t pad;
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
};

struct Fs2UpdateSmem {
    double ox[32], oy[32], zd[32], za[32];
    float oxf[32], oyf[32];
    unsigned char ring[FS2_WPB][FS2_NST][FS2_CHUNK_BYTES];
    int qidx[FS2_WPB][FS2_QCAP];
    unsigned qmask[FS2_WPB][FS2_QCAP];
    Fs2Lm post[FS2_WPB][32];    // speculative result of observation k (lane k)
    float4 pbox[FS2_WPB][32];   // its screen box
    int pidx[FS2_WPB][32];      // landmark index it writes (FS2_NONE: writes nothing)
    Fs2Lm tlm[FS2_WPB][32];     // landmarks touched so far this step (current state)
    float4 tbox[FS2_WPB][32];
    int tidx[FS2_WPB][32];
};

__device__ __forceinline__ void fs2_cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void fs2_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void fs2_cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

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

// per-observation sorted list of exact matches on the pre-step map (registers of lane k)
struct Fs2MatchList {
    int v0, v1, v2, v3, n;
    bool overflow;
    __device__ __forceinline__ void clear() { v0 = v1 = v2 = v3 = FS2_NONE; n = 0; overflow = false; }
    __device__ __forceinline__ void push(int idx)
    {
        if (n == 0) v0 = idx; else if (n == 1) v1 = idx; else if (n == 2) v2 = idx; else if (n == 3) v3 = idx;
        else overflow = true;
        if (n < 4) ++n;
    }
    __device__ __forceinline__ int get(int i) const { return i == 0 ? v0 : i == 1 ? v1 : i == 2 ? v2 : v3; }
};

// phase 2: exact re-test of the queued candidates; lane k appends the survivors of observation k
__device__ __forceinline__ void fs2_drain(Fs2UpdateSmem &sm, int wib, int lane, const double *lm, int qn,
                                          double gate, Fs2MatchList &ml)
{
    for (int e0 = 0; e0 < qn; e0 += 32) {
        int e = e0 + lane;
        unsigned pm = 0;
        if (e < qn) {
            int idx = sm.qidx[wib][e];
            unsigned m = sm.qmask[wib][e];
            Fs2Lm l = fs2_load_lm(lm, idx);
            Fs2Gate g = fs2_gate_prepare(l.c00, l.c01, l.c10, l.c11);
            while (m) {
                int k = __ffs(m) - 1;
                m &= m - 1;
                if (g.singular || fs2_gate_test(g, l.x, l.y, sm.ox[k], sm.oy[k], gate)) pm |= (1u << k);
            }
        }
        unsigned any = __reduce_or_sync(FS2_FULL, pm);
        while (any) {
            int k = __ffs(any) - 1;
            any &= any - 1;
            unsigned b = __ballot_sync(FS2_FULL, (pm >> k) & 1u);
            if (lane == k) {
                while (b) {
                    int src = __ffs(b) - 1;
                    b &= b - 1;
                    ml.push(sm.qidx[wib][e0 + src]);  // queue order == ascending landmark index
                }
            }
        }
    }
    __syncwarp();
}

template <int MP>
__global__ void __launch_bounds__(FS2_WPB * 32, 2)
fs2_update_kernel(const Fs2State st, const __grid_constant__ Fs2ObsBatch ob, const Fs2UpdateArgs ua)
{
    extern __shared__ __align__(16) unsigned char fs2_smem_raw[];
    Fs2UpdateSmem &sm = *reinterpret_cast<Fs2UpdateSmem *>(fs2_smem_raw);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    if (threadIdx.x < 32) {
        sm.ox[lane] = ob.ox[lane]; sm.oy[lane] = ob.oy[lane];
        sm.zd[lane] = ob.zd[lane]; sm.za[lane] = ob.za[lane];
        sm.oxf[lane] = ob.oxf[lane]; sm.oyf[lane] = ob.oyf[lane];
    }
    __syncthreads();
    const int M = ob.M;
    const int lcap = st.lcap;
    const int64_t nwarps = (int64_t)gridDim.x * FS2_WPB;
    const bool is_obs = lane < M;
    const double zd = sm.zd[lane], za = sm.za[lane];
    const double oxd = sm.ox[lane], oyd = sm.oy[lane];
    const float myoxf = sm.oxf[lane], myoyf = sm.oyf[lane];

    for (int64_t p = (int64_t)blockIdx.x * FS2_WPB + wib; p < st.P; p += nwarps) {
        double px = st.x[p], py = st.y[p], pyaw = st.yaw[p], pw = st.w[p];
        int cnt = st.count[p];
        double *lm = st.lm + (size_t)st.slot[p] * 6 * (size_t)lcap;
        int stat = 0;
        if (ua.do_motion) fs2_move(px, py, pyaw, ua.rotation, ua.translation, ua.noise[p]);

        int ks = 0;            // first observation not yet applied
        int nt = 0;            // touched landmarks in sm.tlm / tidx / tbox
        bool seq = (ua.force_seq != 0);
        int my_assoc = -3;     // result of observation `lane`
        Fs2MatchList ml;
        ml.clear();

        if (!seq && M > 0) {
            // ---------------- phase 1: stream + screen ----------------
            const unsigned char *gsrc = reinterpret_cast<const unsigned char *>(lm);
            const int nchunks = (cnt + 31) >> 5;
            unsigned char *ring = &sm.ring[wib][0][0];
            int qn = 0;
#pragma unroll
            for (int c = 0; c < FS2_NST - 1; ++c) {
                if (c < nchunks) {
                    int ng = min(32, cnt - 32 * c) * 3;
                    for (int g = lane; g < ng; g += 32)
                        fs2_cp_async16(ring + c * FS2_CHUNK_BYTES + 16 * g, gsrc + (size_t)c * FS2_CHUNK_BYTES + 16 * g);
                }
                fs2_cp_async_commit();
            }
            for (int c = 0; c < nchunks; ++c) {
                int cn = c + FS2_NST - 1;
                if (cn < nchunks) {
                    int ng = min(32, cnt - 32 * cn) * 3;
                    unsigned char *dst = ring + (cn % FS2_NST) * FS2_CHUNK_BYTES;
                    for (int g = lane; g < ng; g += 32)
                        fs2_cp_async16(dst + 16 * g, gsrc + (size_t)cn * FS2_CHUNK_BYTES + 16 * g);
                }
                fs2_cp_async_commit();
                fs2_cp_async_wait<FS2_NST - 1>();
                __syncwarp();
                const int i = c * 32 + lane;
                Fs2Box b;
                if (i < cnt) {
                    const double2 *src = reinterpret_cast<const double2 *>(ring + (c % FS2_NST) * FS2_CHUNK_BYTES + 48 * lane);
                    double2 a0 = src[0], a1 = src[1], a2 = src[2];
                    b = fs2_box(a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, ua.gate_f, ob.slack);
                } else {
                    b.mx = 0.f; b.my = 0.f; b.rx = -1.f; b.ry = -1.f;
                }
                unsigned mask = 0;
#pragma unroll
                for (int k = 0; k < MP; ++k) {
                    float dx = ob.oxf[k] - b.mx;
                    float dy = ob.oyf[k] - b.my;
                    if (fabsf(dx) < b.rx && fabsf(dy) < b.ry) mask |= (1u << k);
                }
                unsigned has = __ballot_sync(FS2_FULL, mask != 0);
                if (has) {
                    int pos = qn + __popc(has & lt_mask);
                    if (mask) {
                        sm.qidx[wib][pos] = i;
                        sm.qmask[wib][pos] = mask;
                    }
                    qn += __popc(has);
                    if (qn > FS2_QCAP - 32) {
                        __syncwarp();
                        fs2_drain(sm, wib, lane, lm, qn, ua.gate, ml);
                        qn = 0;
                    }
                }
                __syncwarp();  // ring stage is recycled by the next iteration's cp.async
            }
            fs2_cp_async_wait<0>();
            __syncwarp();
            // ---------------- phase 2: exact re-test of what is left in the queue ----------------
            fs2_drain(sm, wib, lane, lm, qn, ua.gate, ml);

            // ---------------- phase 3: speculative, order-preserving application ----------------
            while (ks < M && !seq) {
                const bool active = is_obs && lane >= ks;
                // (a) association of every remaining observation against the current state
                int a_un = FS2_NONE;        // first pre-step match that has not been touched this step
                bool exhausted = false;
                if (active) {
                    int i = 0;
                    for (; i < ml.n; ++i) {
                        int cand = ml.get(i);
                        bool touched = false;
                        for (int t = 0; t < nt; ++t) touched |= (sm.tidx[wib][t] == cand);
                        if (!touched) { a_un = cand; break; }
                    }
                    if (a_un == FS2_NONE && ml.overflow) exhausted = true;  // more pre-step matches exist, unknown
                }
                int a_t = FS2_NONE;         // lowest touched landmark whose CURRENT state stops the scan
                int a_t_pos = -1;
                if (active) {
                    for (int t = 0; t < nt; ++t) {
                        float4 tb = sm.tbox[wib][t];
                        int ti = sm.tidx[wib][t];
                        if (ti < a_t && fabsf(myoxf - tb.x) < tb.z && fabsf(myoyf - tb.y) < tb.w) {
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
                    Fs2Lm in = from_t ? sm.tlm[wib][a_t_pos] : fs2_load_lm(lm, a);
                    double det = __dadd_rn(__dmul_rn(in.c00, in.c11), -__dmul_rn(in.c01, in.c10));
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
                if (widx != FS2_NONE) pb = fs2_box(post.x, post.y, post.c00, post.c01, post.c10, post.c11, ua.gate_f, ob.slack);
                else { pb.mx = 0.f; pb.my = 0.f; pb.rx = -1.f; pb.ry = -1.f; }
                sm.post[wib][lane] = post;
                sm.pbox[wib][lane] = make_float4(pb.mx, pb.my, pb.rx, pb.ry);
                sm.pidx[wib][lane] = widx;
                __syncwarp();
                // (c) does observation `lane` depend on an earlier, not yet committed one?
                bool conflict = false;
                const int bound = matched ? a : FS2_NONE;   // indices below this would pre-empt my choice
                for (int j = ks; j < M - 1; ++j) {
                    int ij = sm.pidx[wib][j];
                    float4 jb = sm.pbox[wib][j];
                    if (active && lane > j && ij != FS2_NONE) {
                        if (ij == a) conflict = true;
                        else if (ij < bound && fabsf(myoxf - jb.x) < jb.z && fabsf(myoyf - jb.y) < jb.w) {
                            if (fs2_stops_here(sm.post[wib][j], oxd, oyd, ua.gate)) conflict = true;
                        }
                    }
                }
                const unsigned cf = __ballot_sync(FS2_FULL, conflict);
                const int kc = cf ? (__ffs(cf) - 1) : M;    // observations [ks, kc) are final
                // (d) commit
                const bool commit = active && lane < kc;
                if (commit) {
                    my_assoc = res;
                    stat |= st_k;
                    if (widx != FS2_NONE) fs2_store_lm(lm, widx, post);
                }
                // touched set: replace an existing entry or append, in lane order
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
                cnt += __popc(__ballot_sync(FS2_FULL, commit && !matched && widx != FS2_NONE));
                for (int j = ks; j < kc; ++j) {             // weight *= likelihood, in observation order
                    double lj = __shfl_sync(FS2_FULL, like, j);
                    pw = __dmul_rn(pw, lj);
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
        __syncwarp();
    }
}
