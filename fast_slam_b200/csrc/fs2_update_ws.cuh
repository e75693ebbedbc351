// fs2_update_ws.cuh -- warp-specialised form of the fused update kernel (same algorithm and device functions as
// fs2_update.cuh; see the description there).
//
// The update has two halves with opposite needs.  Streaming + screening + exact gating is memory-latency bound,
// needs few registers and wants many warps; the EKF / likelihood half is a long fp64 dependency chain that wants
// ~128 registers.  One CTA therefore runs two kinds of warps and moves registers between them with setmaxnreg:
//
//   screeners (FS2_SW warps, FS2_S_REGS registers)   one particle at a time: TMA ring over the map, fp32 box
//             screen through the observation cell tables, exact fp64 gate of the survivors.  The result -- per
//             observation the <= 4 lowest matching landmark indices -- is written straight into a TICKET in
//             shared memory together with the particle's header (pose, weight, noise, count, slot), and
//             published on an mbarrier.  Association does not depend on the pose (quirk Q1), so screeners
//             never touch it beyond fetching it for the applier.
//   appliers  (FS2_AW warps, FS2_A_REGS registers)   take tickets in order and run motion, the speculative
//             order-preserving application of the observations (EKF updates, new landmarks, weight), the
//             sequential fallback and the epilogue stores.
//
// Tickets form a ring of FS2_QS slots with a full and an empty mbarrier each; ticket numbers come from two
// shared-memory counters, so any screener feeds any applier.
#pragma once
#include "fs2_update.cuh"

#ifndef FS2_SW
#define FS2_SW 8
#endif
#ifndef FS2_AW
#define FS2_AW 4
#endif
#ifndef FS2_S_REGS
#define FS2_S_REGS 56
#endif
#ifndef FS2_A_REGS
#define FS2_A_REGS 128
#endif
#ifndef FS2_QS
#define FS2_QS 16
#endif
#ifndef FS2_WS_MINB
#define FS2_WS_MINB 2
#endif
#define FS2_WS_THREADS ((FS2_SW + FS2_AW) * 32)
// setmaxnreg acts on warpgroups (4 consecutive warps, all with the same value): a role boundary inside a warpgroup
// hangs the kernel (seen with 5:3, 6:2 and 4:2 builds)
static_assert(FS2_SW % 4 == 0 && FS2_AW % 4 == 0, "screener and applier warps must come in multiples of four");

struct Fs2Ticket {
    int4 ml[32];                 // per observation: its <= 4 lowest exact matches on the pre-step map
    double px, py, pyaw, pw, nz; // particle header, fetched by the screener one particle ahead
    long long p;
    int cnt, slot;
    unsigned ovf;                // observations with more than 4 matches
    int pad;
};

struct Fs2WsSmem {
    double ox[32], oy[32], zd[32], za[32];
    alignas(8) float2 of[33];
    unsigned tab1[FS2_G1P * FS2_G1P];
    unsigned tab2[FS2_G2P * FS2_G2P];
    alignas(128) unsigned char ring[FS2_SW][FS2_NST][FS2_CHUNK_BYTES];
    alignas(8) unsigned long long bar[FS2_SW][FS2_NST];
    alignas(8) unsigned long long q_full[FS2_QS];
    alignas(8) unsigned long long q_empty[FS2_QS];
    int qidx[FS2_SW][FS2_QCAP];
    unsigned qmask[FS2_SW][FS2_QCAP];
    alignas(16) Fs2Ticket tk[FS2_QS];
    unsigned q_head, q_tail;
    unsigned conf[FS2_AW];
    int bound[FS2_AW][32];
    alignas(16) Fs2Lm tlm[FS2_AW][32];
    alignas(16) float4 tbox[FS2_AW][32];
    int tidx[FS2_AW][32];
};

__device__ __forceinline__ void fs2_mbar_arrive(unsigned long long *bar)
{
    unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(b) : "memory");
}

// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fs2_ws_screener(Fs2WsSmem &sm, const Fs2State &st, const Fs2ObsBatch &ob,
                                                const Fs2UpdateArgs &ua, int sw, int lane)
{
    const unsigned lt_mask = (1u << lane) - 1u;
    const int M = ob.M;
    const int lcap = st.lcap;
    const int64_t step = (int64_t)gridDim.x * FS2_SW;
    const bool streaming = (ua.force_seq == 0) && (M > 0);
    unsigned char *ring = &sm.ring[sw][0][0];
    unsigned long long *bars = &sm.bar[sw][0];
    const unsigned bar_s0 = (unsigned)__cvta_generic_to_shared(bars);
    const unsigned qfull0 = (unsigned)__cvta_generic_to_shared(&sm.q_full[0]);
    const unsigned qempty0 = (unsigned)__cvta_generic_to_shared(&sm.q_empty[0]);
    (void)qfull0;
    unsigned gc = 0, gp = 0;
    int64_t p = (int64_t)blockIdx.x * FS2_SW + sw;
    int cnt_cur = 0, slot_cur = 0;
    double px_n = 0.0, py_n = 0.0, pyaw_n = 0.0, pw_n = 0.0, nz_n = 0.0;
    if (p < st.P) {
        cnt_cur = st.count[p]; slot_cur = st.slot[p];
        px_n = st.x[p]; py_n = st.y[p]; pyaw_n = st.yaw[p]; pw_n = st.w[p];
        if (ua.do_motion) nz_n = ua.noise[p];
    }
    auto issue = [&](const double *map, int cnt_map, int c) {
        const unsigned bytes = (unsigned)min(FS2_CHUNK, cnt_map - FS2_CHUNK * c) * 48u;
        fs2_tma_load(ring + (gp % FS2_NST) * FS2_CHUNK_BYTES,
                     reinterpret_cast<const unsigned char *>(map) + (size_t)c * FS2_CHUNK_BYTES, bytes, bars + (gp % FS2_NST));
    };
    if (streaming && p < st.P) {
        const double *map = st.lm + (size_t)slot_cur * 6 * (size_t)lcap;
        const int pre = min(FS2_NST - 1, (cnt_cur + FS2_CHUNK - 1) / FS2_CHUNK);
        for (int c = 0; c < pre; ++c) { if (lane == 0) issue(map, cnt_cur, c); ++gp; }
    }
    for (; p < st.P; p += step) {
        const int cnt = cnt_cur;
        const double *lm = st.lm + (size_t)slot_cur * 6 * (size_t)lcap;
        // ---- take a ticket and wait for its slot to be free ----
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(&sm.q_head, 1u);
        t = __shfl_sync(FS2_FULL, t, 0);
        const unsigned qs = t % FS2_QS;
        fs2_mbar_wait(qempty0 + 8u * qs, ((t / FS2_QS) & 1u) ^ 1u);
        Fs2Ticket &tk = sm.tk[qs];
        tk.ml[lane] = make_int4(FS2_NONE, FS2_NONE, FS2_NONE, FS2_NONE);
        {
            // __move_particle (fast_slam_2.py:69-87) happens here: the screeners have slack, the appliers do not,
            // and association does not look at the pose (quirk Q1).  All lanes compute it (warp-uniform), lane 0 stores.
            double mx = px_n, my = py_n, myaw = pyaw_n;
            if (ua.do_motion) fs2_move(mx, my, myaw, ua.rotation, ua.translation, nz_n);
            if (lane == 0) {
                if (ua.do_motion) { st.x[p] = mx; st.y[p] = my; st.yaw[p] = myaw; }
                tk.px = mx; tk.py = my; tk.pyaw = myaw; tk.pw = pw_n; tk.nz = nz_n;
                tk.p = p; tk.cnt = cnt; tk.slot = slot_cur; tk.ovf = 0u;
            }
        }
        // ---- next particle's header, one particle ahead ----
        const int64_t pn = p + step;
        int cnt_next = 0, slot_next = 0;
        if (pn < st.P) {
            cnt_next = st.count[pn]; slot_next = st.slot[pn];
            px_n = st.x[pn]; py_n = st.y[pn]; pyaw_n = st.yaw[pn]; pw_n = st.w[pn];
            if (ua.do_motion) nz_n = ua.noise[pn];
        }
        __syncwarp();
        if (streaming) {
            const int nchunks = (cnt + FS2_CHUNK - 1) / FS2_CHUNK;
            int issued = min(FS2_NST - 1, nchunks);
            int qn = 0;
            for (int c = 0; c < nchunks; ++c) {
                if (issued < nchunks) {
                    if (lane == 0) issue(lm, cnt, issued);
                    ++gp; ++issued;
                }
                const unsigned stage = gc % FS2_NST;
                fs2_mbar_wait(bar_s0 + 8u * stage, (gc / FS2_NST) & 1u);
                ++gc;
                const int iA = c * FS2_CHUNK + lane, iB = iA + 32;
                const double2 *srcA = reinterpret_cast<const double2 *>(ring + stage * FS2_CHUNK_BYTES + 48 * lane);
                const double2 *srcB = srcA + 96;
                Fs2Box bA, bB;
                bA.mx = bA.my = 0.f; bA.rx = bA.ry = -1.f;
                bB = bA;
                unsigned maskA = 0, maskB = 0, hasB = 0;
                if (c * FS2_CHUNK + 32 < cnt) {     // warp-uniform: the second half of the chunk holds landmarks
                    if (iA < cnt) {
                        const double2 a0 = srcA[0], a1 = srcA[1], a2 = srcA[2];
                        bA = fs2_box(a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, ua.gate_f, ob.slack);
                    }
                    if (iB < cnt) {
                        const double2 b0 = srcB[0], b1 = srcB[1], b2 = srcB[2];
                        bB = fs2_box(b0.x, b0.y, b1.x, b1.y, b2.x, b2.y, ua.gate_f, ob.slack);
                    }
                    const unsigned candA = fs2_candidates(sm, ob, bA), candB = fs2_candidates(sm, ob, bB);
                    unsigned restA, restB;
                    maskA = fs2_box_filter2(sm, bA, candA, &restA);      // two loop-free streams the compiler interleaves
                    maskB = fs2_box_filter2(sm, bB, candB, &restB);
                    if (__any_sync(FS2_FULL, (restA | restB) != 0u)) {   // rare: a landmark with more than two candidates
                        maskA |= fs2_box_filter_rest(sm, bA, restA);
                        maskB |= fs2_box_filter_rest(sm, bB, restB);
                    }
                    hasB = __ballot_sync(FS2_FULL, maskB != 0);
                } else {                            // tail of the map: at most 32 landmarks left
                    if (iA < cnt) {
                        const double2 a0 = srcA[0], a1 = srcA[1], a2 = srcA[2];
                        bA = fs2_box(a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, ua.gate_f, ob.slack);
                        maskA = fs2_box_filter(sm, bA, fs2_candidates(sm, ob, bA));
                    }
                }
                const unsigned hasA = __ballot_sync(FS2_FULL, maskA != 0);
                if (hasA | hasB) {
                    if (maskA) {
                        const int pos = qn + __popc(hasA & lt_mask);
                        sm.qidx[sw][pos] = iA;
                        sm.qmask[sw][pos] = maskA;
                    }
                    qn += __popc(hasA);
                    if (maskB) {
                        const int pos = qn + __popc(hasB & lt_mask);
                        sm.qidx[sw][pos] = iB;
                        sm.qmask[sw][pos] = maskB;
                    }
                    qn += __popc(hasB);
                    if (qn > FS2_QCAP - 64) {
                        __syncwarp();
                        fs2_drain(sm.qidx[sw], sm.qmask[sw], tk.ml, &tk.ovf, sm.ox, sm.oy, lane, lm, qn, ua.gate);
                        qn = 0;
                    }
                }
                __syncwarp();
            }
            if (pn < st.P) {   // ring empty: start on the next particle's map
                const double *mapn = st.lm + (size_t)slot_next * 6 * (size_t)lcap;
                const int pre = min(FS2_NST - 1, (cnt_next + FS2_CHUNK - 1) / FS2_CHUNK);
                for (int c = 0; c < pre; ++c) { if (lane == 0) issue(mapn, cnt_next, c); ++gp; }
            }
            fs2_drain(sm.qidx[sw], sm.qmask[sw], tk.ml, &tk.ovf, sm.ox, sm.oy, lane, lm, qn, ua.gate);
        }
        __syncwarp();
        if (lane == 0) fs2_mbar_arrive(&sm.q_full[qs]);   // publish the ticket
        cnt_cur = cnt_next;
        slot_cur = slot_next;
    }
}

// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fs2_ws_applier(Fs2WsSmem &sm, const Fs2State &st, const Fs2ObsBatch &ob,
                                               const Fs2UpdateArgs &ua, int aw, int lane, unsigned total)
{
    const unsigned lt_mask = (1u << lane) - 1u;
    const int M = ob.M;
    const int lcap = st.lcap;
    const bool is_obs = lane < M;
    const double zd = sm.zd[lane], za = sm.za[lane];
    const double oxd = sm.ox[lane], oyd = sm.oy[lane];
    const float2 myof = sm.of[lane];
    const unsigned qfull0 = (unsigned)__cvta_generic_to_shared(&sm.q_full[0]);

    // software-pipelined ticket fetch: the NEXT ticket's header, match list and first-match landmark are read
    // (the landmark load only issued) before the current particle is processed
    struct Held { int64_t p; double px, py, pyaw, pw, nz; int cnt, slot; int4 ml; bool ovf; Fs2Lm in; bool valid; };
    auto fetch = [&](Held &h) {
        unsigned u = 0;
        if (lane == 0) u = atomicAdd(&sm.q_tail, 1u);
        u = __shfl_sync(FS2_FULL, u, 0);
        h.valid = u < total;
        if (!h.valid) return;
        const unsigned qs = u % FS2_QS;
        fs2_mbar_wait(qfull0 + 8u * qs, (u / FS2_QS) & 1u);
        const Fs2Ticket &tk = sm.tk[qs];
        h.p = tk.p; h.px = tk.px; h.py = tk.py; h.pyaw = tk.pyaw; h.pw = tk.pw; h.nz = tk.nz;
        h.cnt = tk.cnt; h.slot = tk.slot;
        h.ml = tk.ml[lane];
        h.ovf = (tk.ovf >> lane) & 1u;
        __syncwarp();
        if (lane == 0) fs2_mbar_arrive(&sm.q_empty[qs]);   // everything is in registers: hand the slot back
        h.in.x = h.in.y = h.in.c00 = h.in.c01 = h.in.c10 = h.in.c11 = 0.0;
        if (lane < M && h.ml.x != FS2_NONE)                // first round's landmark, needed ~a particle later
            h.in = fs2_load_lm(st.lm + (size_t)h.slot * 6 * (size_t)lcap, h.ml.x);
    };
    Held cur, nxt;
    fetch(cur);
    while (cur.valid) {
        fetch(nxt);
        const int64_t p = cur.p;
        double px = cur.px, py = cur.py, pyaw = cur.pyaw, pw = cur.pw;
        const double nz = cur.nz;
        int cnt = cur.cnt;
        double *lm = st.lm + (size_t)cur.slot * 6 * (size_t)lcap;
        const int4 ml = cur.ml;
        const bool ml_overflow = cur.ovf;
        const Fs2Lm in0 = cur.in;

        int stat = 0;
        (void)nz;                                   // the screener already moved the particle
        int ks = 0, nt = 0;
        bool seq = (ua.force_seq != 0);
        int my_assoc = -3;

#ifdef FS2_DEBUG_ROUNDS
        int dbg_rounds = 0;
#endif
        while (ks < M && !seq) {
#ifdef FS2_DEBUG_ROUNDS
            ++dbg_rounds;
#endif
            const bool active = is_obs && lane >= ks;
            int a_un = FS2_NONE;
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
                            for (int t = 0; t < nt; ++t) touched |= (sm.tidx[aw][t] == cand);
                            if (!touched) a_un = cand;
                        }
                    }
                    if (a_un == FS2_NONE && ml_overflow) exhausted = true;
                }
            }
            int a_t = FS2_NONE, a_t_pos = -1;
            if (active) {
                for (int t = 0; t < nt; ++t) {
                    const float4 tb = sm.tbox[aw][t];
                    const int ti = sm.tidx[aw][t];
                    if (ti < a_t && fabsf(myof.x - tb.x) < tb.z && fabsf(myof.y - tb.y) < tb.w) {
                        if (fs2_stops_here(sm.tlm[aw][t], oxd, oyd, ua.gate)) { a_t = ti; a_t_pos = t; }
                    }
                }
            }
            if (__any_sync(FS2_FULL, exhausted)) { seq = true; break; }
            const bool from_t = a_t < a_un;
            const int a = from_t ? a_t : a_un;
            const bool matched = active && a != FS2_NONE;
            const unsigned unm = __ballot_sync(FS2_FULL, active && !matched);
            const int app_rank = __popc(unm & lt_mask);
            Fs2Lm post;
            post.x = post.y = post.c00 = post.c01 = post.c10 = post.c11 = 0.0;
            double like = 1.0;
            int st_k = 0, widx = FS2_NONE, res = -3;
            if (matched) {
                const Fs2Lm in = from_t ? sm.tlm[aw][a_t_pos] : ((nt == 0) ? in0 : fs2_load_lm(lm, a));
                const double det = __dadd_rn(__dmul_rn(in.c00, in.c11), -__dmul_rn(in.c01, in.c10));
                if (det == 0.0) {
                    st_k = 1; res = -2;
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
                    st_k = 8;
                }
            }
            Fs2Box pb;
            pb.mx = 0.f; pb.my = 0.f; pb.rx = -1.f; pb.ry = -1.f;
            if (widx != FS2_NONE) pb = fs2_box(post.x, post.y, post.c00, post.c01, post.c10, post.c11, ua.gate_f, ob.slack);
            const int key = matched ? a : (widx != FS2_NONE ? widx : (0x40000000 | lane));
            const unsigned act_mask = __ballot_sync(FS2_FULL, active);
            const unsigned same = __match_any_sync(FS2_FULL, active ? key : (0x50000000 | lane));
            sm.bound[aw][lane] = matched ? a : FS2_NONE;
            // same landmark as an earlier observation of the round: one ballot, no shared-memory atomic
            const unsigned cf_same = __ballot_sync(FS2_FULL, matched && (same & lt_mask & act_mask) != 0u);
            if (lane == 0) sm.conf[aw] = cf_same;
            __syncwarp();
            if (widx != FS2_NONE) {
                unsigned later = fs2_candidates(sm, ob, pb) & act_mask & ~(lt_mask | (1u << lane));
                later = fs2_box_filter(sm, pb, later);
                while (later) {
                    const int k2 = __ffs(later) - 1;
                    later &= later - 1;
                    if (widx < sm.bound[aw][k2] && fs2_stops_here(post, sm.ox[k2], sm.oy[k2], ua.gate))
                        atomicOr(&sm.conf[aw], 1u << k2);
                }
            }
            __syncwarp();
            const unsigned cf = sm.conf[aw];
            const int kc = cf ? (__ffs(cf) - 1) : M;
#ifdef FS2_DEBUG_ROUNDS
            if (kc < M && dbg_rounds == 1) {
                stat |= ((cf_same >> kc) & 1u) ? 256 : 512;               // same landmark twice / captured by a new state
                stat |= (sm.bound[aw][kc] == FS2_NONE) ? 1024 : 2048;     // the dependent observation had no match / had one
            }
#endif
            const bool commit = active && lane < kc;
            if (commit) {
                my_assoc = res;
                stat |= st_k;
                if (widx != FS2_NONE) fs2_store_lm(lm, widx, post);
            }
            if (kc < M) {
                int tpos = -1;
                if (commit && widx != FS2_NONE) {
                    for (int t = 0; t < nt; ++t) if (sm.tidx[aw][t] == widx) tpos = t;
                }
                const unsigned newt = __ballot_sync(FS2_FULL, commit && widx != FS2_NONE && tpos < 0);
                if (commit && widx != FS2_NONE) {
                    if (tpos < 0) tpos = nt + __popc(newt & lt_mask);
                    sm.tidx[aw][tpos] = widx;
                    sm.tlm[aw][tpos] = post;
                    sm.tbox[aw][tpos] = make_float4(pb.mx, pb.my, pb.rx, pb.ry);
                }
                nt += __popc(newt);
            }
            cnt += __popc(__ballot_sync(FS2_FULL, commit && !matched && widx != FS2_NONE));
            {
                double lk = commit ? like : 1.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) lk *= __shfl_xor_sync(FS2_FULL, lk, o);
                pw *= lk;
            }
            ks = kc;
            __syncwarp();
        }

        if (M > 0 && ks < M && seq) {   // the reference's loop, observation by observation
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
                int res, st_k = 0;
                if (found != FS2_NONE) {
                    Fs2Lm in = fs2_load_lm(lm, found);
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

#ifdef FS2_DEBUG_ROUNDS      // diagnostics build only: how the step went for this particle, in spare status bits
        if (seq) stat |= 16;
        if (dbg_rounds > 1) stat |= 32;
        if (dbg_rounds > 2) stat |= 64;
        if (dbg_rounds > 4) stat |= 128;
#endif
        stat = __reduce_or_sync(FS2_FULL, stat);
        if (lane == 0) {
            if (M > 0) { st.w[p] = pw; st.count[p] = cnt; }
            if (stat) st.status[p] |= stat;
        }
        if (ua.assoc && is_obs) ua.assoc[(size_t)(ob.k0 + lane) * (size_t)st.P + p] = my_assoc;
        __syncwarp();
        cur = nxt;
    }
}

__global__ void __launch_bounds__(FS2_WS_THREADS, FS2_WS_MINB)
fs2_update_ws_kernel(const Fs2State st, const __grid_constant__ Fs2ObsBatch ob, const Fs2UpdateArgs ua)
{
    extern __shared__ __align__(128) unsigned char fs2_smem_raw[];
    Fs2WsSmem &sm = *reinterpret_cast<Fs2WsSmem *>(fs2_smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 32) {
        sm.ox[lane] = ob.ox[lane]; sm.oy[lane] = ob.oy[lane];
        sm.zd[lane] = ob.zd[lane]; sm.za[lane] = ob.za[lane];
        sm.of[lane] = make_float2(ob.oxf[lane], ob.oyf[lane]);
        if (lane == 0) {
            sm.of[32] = make_float2(__int_as_float(0x7f800000), __int_as_float(0x7f800000));
            sm.q_head = 0u; sm.q_tail = 0u;
        }
        if (lane < FS2_QS) { fs2_mbar_init(&sm.q_full[lane], 1); fs2_mbar_init(&sm.q_empty[lane], 1); }
    }
    for (int i = threadIdx.x; i < FS2_G1P * FS2_G1P; i += blockDim.x) sm.tab1[i] = ob.tab1[i];
    for (int i = threadIdx.x; i < FS2_G2P * FS2_G2P; i += blockDim.x) sm.tab2[i] = ob.tab2[i];
    if (warp < FS2_SW && lane == 0) {
#pragma unroll
        for (int s = 0; s < FS2_NST; ++s) fs2_mbar_init(&sm.bar[warp][s], 1);
    }
    fs2_fence_mbar_init();
    __syncthreads();
    // particles of this CTA: screener w takes p = blockIdx*SW + w, + k*gridDim*SW
    const int64_t step = (int64_t)gridDim.x * FS2_SW;
    unsigned total = 0;
    for (int w = 0; w < FS2_SW; ++w) {
        const int64_t p0 = (int64_t)blockIdx.x * FS2_SW + w;
        if (p0 < st.P) total += (unsigned)((st.P - p0 + step - 1) / step);
    }
    if (warp < FS2_SW) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(FS2_S_REGS));
        fs2_ws_screener(sm, st, ob, ua, warp, lane);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(FS2_A_REGS));
        fs2_ws_applier(sm, st, ob, ua, warp - FS2_SW, lane, total);
    }
}
