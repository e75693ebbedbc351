// fs2_update_ws.cuh -- warp-specialised form of the fused update kernel (same algorithm and device functions as
// fs2_update.cuh; see the description there).
//
// The update has two halves with opposite needs.  Streaming + screening + exact gating is memory-latency bound,
// needs few registers and wants many warps; the EKF / likelihood half is a long fp64 dependency chain that wants
// ~128 registers.  One CTA therefore runs two kinds of warps and moves registers between them with setmaxnreg:
//
//   screeners (FS2_SW warps, FS2_S_REGS registers)   one particle at a time: TMA ring over the map, then per
//             landmark a LEVEL test instead of a box: a safe covariance (fs2_box's conditions) with
//             max(c00, c11) <= amax1 / amax2 is known to have a gate box no wider than e1 / e2, so the
//             observation cell table of that level, looked up at the landmark's mean, already lists every
//             observation the landmark could gate.  With the fine table (48 x 48 cells over the observations'
//             bounding square) that list is empty for almost every landmark no observation belongs to, so the
//             loop carries no box, no sqrt and no per-observation compare.  Landmarks with a non-empty list are
//             queued; the drain builds their box (fs2_box), filters the list through it and runs the exact fp64
//             gate on what is left.  The result -- per observation the <= 4 lowest matching landmark indices --
//             is written straight into a TICKET in shared memory together with the particle's header (pose after
//             __move_particle, weight, count, slot).  Association does not depend on the pose (quirk Q1).
//   appliers  (FS2_AW warps, FS2_A_REGS registers)   take tickets and run the speculative order-preserving
//             application of the observations (EKF updates, new landmarks, weight), the sequential fallback
//             and the epilogue stores.
//
// Ticket hand-over.  Every screener owns TWO ticket slots and uses them alternately; each slot has a full and an
// empty mbarrier (one arrival each).  Applier aw serves the screeners aw, aw + AW, ...: a slot has exactly one
// producer and one consumer, each of which works through its slots in order and waits for the other side before
// reusing one, so a barrier is never more than one phase away from what a waiter asks for and the one-bit phase
// parity is unambiguous (a shared ring with free-running ticket counters is not: a slow screener can be lapped by
// a whole generation).  An applier looks at its screeners in turn with a non-blocking test of the full barrier.
#pragma once
#include "fs2_update.cuh"
#ifdef FS2_ASSERTS
#include <assert.h>
#define FS2_CHECK(c) assert(c)
#else
#define FS2_CHECK(c) ((void)0)
#endif

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
#ifndef FS2_WS_MINB
#define FS2_WS_MINB 2
#endif
#ifndef FS2_TCAP
#define FS2_TCAP 32        // landmarks a step may touch before its LAST round; beyond: the literal loop takes over
#endif
#ifndef FS2_TAIL
#define FS2_TAIL 32        // a map that ends with 1 .. FS2_TAIL landmarks past a full chunk: no ring round for them
#endif
#ifndef FS2_COMPACT_SCREEN
#define FS2_COMPACT_SCREEN DEFER      // one-landmark-at-a-time screening loop (smaller code) in the deferred launch only
#endif
#define FS2_TBW 64         // words of the touched-landmark bitmap: maps of up to 2048 landmarks re-speculate, larger ones take the literal loop
#define FS2_WS_THREADS ((FS2_SW + FS2_AW) * 32)
// setmaxnreg acts on warpgroups (4 consecutive warps, all with the same value): a role boundary inside a warpgroup
// hangs the kernel (seen with 5:3, 6:2 and 4:2 builds)
static_assert(FS2_SW % 4 == 0 && FS2_AW % 4 == 0, "screener and applier warps must come in multiples of four");
static_assert(FS2_SW % FS2_AW == 0, "every applier serves the same number of screeners");
#define FS2_NS (FS2_SW / FS2_AW)   // screeners per applier

struct Fs2Ticket {
    int4 ml[32];                 // per observation: its <= 4 lowest exact matches on the pre-step map
    double pose[1 + FS2_SIBMAX][5];   // pose after the motion step (x, y, yaw, sin yaw, cos yaw): [0] the ticket's particle,
                                      // [f] its f-th follower (DEFER)
    double pw;                   // weight before the step
    long long p;
    int cnt, slot;
    unsigned ovf;                // observations with more than 4 matches
    int nfol;                    // deferred-copy step: the next nfol particles share this ticket (same pre-step map)
    int fslot[8];                // ... on these map slots (copies written by the screener)
};

struct Fs2WsSmem {
    double ox[32], oy[32], zd[32], za[32];
    double sza[32], cza[32];     // sin / cos of the bearings (new landmarks by angle addition: see the appliers)
    alignas(8) float2 of[33];
    unsigned tab1[FS2_G1P * FS2_G1P];
    unsigned tab2[FS2_G2P * FS2_G2P];
    alignas(128) unsigned char ring[FS2_SW][FS2_NST][FS2_CHUNK_BYTES];
    alignas(8) unsigned long long bar[FS2_SW][FS2_NST];
    alignas(8) unsigned long long q_full[FS2_SW][2];
    alignas(8) unsigned long long q_empty[FS2_SW][2];
    int qidx[FS2_SW][FS2_QCAP];
    unsigned qmask[FS2_SW][FS2_QCAP];
    alignas(16) Fs2Ticket tk[FS2_SW][2];
    unsigned nper[FS2_SW];       // particles each screener will process
    unsigned ktaken[FS2_SW];     // tickets of each screener its applier has taken
    unsigned nlist;              // particles of this launch
    unsigned conf[FS2_AW];
    int bound[FS2_AW][32];
    alignas(16) Fs2Lm tlm[FS2_AW][FS2_TCAP];      // landmarks written by the rounds so far (multi-round steps only)
    unsigned tlater[FS2_AW][FS2_TCAP];            // ... the observations each of them could gate (level test of its new state)
    int tidx[FS2_AW][FS2_TCAP];
    unsigned tbits[FS2_AW][FS2_TBW];              // ... and a bitmap over landmark indices: touched this step (all zero between particles)
};

__device__ __forceinline__ void fs2_mbar_arrive(unsigned long long *bar)
{
    unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(b) : "memory");
}

// fs2_tma_load with shared-space addresses already in hand
__device__ __forceinline__ void fs2_tma_load_s(unsigned dst_saddr, const void *src, unsigned bytes, unsigned bar_saddr)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar_saddr), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(dst_saddr), "l"(src), "r"(bytes), "r"(bar_saddr) : "memory");
}

// one look at a barrier: has the phase of this parity completed?  (no waiting)
__device__ __forceinline__ bool fs2_mbar_test(unsigned bar_saddr, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar_saddr), "r"(parity) : "memory");
    return ok != 0u;
}

// wait at most ~ns for the phase, then report
__device__ __forceinline__ bool fs2_mbar_try(unsigned bar_saddr, unsigned parity, unsigned ns)
{
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar_saddr), "r"(parity), "r"(ns) : "memory");
    return ok != 0u;
}

// observations a landmark could gate (superset), straight from its mean and covariance: the level test described at
// the top of the file.  Everything fs2_box would give an infinite box (unsafe covariance, non-finite mean) or a box
// wider than e2 gets all observations; the exact test then decides alone.
// The usual landmark -- safe, updated at least once (level 1), within reach of the observations -- is settled by
// straight-line code (fs2_screen_fast); everything else is patched up afterwards behind one warp-wide branch.
struct Fs2Scr {
    unsigned cand;      // tab1 entry at the landmark's mean (meaningful only if usual)
    float x, y, big;
    bool safe, usual;
};

template <class SM>
__device__ __forceinline__ Fs2Scr fs2_screen_fast(const SM &sm, const Fs2ObsBatch &ob, const double2 &m, const double2 &c0,
                                                  const double2 &c1)
{
    Fs2Scr r;
    r.x = (float)m.x; r.y = (float)m.y;
    const float a = (float)c0.x, b = (float)c0.y, c = (float)c1.x, d = (float)c1.y;
    const float ad = a * d;
    const float det = fmaf(-b, c, ad);
    const float asym = b - c;
    r.big = fmaxf(a, d);
    // fs2_box's `safe` without its ad < 1e30 (implied by big <= amax2) and with the position test split off
    r.safe = (a > 0.f) & (ad > 1e-30f) & (det > 4e-6f * ad) & (asym * asym <= 1e-12f * ad) &
             (r.big * r.big < 1e4f * ad) & (r.big <= ob.amax2);
    r.usual = r.safe & (r.big <= ob.amax1) & (fmaxf(fabsf(r.x), fabsf(r.y)) < ob.xymax);
    r.cand = sm.tab1[fs2_cell<FS2_G1>(r.x, r.y, ob.inv_s1, ob.cx1, ob.cy1)];   // harmless when not usual (index clamped)
    return r;
}

template <class SM>
__device__ __forceinline__ unsigned fs2_screen_rare(const SM &sm, const Fs2ObsBatch &ob, const Fs2Scr &r)
{
    if (!r.safe) return ob.all_mask;
    if (!(fmaxf(fabsf(r.x), fabsf(r.y)) < ob.xymax)) return 0u;      // beyond every observation by more than e2 (NaN too)
    if (r.big <= ob.amax1) return r.cand;
    return sm.tab2[fs2_cell<FS2_G2>(r.x, r.y, ob.inv_s2, ob.cx2, ob.cy2)];
}

template <class SM>
__device__ __forceinline__ unsigned fs2_screen(const SM &sm, const Fs2ObsBatch &ob, const double2 &m, const double2 &c0,
                                               const double2 &c1)
{
    const Fs2Scr r = fs2_screen_fast(sm, ob, m, c0, c1);
    return r.usual ? r.cand : fs2_screen_rare(sm, ob, r);
}

// phase 2 of the screeners: box-filter and exact re-test of the queued (landmark, candidate observations) pairs
__device__ __forceinline__ void fs2_drain_ws(const Fs2WsSmem &sm, const int *qidx, const unsigned *qmask, int4 *ml,
                                             unsigned *ovf, int lane, const double *lm, int qn, float gate_f,
                                             float slack, double gate)
{
    for (int e = lane; e < qn; e += 32) {
        const int idx = qidx[e];
        const Fs2Lm l = fs2_load_lm(lm, idx);
        const Fs2Box bx = fs2_box(l.x, l.y, l.c00, l.c01, l.c10, l.c11, gate_f, slack);
        unsigned m = fs2_box_filter(sm, bx, qmask[e]);
        if (m) {
            const Fs2Gate g = fs2_gate_prepare(l.c00, l.c01, l.c10, l.c11);
            while (m) {
                const int k = __ffs(m) - 1;
                m &= m - 1;
                if (g.singular || fs2_gate_test(g, l.x, l.y, sm.ox[k], sm.oy[k], gate)) fs2_ml_insert(ml, ovf, k, idx);
            }
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------------
// DEFER: the launch that follows a resample whose map copies were deferred (fs2_step.cuh).  After a resample the
// offspring of one ancestor are consecutive particles with the SAME pose, weight and map; they differ only in the
// motion noise they draw next.  Association never looks at the pose (quirk Q1), so all of them have the same match
// lists.  The resampler therefore copies no map for them: it cuts every lineage into groups of a LEADER and up to
// FS2_SIBMAX FOLLOWERS (the leader owns a map: the first offspring keeps the ancestor's, every 8th sibling got a real
// copy), gives the followers a free map slot each and hands this kernel the list of leaders.  A screener streams and
// screens the leader's map once and, chunk by chunk as it sits in shared memory, writes it out again into the
// followers' slots (bulk TMA stores from the ring stage: the copy costs no second read of the map).  The ticket then
// serves the whole group: the applier runs the leader and, from the same match lists, every follower on its own copy
// with its own motion draw.  The copies are the map BEFORE this step's updates: they leave from shared memory, and the
// ticket is published only after they have landed (cp.async.bulk.wait_group), so no applier writes before that.
template <bool DEFER>
__device__ __forceinline__ void fs2_ws_screener(Fs2WsSmem &sm, const Fs2State &st, const Fs2ObsBatch &ob,
                                                const Fs2UpdateArgs &ua, int sw, int lane)
{
    const unsigned lt_mask = (1u << lane) - 1u;
    const int M = ob.M;
    const size_t map_bytes = (size_t)st.lcap * 48u;
    const unsigned step = gridDim.x * FS2_SW;            // (32-bit particle arithmetic: P < 2^31, fs2_create)
    const bool streaming = (ua.force_seq == 0) && (M > 0);
    unsigned char *ring = &sm.ring[sw][0][0];
    unsigned long long *bars = &sm.bar[sw][0];
    const unsigned bar_s0 = (unsigned)__cvta_generic_to_shared(bars);
    const unsigned qempty0 = (unsigned)__cvta_generic_to_shared(&sm.q_empty[sw][0]);
    const unsigned char *lm_base = reinterpret_cast<const unsigned char *>(st.lm);

    // The particle header is fetched ONE PARTICLE AHEAD, one field per lane (lane 0..4: x, y, yaw, w, noise as 8
    // bytes; lane 5, 6: count, slot as 4 bytes), so the pending loads occupy one register pair instead of twelve
    // registers that the 56-register budget would spill -- a spill store waits for the load it spills.
    const unsigned char *hptr = nullptr;
    {   // (a select chain, not an array indexed by the lane: that would live in local memory)
        const void *hp = nullptr;
        hp = (lane == 0) ? (const void *)st.x : hp;
        hp = (lane == 1) ? (const void *)st.y : hp;
        hp = (lane == 2) ? (const void *)st.yaw : hp;
        hp = (lane == 3) ? (const void *)st.w : hp;
        hp = (lane == 4 && ua.do_motion) ? (const void *)ua.noise : hp;
        hp = (lane == 5) ? (const void *)st.count : hp;
        hp = (lane == 6) ? (const void *)st.slot : hp;
        hptr = reinterpret_cast<const unsigned char *>(hp);
    }
    const bool sib_on = DEFER;
    unsigned long long hraw = 0ull;
    unsigned hcs = 0u;                   // lanes 5, 6: count, slot (their own register: widening a pending 4-byte load into hraw
                                         // would wait for it on the spot)
    unsigned hsib = 0u;                  // DEFER: lane 7: followers of the leader, lanes 8..14: their map slots
    auto hload = [&](unsigned q) {
        if (hptr) {
            if (lane < 5) hraw = *reinterpret_cast<const unsigned long long *>(hptr + 8 * (size_t)q);
            else hcs = *reinterpret_cast<const unsigned *>(hptr + 4 * (size_t)q);
        }
        if (sib_on) {
            if (lane == 7) hsib = (unsigned)ua.nfol[q];
            else if (lane >= 8 && lane < 8 + FS2_SIBMAX && q + (unsigned)(lane - 7) < (unsigned)st.P) hsib = (unsigned)st.slot[q + (unsigned)(lane - 7)];
            // the followers' motion draws ride in the header register of lanes 16 .. (lanes 0 .. 4 hold the leader's header)
            if (ua.do_motion && lane >= 16 && lane < 16 + FS2_SIBMAX && q + (unsigned)(lane - 15) < (unsigned)st.P)
                hraw = *reinterpret_cast<const unsigned long long *>(ua.noise + q + (unsigned)(lane - 15));
        }
    };
    // particles of this launch: all of them, or the leaders of a deferred-copy step
    const unsigned n = DEFER ? sm.nlist : (unsigned)st.P;
    const int32_t *plist = DEFER ? ua.leaders : nullptr;
    auto pid = [&](unsigned i) -> unsigned { return DEFER ? (unsigned)plist[i] : i; };
    unsigned gc = 0, gp = 0;             // ring chunks consumed / issued over the warp's life
    const unsigned char *isrc = nullptr; // next chunk of the map being issued
    int irem = 0;                        // landmarks of that map not yet issued
    const unsigned ring_s0 = (unsigned)__cvta_generic_to_shared(ring);
    auto issue_one = [&]() {
        const int nl = min(irem, FS2_CHUNK);
        const unsigned stg = gp % FS2_NST;
        if (lane == 0) {
            // follower copies still reading this stage out of shared memory (issued a round ago) must be through with it
            if (sib_on) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
            fs2_tma_load_s(ring_s0 + stg * FS2_CHUNK_BYTES, isrc, (unsigned)nl * 48u, bar_s0 + 8u * stg);
        }
        isrc += FS2_CHUNK_BYTES;
        irem -= nl;
        ++gp;
    };
    // A map of n landmarks goes through the ring in chunks of FS2_CHUNK -- except a short end: 1 .. FS2_TAIL landmarks
    // past at least one full chunk are read straight from global memory after the last round (prefetched into L2 when
    // the particle starts).  Maps grow by a landmark now and then, and a whole ring round for two landmarks cost 11-17 %
    // of the kernel; a tail buffer filled by the TMA cost even more (registers in the hot loop).
    auto ring_part = [](int n) {
        const int nfull = n / FS2_CHUNK, ntail = n - nfull * FS2_CHUNK;
        return (FS2_TAIL > 0 && nfull >= 1 && ntail >= 1 && ntail <= FS2_TAIL) ? nfull * FS2_CHUNK : n;
    };
    // consecutive particles go to different CTAs: after a resample they are siblings, and a lineage that is slow to apply
    // (exhausted match lists, many rounds) would otherwise queue up behind the four appliers of one CTA
    unsigned idx = (unsigned)sw * gridDim.x + blockIdx.x;
    unsigned p = (idx < n) ? pid(idx) : 0u;
    unsigned p_next = (idx + step < n) ? pid(idx + step) : 0u;     // a list entry is needed one particle before its header
    if (idx < n) hload(p);
    int cnt_cur = (int)__shfl_sync(FS2_FULL, hcs, 5);
    int slot_cur = (int)__shfl_sync(FS2_FULL, hcs, 6);
    if (streaming && idx < n) {
        isrc = lm_base + (size_t)slot_cur * map_bytes;
        irem = ring_part(cnt_cur);
        for (int c = 0; c < FS2_NST - 1 && irem > 0; ++c) issue_one();
    }
    for (unsigned k = 0; idx < n; idx += step, ++k) {
        const int cnt = cnt_cur;
        FS2_CHECK(p < (unsigned)st.P);
        FS2_CHECK(cnt >= 0 && cnt <= st.lcap);
        FS2_CHECK(slot_cur >= 0 && (int64_t)slot_cur < 2 * st.P);
        const unsigned cursib = hsib;            // (the next particle's header load overwrites hsib below)
        const int nsib = sib_on ? (int)__shfl_sync(FS2_FULL, cursib, 7) : 0;
        FS2_CHECK(nsib >= 0 && nsib <= FS2_SIBMAX && p + nsib < (unsigned)st.P);
        const double *lm = reinterpret_cast<const double *>(lm_base + (size_t)slot_cur * map_bytes);
        double mx = __longlong_as_double((long long)__shfl_sync(FS2_FULL, hraw, 0));
        double my = __longlong_as_double((long long)__shfl_sync(FS2_FULL, hraw, 1));
        double myaw = __longlong_as_double((long long)__shfl_sync(FS2_FULL, hraw, 2));
        const double pw = __longlong_as_double((long long)__shfl_sync(FS2_FULL, hraw, 3));
        const unsigned long long curh = hraw;    // (lane 4: the particle's motion draw, lanes 16 ..: its followers')
        // ---- next particle's header, one particle ahead (and the list entry of the one after it) ----
        const bool have_next = idx + step < n;
        const unsigned pn = p_next;
        if (have_next) hload(pn);
        if (idx + 2 * step < n) p_next = pid(idx + 2 * step);
        // ---- my ticket slot of this turn: wait until its previous use has been taken over by an applier ----
        const unsigned j = k & 1u;
        // (a blocked try_wait is woken by every barrier event of the CTA -- over a hundred times per particle; sleep
        // in earnest instead: the appliers that would free the slot share this scheduler)
        while (!fs2_mbar_test(qempty0 + 8u * j, ((k >> 1) & 1u) ^ 1u)) __nanosleep(1500);
        Fs2Ticket &tk = sm.tk[sw][j];
        tk.ml[lane] = make_int4(FS2_NONE, FS2_NONE, FS2_NONE, FS2_NONE);
        if (DEFER) {
            if (lane == 0) tk.nfol = nsib;
            if (lane >= 8 && lane < 8 + FS2_SIBMAX) tk.fslot[lane - 8] = (int)cursib;
        }
        // __move_particle (fast_slam_2.py:69-87) happens here: association does not look at the pose (quirk Q1).
        // All lanes compute it, lane 0 stores.  On a deferred-copy step the particle's followers start from the same pose
        // and differ in their draw only: lane f computes follower f's move in the same instructions.
        {
            double fx = mx, fy = my, fyaw = myaw;
            const int from = (DEFER && lane >= 1 && lane <= FS2_SIBMAX) ? 15 + lane : 4;
            const double nz = __longlong_as_double((long long)__shfl_sync(FS2_FULL, curh, from));
            double fs, fc;
            if (ua.do_motion) fs2_move(fx, fy, fyaw, ua.rotation, ua.translation, nz, &fs, &fc);
            else sincospi(fyaw * 0.31830988618379067154, &fs, &fc);
            if (lane <= nsib) {
                if (ua.do_motion) { st.x[p + lane] = fx; st.y[p + lane] = fy; st.yaw[p + lane] = fyaw; }
                tk.pose[lane][0] = fx; tk.pose[lane][1] = fy; tk.pose[lane][2] = fyaw;
                tk.pose[lane][3] = fs; tk.pose[lane][4] = fc;
            }
        }
        if (lane == 0) {
            tk.pw = pw;
            tk.p = (long long)p; tk.cnt = cnt; tk.slot = slot_cur; tk.ovf = 0u;
        }
        __syncwarp();
        if (streaming) {
            const int nring = ring_part(cnt);                 // landmarks that come through the ring
            const int nchunks = (nring + FS2_CHUNK - 1) / FS2_CHUNK;
            if (nring < cnt && nring + lane < cnt) {          // the short end: on its way into L2 while the ring is worked through
                const unsigned char *t = reinterpret_cast<const unsigned char *>(lm) + (size_t)(nring + lane) * 48u;
                asm volatile("prefetch.global.L2 [%0];\n" ::"l"(t));
                asm volatile("prefetch.global.L2 [%0];\n" ::"l"(t + 32));
            }
            int qn = 0;
            for (int c = 0; c < nchunks; ++c) {
                if (irem > 0) issue_one();            // keep NST-1 chunks in flight
                const unsigned stage = gc % FS2_NST;
                fs2_mbar_wait(bar_s0 + 8u * stage, (gc / FS2_NST) & 1u);
                ++gc;
                if (DEFER && nsib > 0) {      // this chunk, as it sits in shared memory, into every follower's map slot
                    const unsigned nbytes = (unsigned)min(FS2_CHUNK, nring - c * FS2_CHUNK) * 48u;
                    for (int jj = 0; jj < nsib; ++jj) {
                        const unsigned sslot = __shfl_sync(FS2_FULL, cursib, 8 + jj);
                        FS2_CHECK((int64_t)sslot < 2 * st.P);
                        if (lane == 0) {
                            unsigned char *dst = const_cast<unsigned char *>(lm_base) + (size_t)sslot * map_bytes + (size_t)c * FS2_CHUNK_BYTES;
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
                                         ::"l"(dst), "r"(ring_s0 + stage * FS2_CHUNK_BYTES), "r"(nbytes) : "memory");
                        }
                    }
                    if (lane == 0) asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                }
                const int iA = c * FS2_CHUNK + lane, iB = iA + 32;
                const double2 *srcA = reinterpret_cast<const double2 *>(ring + stage * FS2_CHUNK_BYTES + 48 * lane);
                const double2 *srcB = srcA + 96;    // 32 landmarks * 48 B / 16 B
                unsigned candA = 0u, candB = 0u;
                if (FS2_COMPACT_SCREEN) {
                    // Compact form: one landmark at a time through one copy of the screen.  The deferred launch is bound by
                    // its appliers (its screeners sleep a quarter of the time), and with the copy code its hot instructions
                    // were 33 KB -- just over the 32 KB instruction cache: the appliers then waited for instructions a third
                    // of their time (ncu, stall_no_inst: twice the plain launch's).
#pragma unroll 1
                    for (int h = 0; h < 2; ++h) {
                        const int ih = iA + 32 * h;
                        const double2 *src = srcA + 96 * h;
                        unsigned cand = 0u;
                        if (ih < cnt) cand = fs2_screen(sm, ob, src[0], src[1], src[2]);
                        const unsigned has = __ballot_sync(FS2_FULL, cand != 0u);
                        if (cand) {
                            const int pos = qn + __popc(has & lt_mask);
                            sm.qidx[sw][pos] = ih;
                            sm.qmask[sw][pos] = cand;
                        }
                        qn += __popc(has);
                    }
                } else if (c * FS2_CHUNK + FS2_CHUNK <= cnt) {
                    // a full chunk (warp-uniform): two landmarks per lane as two independent branch-free streams the
                    // compiler interleaves; the rare non-usual landmark is patched up behind one vote
                    const Fs2Scr ra = fs2_screen_fast(sm, ob, srcA[0], srcA[1], srcA[2]);
                    const Fs2Scr rb = fs2_screen_fast(sm, ob, srcB[0], srcB[1], srcB[2]);
                    candA = ra.cand; candB = rb.cand;
                    if (__any_sync(FS2_FULL, !(ra.usual & rb.usual))) {
                        if (!ra.usual) candA = fs2_screen_rare(sm, ob, ra);
                        if (!rb.usual) candB = fs2_screen_rare(sm, ob, rb);
                    }
                } else {
                    if (iA < cnt) candA = fs2_screen(sm, ob, srcA[0], srcA[1], srcA[2]);
                    if (iB < cnt) candB = fs2_screen(sm, ob, srcB[0], srcB[1], srcB[2]);
                }
                const unsigned hasA = FS2_COMPACT_SCREEN ? 0u : __ballot_sync(FS2_FULL, candA != 0u);
                const unsigned hasB = FS2_COMPACT_SCREEN ? 0u : __ballot_sync(FS2_FULL, candB != 0u);
                if (FS2_COMPACT_SCREEN) {
                    if (qn > FS2_QCAP - 96) {
                        __syncwarp();
                        fs2_drain_ws(sm, sm.qidx[sw], sm.qmask[sw], tk.ml, &tk.ovf, lane, lm, qn, ua.gate_f, ob.slack, ua.gate);
                        qn = 0;
                    }
                } else if (hasA | hasB) {
                    // queue order must stay ascending in landmark index: all of A (iA < iB) first
                    if (candA) {
                        const int pos = qn + __popc(hasA & lt_mask);
                        sm.qidx[sw][pos] = iA;
                        sm.qmask[sw][pos] = candA;
                    }
                    qn += __popc(hasA);
                    if (candB) {
                        const int pos = qn + __popc(hasB & lt_mask);
                        sm.qidx[sw][pos] = iB;
                        sm.qmask[sw][pos] = candB;
                    }
                    qn += __popc(hasB);
                    if (qn > FS2_QCAP - 96) {
                        __syncwarp();
                        fs2_drain_ws(sm, sm.qidx[sw], sm.qmask[sw], tk.ml, &tk.ovf, lane, lm, qn, ua.gate_f, ob.slack, ua.gate);
                        qn = 0;
                    }
                }
                __syncwarp();   // every lane is done with this stage before lane 0 hands it back to the TMA
            }
            cnt_cur = (int)__shfl_sync(FS2_FULL, hcs, 5);
            slot_cur = (int)__shfl_sync(FS2_FULL, hcs, 6);
            if (have_next) {   // ring empty: start on the next particle's map
                isrc = lm_base + (size_t)slot_cur * map_bytes;
                irem = ring_part(cnt_cur);
                for (int c = 0; c < FS2_NST - 1 && irem > 0; ++c) issue_one();
            }
            if (nring < cnt) {   // the short end of this map, straight from global memory (L2 by now)
                unsigned candT = 0u;
                const bool mine = nring + lane < cnt;
                double2 g0 = make_double2(0.0, 0.0), g1 = g0, g2 = g0;
                if (mine) {
                    const double2 *g = reinterpret_cast<const double2 *>(lm + 6 * (size_t)(nring + lane));
                    g0 = g[0]; g1 = g[1]; g2 = g[2];
                    candT = fs2_screen(sm, ob, g0, g1, g2);
                }
                if (DEFER) {
                    for (int jj = 0; jj < nsib; ++jj) {           // the followers' copies of the short end, by plain stores
                        const unsigned sslot = __shfl_sync(FS2_FULL, cursib, 8 + jj);
                        if (mine) {
                            double2 *d = reinterpret_cast<double2 *>(const_cast<unsigned char *>(lm_base) + (size_t)sslot * map_bytes) + 3 * (size_t)(nring + lane);
                            d[0] = g0; d[1] = g1; d[2] = g2;
                        }
                    }
                }
                const unsigned hasT = __ballot_sync(FS2_FULL, candT != 0u);
                if (candT) {
                    const int pos = qn + __popc(hasT & lt_mask);
                    sm.qidx[sw][pos] = nring + lane;
                    sm.qmask[sw][pos] = candT;
                }
                qn += __popc(hasT);
                __syncwarp();
            }
            fs2_drain_ws(sm, sm.qidx[sw], sm.qmask[sw], tk.ml, &tk.ovf, lane, lm, qn, ua.gate_f, ob.slack, ua.gate);
        } else {
            cnt_cur = (int)__shfl_sync(FS2_FULL, hcs, 5);
            slot_cur = (int)__shfl_sync(FS2_FULL, hcs, 6);
        }
        __syncwarp();
        if (lane == 0) {
            // the followers' copies have landed before their applier reads or writes them
            if (DEFER && nsib > 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
            fs2_mbar_arrive(&sm.q_full[sw][j]);   // publish the ticket
        }
        p = pn;
    }
}

// The reference's loop, observation by observation (landmark_utils.py:103-117 + fast_slam_2.py:100-159): the fallback
// of the appliers for exhausted match lists and FS2_FLAG_FORCE_SEQUENTIAL.  (Calling it out of line was measured: the
// appliers' hot loop shrinks by 40 % of its instructions but the calls cost more than the shorter code saves.)
struct Fs2SeqOut {
    double pw;
    int cnt, stat, my_assoc;
};

__device__ __forceinline__ Fs2SeqOut fs2_apply_sequential(const double *s_ox, const double *s_oy, const double *s_zd, const double *s_za,
                                                       double *lm, int lcap, double r00, double r01, double r10, double r11,
                                                       double gate, double px, double py, double pyaw, double pw, int cnt, int ks,
                                                       int M, int lane, int stat, int my_assoc)
{
    for (int k = ks; k < M; ++k) {
        const double kox = s_ox[k], koy = s_oy[k];
        int found = FS2_NONE;
        for (int base = 0; base < cnt && found == FS2_NONE; base += 32) {
            int i = base + lane;
            bool stop = false;
            if (i < cnt) stop = fs2_stops_here(fs2_load_lm(lm, i), kox, koy, gate);
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
                st_k = fs2_ekf(px, py, pyaw, s_zd[k], s_za[k], r00, r01, r10, r11, in, &post, &like);
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
                Fs2Lm post = fs2_new_landmark(px, py, pyaw, s_zd[k], s_za[k]);
                if (lane == 0) fs2_store_lm(lm, cnt, post);
                cnt += 1;
            } else st_k = 8;
        }
        stat |= st_k;
        if (lane == k) my_assoc = res;
        __syncwarp();
    }
    Fs2SeqOut o;
    o.pw = pw; o.cnt = cnt; o.stat = stat; o.my_assoc = my_assoc;
    return o;
}

// ------------------------------------------------------------------------------------------------------
template <bool DEFER>
__device__ __forceinline__ void fs2_ws_applier(Fs2WsSmem &sm, const Fs2State &st, const Fs2ObsBatch &ob,
                                               const Fs2UpdateArgs &ua, int aw, int lane)
{
    const unsigned lt_mask = (1u << lane) - 1u;
    const int M = ob.M;
    const int lcap = st.lcap;
    const bool is_obs = lane < M;
    const double zd = sm.zd[lane], za = sm.za[lane];
    // (the observation's map-frame point is needed on multi-round steps only: read from shared memory there, the
    // appliers have no register to spare)
    const unsigned qfull0 = (unsigned)__cvta_generic_to_shared(&sm.q_full[0][0]);

    // software-pipelined ticket fetch: the NEXT ticket is claimed and its first-match landmark load issued before
    // the current particle is processed -- if a ticket is ready by then.  The rest of the ticket stays in shared
    // memory until its turn (a screener has two slots and needs one back only a whole particle later).
    // DEFER: a ticket serves its leader (f = 0) and then the followers f = 1 .. nf, one work item each.
    struct Held { int s, j, ml0; Fs2Lm in; bool valid; int f, nf; };
    // Applier aw serves the screeners aw, aw + AW, ... (FS2_NS of them): every ticket slot has exactly one producer and
    // one consumer, so the consumed counts live in registers and nothing has to be claimed.
    // (the counts live in shared memory -- sm.ktaken, sm.nper -- and are touched once per ticket)
    int pref = 0;                        // the screener looked at first (alternates: neither of them starves)
    // take one published ticket.  Returns false if none is ready right now and the caller does not want to wait
    // (h untouched); sets h.valid = false and returns true when all tickets of this applier's screeners are taken.
    auto take = [&](Held &h, bool blocking) -> bool {
        for (;;) {
            bool left = false;
#pragma unroll
            for (int t = 0; t < FS2_NS; ++t) {
                int i = pref + t;
                if (i >= FS2_NS) i -= FS2_NS;
                const int s = aw + i * FS2_AW;
                const unsigned k = sm.ktaken[s], np = sm.nper[s];
                if (k >= np) continue;
                left = true;
                const unsigned j = k & 1u;
                // ONE decision for the warp: with fewer than 32 observations the lanes beyond them skip the landmark loads
                // and can run a few instructions apart; lanes that saw the barrier at different moments would take the
                // ticket on different turns, and the late ones would read it after lane 0 had handed the slot back.
                // (A published ticket stays published until this warp takes it, so "any lane saw it" is exact.)
                if (!__any_sync(FS2_FULL, fs2_mbar_test(qfull0 + 8u * (2u * s + j), (k >> 1) & 1u))) continue;
                __syncwarp();
                if (lane == 0) sm.ktaken[s] = k + 1u;
                __syncwarp();
                const Fs2Ticket &tk = sm.tk[s][j];
                h.s = s; h.j = (int)j;
                h.ml0 = tk.ml[lane].x;
                h.in.x = h.in.y = h.in.c00 = h.in.c01 = h.in.c10 = h.in.c11 = 0.0;
                if (lane < M && h.ml0 != FS2_NONE)                   // first round's landmark, needed ~a particle later
                    h.in = fs2_load_lm(st.lm + (size_t)tk.slot * 6 * (size_t)lcap, h.ml0);
                h.valid = true;
                h.f = 0; h.nf = DEFER ? tk.nfol : 0;
                pref = (i + 1 < FS2_NS) ? i + 1 : 0;
                return true;
            }
            if (!left) { h.valid = false; return true; }
            if (!blocking) return false;
            // nothing published: sleep on the full barrier of the preferred screener if it still has tickets to come
            // (bounded, so that a ticket published by the other one is not left waiting), then look again
            const unsigned k = sm.ktaken[aw + pref * FS2_AW], np = sm.nper[aw + pref * FS2_AW];
            if (k < np) (void)fs2_mbar_try(qfull0 + 8u * (2u * (aw + pref * FS2_AW) + (k & 1u)), (k >> 1) & 1u, 400u);
            pref = (pref + 1 < FS2_NS) ? pref + 1 : 0;
        }
    };
    // one call site (the look is ~60 instructions): wait for a ticket only when there is nothing to work on,
    // otherwise take the next one if it happens to be ready (its first landmark then loads while this one runs)
    Held cur, nxt;
    cur.valid = false;
    bool done = false;
    for (;;) {
        nxt.valid = false;
        if (DEFER && cur.valid && cur.f < cur.nf) {
            // the next follower of the ticket in hand: same match lists, its own map copy and motion draw
            const Fs2Ticket &tk = sm.tk[cur.s][cur.j];
            nxt.s = cur.s; nxt.j = cur.j; nxt.ml0 = cur.ml0; nxt.f = cur.f + 1; nxt.nf = cur.nf;
            nxt.in.x = nxt.in.y = nxt.in.c00 = nxt.in.c01 = nxt.in.c10 = nxt.in.c11 = 0.0;
            if (lane < M && nxt.ml0 != FS2_NONE)
                nxt.in = fs2_load_lm(st.lm + (size_t)tk.fslot[cur.f] * 6 * (size_t)lcap, nxt.ml0);
            nxt.valid = true;
        } else if (!done) {
            const bool got = take(nxt, !cur.valid);
            if (got && !nxt.valid) done = true;
        }
        if (!cur.valid) {
            if (done) break;
            cur = nxt;
            continue;
        }
        const Fs2Ticket &ctk = sm.tk[cur.s][cur.j];
        int p = (int)ctk.p;                          // P < 2^31 (fs2_create)
        const int fi = DEFER ? cur.f : 0;            // a follower: the leader's pre-step state, its own pose and map copy
        p += fi;
        const double px = ctk.pose[fi][0], py = ctk.pose[fi][1], pyaw = ctk.pose[fi][2];
        const double psy = ctk.pose[fi][3], pcy = ctk.pose[fi][4];
        double pw = ctk.pw;
        int cnt = ctk.cnt;
        const int myslot = (DEFER && fi > 0) ? ctk.fslot[fi - 1] : ctk.slot;
        FS2_CHECK(fi >= 0 && fi <= FS2_SIBMAX && (!DEFER || fi <= ctk.nfol));
        FS2_CHECK(p >= 0 && p < st.P);
        FS2_CHECK(myslot >= 0 && (int64_t)myslot < 2 * st.P);
        FS2_CHECK(cnt >= 0 && cnt <= lcap);
        FS2_CHECK(cur.ml0 == ctk.ml[lane].x);
        double *lm = st.lm + (size_t)myslot * 6 * (size_t)lcap;
        const int4 ml = ctk.ml[lane];
        const bool ml_overflow = (ctk.ovf >> lane) & 1u;
        const Fs2Lm in0 = cur.in;
        __syncwarp();
        // everything is in registers: hand the slot back (once its last follower is under way)
        if (lane == 0 && (!DEFER || cur.f == cur.nf)) fs2_mbar_arrive(&sm.q_empty[cur.s][cur.j]);

        int stat = 0;
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
                            const bool touched = (sm.tbits[aw][cand >> 5] >> (cand & 31)) & 1u;
                            if (!touched) a_un = cand;
                        }
                    }
                    if (a_un == FS2_NONE && ml_overflow) exhausted = true;
                }
            }
            int a_t = FS2_NONE, a_t_pos = -1;
            if (active) {
                for (int t = 0; t < nt; ++t) {          // (the level test of the new state said whom it could gate: usually nobody)
                    const int ti = sm.tidx[aw][t];
                    if (((sm.tlater[aw][t] >> lane) & 1u) && ti < a_t) {
                        if (fs2_stops_here(sm.tlm[aw][t], sm.ox[lane], sm.oy[lane], ua.gate)) { a_t = ti; a_t_pos = t; }
                    }
                }
            }
            if (__any_sync(FS2_FULL, exhausted)) { seq = true; break; }
            const bool from_t = a_t < a_un;
            const int a = from_t ? a_t : a_un;
            const bool matched = active && a != FS2_NONE;
            const unsigned unm = __ballot_sync(FS2_FULL, active && !matched);
            const int app_rank = __popc(unm & lt_mask);

            // Every lane computes BOTH outcomes -- the EKF update of a (possibly dummy) landmark and a new landmark -- and
            // selects.  A warp that splits at an if / else here was seen to stay split through the votes and shuffles that
            // follow (each one then takes the compiler's slow "collective" path: +60 % instructions on a step where a
            // single observation was new, and lanes running a turn apart when fewer than 32 hold an observation).
            Fs2Lm in = in0;                                      // first round: loaded with the ticket (zeros without a match)
            if (nt != 0) {                                        // (warp-uniform) later rounds: the map, or the touched table
                // an observation that stays with its first-round landmark still has it in registers (in0); the map is read
                // again only if some lane moved on to a later entry of its match list
                const bool moved = matched && !from_t && a != ml.x;
                if (__any_sync(FS2_FULL, moved)) {
                    const Fs2Lm g = fs2_load_lm(lm, moved ? a : 0);
                    if (moved) in = g;
                }
                const Fs2Lm tl = sm.tlm[aw][from_t ? a_t_pos : 0];
                if (from_t) in = tl;
            }
            // dummy input of a lane without a match (result unused): nothing in it may send the warp down a slow path -- a
            // zero numerator in atan2's division did (the library's division subroutine, on every particle with a new landmark)
            // (nor a Mahalanobis distance beyond 1390, where the likelihood leaves the short exp: hence the huge covariance)
            if (!matched) { in.x = px + 1.0; in.y = py + 0.75; in.c00 = 1e4; in.c01 = 0.0; in.c10 = 0.0; in.c11 = 1e4; }
            const bool sing = matched && __dadd_rn(__dmul_rn(in.c00, in.c11), -__dmul_rn(in.c01, in.c10)) == 0.0;
            Fs2Lm post;
            double like;
            const int st_e = fs2_ekf(px, py, pyaw, zd, za, ua.r00, ua.r01, ua.r10, ua.r11, in, &post, &like);
            const bool ekf_ok = matched && !sing && st_e != 2;
            const bool app_ok = active && !matched && (cnt + app_rank < lcap);
            int st_k = 0, res = -3;
            if (matched) { st_k = sing ? 1 : st_e; res = (sing || st_e == 2) ? -2 : a; }
            else if (active) { res = -1; st_k = app_ok ? 0 : 8; }
            const int widx = ekf_ok ? a : (app_ok ? cnt + app_rank : FS2_NONE);
            if (!ekf_ok) {
                // fs2_new_landmark (fast_slam_2.py:108-111) with cos / sin (yaw + bearing) by angle addition: sin / cos of
                // the yaw come with the ticket (the motion step computed them), those of the bearing with the batch
                const double sz = sm.sza[lane], cz = sm.cza[lane];
                const double cs = fma(pcy, cz, -(psy * sz)), sn = fma(psy, cz, pcy * sz);
                post.x = __dadd_rn(px, __dmul_rn(zd, cs));
                post.y = __dadd_rn(py, __dmul_rn(zd, sn));
                post.c00 = 0.1; post.c01 = 0.0; post.c10 = 0.0; post.c11 = 0.1;
                like = 1.0;
            }
            // same landmark as an earlier observation of the round (only matched observations can share one: new
            // landmarks get distinct slots)
            const unsigned act_mask = __ballot_sync(FS2_FULL, active);
            const unsigned same = __match_any_sync(FS2_FULL, matched ? a : (0x40000000 | lane));
            sm.bound[aw][lane] = matched ? a : FS2_NONE;
            // same landmark as an earlier observation of the round: one ballot, no shared-memory atomic
            const unsigned cf_same = __ballot_sync(FS2_FULL, matched && (same & lt_mask & act_mask) != 0u);
            if (lane == 0) sm.conf[aw] = cf_same;
            __syncwarp();
            unsigned lvl;
            // my post-state landmark, at a lower index than a LATER observation's choice, may stop its scan: the level
            // test of the screeners (no box) says which later observations it could gate at all -- usually none
            {
                lvl = fs2_screen(sm, ob, make_double2(post.x, post.y), make_double2(post.c00, post.c01), make_double2(post.c10, post.c11));
                unsigned later = lvl & act_mask & ~(lt_mask | (1u << lane));
                if (widx == FS2_NONE) later = 0u;
                if (!matched) later &= unm;      // a new landmark sits beyond every matched index: only unmatched observations can meet it
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
                const bool tw = commit && widx != FS2_NONE;
                const bool was = tw && lcap <= 32 * FS2_TBW && ((sm.tbits[aw][widx >> 5] >> (widx & 31)) & 1u);
                if (was) {                                   // touched before (the same landmark twice): its entry is replaced
                    for (int t = 0; t < nt; ++t) if (sm.tidx[aw][t] == widx) tpos = t;
                }
                const unsigned newt = __ballot_sync(FS2_FULL, tw && tpos < 0);
                if (nt + __popc(newt) > FS2_TCAP || lcap > 32 * FS2_TBW) {
                    // more touched landmarks than the table holds: what is committed stands, the reference's own loop
                    // does the rest against the map in global memory
                    seq = true;
                } else if (commit && widx != FS2_NONE) {
                    if (tpos < 0) tpos = nt + __popc(newt & lt_mask);
                    // another round follows: it meets the touched landmarks through the level test of their new state
                    sm.tidx[aw][tpos] = widx;
                    sm.tlm[aw][tpos] = post;
                    sm.tlater[aw][tpos] = lvl;
                    atomicOr(&sm.tbits[aw][widx >> 5], 1u << (widx & 31));
                }
                if (!seq) nt += __popc(newt);
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

        if (nt > 0) {                   // the bitmap goes back to all zero
            __syncwarp();
            if (lane < nt) sm.tbits[aw][sm.tidx[aw][lane] >> 5] = 0u;
            __syncwarp();
        }
        if (M > 0 && ks < M && seq) {   // the reference's loop, observation by observation 
            __syncwarp();
            const Fs2SeqOut so = fs2_apply_sequential(sm.ox, sm.oy, sm.zd, sm.za, lm, lcap, ua.r00, ua.r01, ua.r10, ua.r11, ua.gate,
                                                      px, py, pyaw, pw, cnt, ks, M, lane, stat, my_assoc);
            pw = so.pw; cnt = so.cnt; stat = so.stat; my_assoc = so.my_assoc;
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
        if (done && !cur.valid) break;
    }
}

template <bool DEFER>
__global__ void __launch_bounds__(FS2_WS_THREADS, FS2_WS_MINB)
fs2_update_ws_kernel(const Fs2State st, const __grid_constant__ Fs2ObsBatch ob, const Fs2UpdateArgs ua)
{
    // A step whose resample deferred its map copies is run by the DEFER form over the leaders, every other step by the
    // plain kernel; which of the two it is stands in device memory (the resample decision is taken on the device), so
    // both are launched and the one that is not needed returns here.
    const int pending = ua.dctl ? ((volatile const int32_t *)ua.dctl)[0] : 0;
    if (DEFER != (pending != 0)) return;
    const int64_t nl = DEFER ? (int64_t)((volatile const int32_t *)ua.dctl)[1] : st.P;
    extern __shared__ __align__(128) unsigned char fs2_smem_raw[];
    Fs2WsSmem &sm = *reinterpret_cast<Fs2WsSmem *>(fs2_smem_raw);
    // (the broadcast tells the compiler that the warp index is warp-uniform: everything derived from it -- ring, barrier and
    // ticket addresses, the particle counter -- can live in uniform registers instead of the role's small register budget)
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (threadIdx.x < 32) {
        sm.ox[lane] = ob.ox[lane]; sm.oy[lane] = ob.oy[lane];
        sm.zd[lane] = ob.zd[lane]; sm.za[lane] = ob.za[lane];
        sm.sza[lane] = ob.sza[lane]; sm.cza[lane] = ob.cza[lane];
        sm.of[lane] = make_float2(ob.oxf[lane], ob.oyf[lane]);
        if (lane == 0) { sm.of[32] = make_float2(__int_as_float(0x7f800000), __int_as_float(0x7f800000)); sm.nlist = (unsigned)nl; }
        if (lane < FS2_SW) {
            // particles of this CTA: screener w takes entries w * gridDim + blockIdx, + k * gridDim * SW of the launch's list
            const int64_t step = (int64_t)gridDim.x * FS2_SW;
            const int64_t p0 = (int64_t)lane * gridDim.x + blockIdx.x;
            sm.nper[lane] = (p0 < nl) ? (unsigned)((nl - p0 + step - 1) / step) : 0u;
            sm.ktaken[lane] = 0u;
            fs2_mbar_init(&sm.q_full[lane][0], 1); fs2_mbar_init(&sm.q_full[lane][1], 1);
            fs2_mbar_init(&sm.q_empty[lane][0], 1); fs2_mbar_init(&sm.q_empty[lane][1], 1);
        }
    }
    for (int i = threadIdx.x; i < FS2_G1P * FS2_G1P; i += blockDim.x) sm.tab1[i] = ob.tab1[i];
    for (int i = threadIdx.x; i < FS2_G2P * FS2_G2P; i += blockDim.x) sm.tab2[i] = ob.tab2[i];
    for (int i = threadIdx.x; i < FS2_AW * FS2_TBW; i += blockDim.x) (&sm.tbits[0][0])[i] = 0u;
    if (warp < FS2_SW && lane == 0) {
#pragma unroll
        for (int s = 0; s < FS2_NST; ++s) fs2_mbar_init(&sm.bar[warp][s], 1);
    }
    fs2_fence_mbar_init();
    __syncthreads();
    if (warp < FS2_SW) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(FS2_S_REGS));
        fs2_ws_screener<DEFER>(sm, st, ob, ua, warp, lane);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(FS2_A_REGS));
        // (the appliers get their index WITHOUT the uniformity hint: with it the compiler keeps the ticket in hand -- slot,
        // follower index -- in uniform registers across the EKF; when fewer than 32 lanes hold an observation the idle
        // lanes were seen a turn ahead of the busy ones (cuda-gdb), reading those registers while the others' code had
        // reused them.  Plain registers are private to a lane.)
        fs2_ws_applier<DEFER>(sm, st, ob, ua, (int)(threadIdx.x >> 5) - FS2_SW, lane);
    }
}
