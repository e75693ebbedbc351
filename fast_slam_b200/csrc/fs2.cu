// fs2.cu -- C ABI (include/fs2.h) of the B200 FastSLAM filter-step library.  Host orchestration only;
// the kernels live in fs2_update.cuh, fs2_weights.cuh, fs2_resample.cuh.  Compiled for sm_100a.
#include "../../include/fs2.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fs2_update.cuh"
#include "fs2_update_ws.cuh"
#include "fs2_weights.cuh"
#include "fs2_resample.cuh"
#include "fs2_step.cuh"
#include "fs2_place.cuh"
#include "fs2_frontend.cuh"
#include "fs2_known.cuh"
#include "fs2_icp.cuh"

static thread_local char g_cuda_err[512] = "";

#define FS2_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            snprintf(g_cuda_err, sizeof(g_cuda_err), "%s:%d %s: %s", __FILE__, __LINE__, #call, \
                     cudaGetErrorString(e__));                                                  \
            return FS2_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

struct fs2_filter_s {
    fs2_config cfg;
    int64_t P, Pglobal;
    int64_t S;                // map slots in the pool (P + spare)
    int lcap;
    int sm_count;
    // state
    double *x, *y, *yaw, *w;
    int32_t *count, *slot, *status;
    double *lm;
    // scratch
    double *noise;
    double *x2, *y2, *yaw2, *w2;
    int32_t *count2, *slot2;
    int32_t *alive, *extra, *tasks, *freeslot, *ncopies;
    int2 *iscan_bs;
    int iscan_nb;
    double *stats;            // FS2_STATS_LEN
    double *partial;          // reductions
    double *partial_sq;
    Fs2MaxIdx *partial_best;
    unsigned int *counters;   // [2]
    // resample scan (sized for Pglobal)
    double *cumsum, *bsum, *bpre, *cstart, *scan_total;
    unsigned long long *A0, *A1;
    int *eb, *mode, *anomaly, *stuck;
    unsigned long long *gA0, *gA1;   // per group of FS2_GRP scan blocks
    int *gE, *gK, *gV;
    double *cg;
    int scan_nb, scan_ng;
    int32_t *ancestor;
    // decoupled placement of a sharded filter (fs2_place.cuh): logical ids of the local particles, plan scratch
    int32_t *logi, *logi_new, *src, *ancl;
    PlPlan pl;
    void *pl_block;            // one allocation behind pl's arrays
    int pl_on;
    // fused step (fs2_step.cuh): control words, one-pass scan state
    int *ctl;
    unsigned long long *is_state;
    unsigned int *is_ticket;
    int is_tiles;
    // deferred map copies (fs2_update_ws.cuh, DEFER): control words, leaders of the last resample, followers per leader
    int32_t *dctl, *leaders, *nfolv;
    int defer;                // the fused step may defer (one GPU, warp-specialised kernel, FS2_DEFER != 0)
    int defer_pending;        // a deferring resample chain was enqueued and no update has run since
    // host staging
    double *h_stats;          // pinned
    int *h_flags;             // pinned [2]
    int64_t launches;
    int red_blocks;
    int use_ws;               // warp-specialised update kernel (default) or the single-role one (FS2_KERNEL=v3)
    void *peers_dev;          // Fs2Peers: peer stores mapped with CUDA IPC (fs2_ipc_open_peers), or nullptr
    void *peer_bases[16][7];
    int peer_world;
    struct KlWork *kl;        // map-clustering workspace (fs2_known_landmarks), allocated on first use
};

static void kl_work_destroy(struct KlWork *w);
static int sync_maps(struct fs2_filter_s *h, cudaStream_t s);

extern "C" int fs2_abi_version(void) { return FS2_ABI_VERSION; }

extern "C" const char *fs2_strerror(int s)
{
    switch (s) {
        case FS2_OK: return "ok";
        case FS2_ERR_INVALID: return "invalid argument";
        case FS2_ERR_CUDA: return "CUDA runtime error";
        case FS2_ERR_NOMEM: return "out of device memory";
        case FS2_ERR_UNSUPPORTED: return "unsupported";
        default: return "unknown status";
    }
}

extern "C" const char *fs2_last_cuda_error(void) { return g_cuda_err; }

template <typename T>
static int dev_alloc(T **p, size_t n)
{
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T) + 16);
    if (e != cudaSuccess) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "cudaMalloc(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(e));
        (void)cudaGetLastError();
        return FS2_ERR_NOMEM;
    }
    *p = (T *)q;
    return FS2_OK;
}

#define FS2_ALLOC(ptr, n)                    \
    do {                                     \
        int r__ = dev_alloc(&(ptr), (n));    \
        if (r__ != FS2_OK) {                 \
            fs2_destroy(h);                  \
            return r__;                      \
        }                                    \
    } while (0)

__global__ void fs2_reset_kernel(Fs2State st, int64_t Pglobal)
{
    const double w0 = 1.0 / (double)Pglobal;  // particle.py:19
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < st.P; i += (int64_t)gridDim.x * blockDim.x) {
        st.x[i] = 0.0; st.y[i] = 0.0; st.yaw[i] = 0.0; st.w[i] = w0;
        st.count[i] = 0; st.slot[i] = (int32_t)i; st.status[i] = 0;
    }
}

__global__ void fs2_noise_kernel(double *noise, int64_t P, int64_t goff, const int32_t *ids, double sigma, uint64_t step, uint64_t seed)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
        // the draw belongs to the LOGICAL particle: ids[i] once a sharded filter places particles freely (fs2_place.cuh)
        uint64_t g = ids ? (uint64_t)ids[i] : (uint64_t)(goff + i);
        uint32_t r[4];
        fs2_philox4x32((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)step, (uint32_t)(step >> 32), (uint32_t)seed,
                       (uint32_t)(seed >> 32), r);
        noise[i] = sigma * fs2_normal_from_bits(r);   // loc + scale * gauss with loc = 0 (fast_slam_2.py:79/81)
    }
}

// canonical-order copy of selected particles' maps (slot indirection resolved) for host download
__global__ void fs2_pack_maps_kernel(const double *lm, const int32_t *slot, const int64_t *sel, int64_t nsel, int lcap, double *out)
{
    const int64_t per = 6 * (int64_t)lcap;
    for (int64_t r = blockIdx.x; r < nsel; r += gridDim.x) {
        int64_t p = sel ? sel[r] : r;
        const double *src = lm + (size_t)slot[p] * per;
        double *dst = out + (size_t)r * per;
        for (int64_t j = threadIdx.x; j < per; j += blockDim.x) dst[j] = src[j];
    }
}

static Fs2Scan make_scan(fs2_handle h)
{
    Fs2Scan sc;
    sc.A0 = h->A0; sc.A1 = h->A1; sc.eb = h->eb; sc.mode = h->mode; sc.cstart = h->cstart;
    sc.gA0 = h->gA0; sc.gA1 = h->gA1; sc.gE = h->gE; sc.gK = h->gK; sc.gV = h->gV; sc.cg = h->cg; sc.total = h->scan_total;
    return sc;
}

static Fs2State make_state(fs2_handle h)
{
    Fs2State st;
    st.x = h->x; st.y = h->y; st.yaw = h->yaw; st.w = h->w;
    st.count = h->count; st.slot = h->slot; st.lm = h->lm; st.status = h->status;
    st.P = h->P; st.lcap = h->lcap; st.pad = 0;
    return st;
}

extern "C" int fs2_destroy(fs2_handle h)
{
    if (!h) return FS2_OK;
    cudaSetDevice(h->cfg.device);
    void *ptrs[] = {h->x, h->y, h->yaw, h->w, h->count, h->slot, h->status, h->lm, h->noise, h->x2, h->y2, h->yaw2,
                    h->w2, h->count2, h->slot2, h->alive, h->extra, h->tasks, h->freeslot, h->ncopies, h->iscan_bs,
                    h->stats, h->partial, h->partial_sq, h->partial_best, h->counters, h->cumsum, h->bsum, h->bpre,
                    h->cstart, h->scan_total, h->A0, h->A1, h->eb, h->mode, h->anomaly, h->stuck, h->ancestor,
                    h->ctl, h->is_state, h->is_ticket, h->gA0, h->gA1, h->gE, h->gK, h->gV, h->cg,
                    h->logi, h->logi_new, h->src, h->ancl, h->pl_block, h->dctl, h->leaders, h->nfolv};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (int r = 0; r < h->peer_world; ++r)
        for (int i = 0; i < 7; ++i)
            if (h->peer_bases[r][i]) cudaIpcCloseMemHandle(h->peer_bases[r][i]);
    if (h->peers_dev) cudaFree(h->peers_dev);
    if (h->kl) kl_work_destroy(h->kl);
    if (h->h_stats) cudaFreeHost(h->h_stats);
    if (h->h_flags) cudaFreeHost(h->h_flags);
    free(h);
    return FS2_OK;
}

extern "C" int fs2_create(const fs2_config *cfg, fs2_handle *out)
{
    if (!cfg || !out) return FS2_ERR_INVALID;
    if (cfg->num_particles <= 0 || cfg->num_particles > 0x7fffffffLL || cfg->landmark_capacity <= 0) return FS2_ERR_INVALID;
    fs2_handle h = (fs2_handle)calloc(1, sizeof(fs2_filter_s));
    if (!h) return FS2_ERR_NOMEM;
    h->cfg = *cfg;
    h->P = cfg->num_particles;
    h->Pglobal = cfg->global_particles > 0 ? cfg->global_particles : cfg->num_particles;
    if (h->Pglobal > 0x7fffffffLL || cfg->global_offset < 0 || cfg->global_offset + h->P > h->Pglobal) {
        free(h);
        return FS2_ERR_INVALID;
    }
    h->lcap = cfg->landmark_capacity;
    if (cudaSetDevice(cfg->device) != cudaSuccess) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "cudaSetDevice(%d) failed", cfg->device);
        free(h);
        return FS2_ERR_CUDA;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) { free(h); return FS2_ERR_CUDA; }
    h->sm_count = prop.multiProcessorCount;
    const size_t P = (size_t)h->P, PG = (size_t)h->Pglobal;
    FS2_ALLOC(h->x, P); FS2_ALLOC(h->y, P); FS2_ALLOC(h->yaw, P); FS2_ALLOC(h->w, P);
    FS2_ALLOC(h->count, P); FS2_ALLOC(h->slot, P); FS2_ALLOC(h->status, P);
    h->S = h->P + (cfg->spare_slots > 0 ? cfg->spare_slots : 0);
    const size_t S = (size_t)h->S;
    FS2_ALLOC(h->lm, S * 6 * (size_t)h->lcap);
    FS2_ALLOC(h->noise, P);
    FS2_ALLOC(h->x2, P); FS2_ALLOC(h->y2, P); FS2_ALLOC(h->yaw2, P); FS2_ALLOC(h->w2, P);
    FS2_ALLOC(h->count2, P); FS2_ALLOC(h->slot2, P);
    FS2_ALLOC(h->alive, S); FS2_ALLOC(h->extra, P); FS2_ALLOC(h->tasks, P); FS2_ALLOC(h->freeslot, S);
    FS2_ALLOC(h->ncopies, 4);
    h->iscan_nb = (int)((S + FS2_ISCAN_B - 1) / FS2_ISCAN_B);
    FS2_ALLOC(h->iscan_bs, (size_t)h->iscan_nb);
    FS2_ALLOC(h->stats, FS2_STATS_LEN);
    h->red_blocks = h->sm_count * 8 < FS2_RED_MAX_BLOCKS ? h->sm_count * 8 : FS2_RED_MAX_BLOCKS;
    FS2_ALLOC(h->partial, FS2_RED_MAX_BLOCKS); FS2_ALLOC(h->partial_sq, FS2_RED_MAX_BLOCKS);
    FS2_ALLOC(h->partial_best, FS2_RED_MAX_BLOCKS); FS2_ALLOC(h->counters, 4);
    h->scan_nb = (int)((PG + FS2_SCAN_B - 1) / FS2_SCAN_B);
    FS2_ALLOC(h->cumsum, PG); FS2_ALLOC(h->bsum, (size_t)h->scan_nb); FS2_ALLOC(h->bpre, (size_t)h->scan_nb);
    FS2_ALLOC(h->cstart, (size_t)h->scan_nb); FS2_ALLOC(h->scan_total, 2);
    FS2_ALLOC(h->A0, (size_t)h->scan_nb); FS2_ALLOC(h->A1, (size_t)h->scan_nb);
    FS2_ALLOC(h->eb, (size_t)h->scan_nb); FS2_ALLOC(h->mode, (size_t)h->scan_nb);
    FS2_ALLOC(h->anomaly, 4); FS2_ALLOC(h->stuck, 4);
    h->scan_ng = (h->scan_nb + FS2_GRP - 1) / FS2_GRP;
    FS2_ALLOC(h->gA0, (size_t)h->scan_ng); FS2_ALLOC(h->gA1, (size_t)h->scan_ng);
    FS2_ALLOC(h->gE, (size_t)h->scan_ng); FS2_ALLOC(h->gK, (size_t)h->scan_ng); FS2_ALLOC(h->gV, (size_t)h->scan_ng);
    FS2_ALLOC(h->cg, (size_t)h->scan_ng);
    FS2_ALLOC(h->ancestor, P);
    FS2_ALLOC(h->ctl, FS2_CTL_LEN);
    h->is_tiles = (int)((S + FS2_IS_TILE - 1) / FS2_IS_TILE);
    FS2_ALLOC(h->is_state, (size_t)h->is_tiles);
    FS2_ALLOC(h->is_ticket, 4);
    FS2_ALLOC(h->dctl, 4); FS2_ALLOC(h->leaders, P); FS2_ALLOC(h->nfolv, P);
    if (cudaMallocHost((void **)&h->h_stats, FS2_STATS_LEN * sizeof(double)) != cudaSuccess ||
        cudaMallocHost((void **)&h->h_flags, 4 * sizeof(int)) != cudaSuccess) {
        fs2_destroy(h);
        return FS2_ERR_NOMEM;
    }
    cudaMemset(h->counters, 0, 4 * sizeof(unsigned int));
    cudaMemset(h->ctl, 0, FS2_CTL_LEN * sizeof(int));
    cudaMemset(h->is_ticket, 0, 4 * sizeof(unsigned int));
    cudaMemset(h->dctl, 0, 4 * sizeof(int32_t));
    cudaMemset(h->stats, 0, FS2_STATS_LEN * sizeof(double));
    // opt in to the update kernel's shared memory once
    const int smem = (int)sizeof(Fs2UpdateSmem);
    cudaFuncSetAttribute(fs2_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(fs2_update_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Fs2WsSmem));
    cudaFuncSetAttribute(fs2_update_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Fs2WsSmem));
    {
        const char *k = getenv("FS2_KERNEL");
        h->use_ws = !(k && strcmp(k, "v3") == 0);
        // experiments: shared-memory carve-out of the update kernel in per cent of the maximum (the rest is L1)
        const char *c = getenv("FS2_CARVEOUT");
        if (c) {
            cudaFuncSetAttribute(fs2_update_ws_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(c));
            cudaFuncSetAttribute(fs2_update_ws_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(c));
        }
        const char *d = getenv("FS2_DEFER");
        h->defer = h->use_ws && h->Pglobal == h->P && !(cfg->flags & FS2_FLAG_FORCE_SEQUENTIAL) && !(d && atoi(d) == 0);
    }
    int r = fs2_reset(h, nullptr);
    if (r != FS2_OK) { fs2_destroy(h); return r; }
    {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            snprintf(g_cuda_err, sizeof(g_cuda_err), "fs2_create: cudaDeviceSynchronize: %s", cudaGetErrorString(e));
            fs2_destroy(h);
            return FS2_ERR_CUDA;
        }
    }
    *out = h;
    return FS2_OK;
}

extern "C" int fs2_reset(fs2_handle h, void *stream)
{
    if (!h) return FS2_ERR_INVALID;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    { int r_ = sync_maps(h, (cudaStream_t)stream); if (r_ != FS2_OK) return r_; }
    cudaStream_t s = (cudaStream_t)stream;
    int blocks = (int)((h->P + 255) / 256);
    if (blocks > h->sm_count * 16) blocks = h->sm_count * 16;
    fs2_reset_kernel<<<blocks, 256, 0, s>>>(make_state(h), h->Pglobal);
    h->launches++;
    FS2_CUDA(cudaGetLastError());
    return FS2_OK;
}

extern "C" int fs2_get_ptrs(fs2_handle h, fs2_ptrs *o)
{
    if (!h || !o) return FS2_ERR_INVALID;
    o->x = h->x; o->y = h->y; o->yaw = h->yaw; o->w = h->w; o->count = h->count; o->lm = h->lm;
    o->status = h->status; o->noise = h->noise; o->cumsum = h->cumsum; o->ancestor = h->ancestor;
    o->stats = h->stats; o->num_particles = h->P; o->landmark_capacity = h->lcap; o->reserved = 0;
    return FS2_OK;
}

extern "C" int64_t fs2_launch_count(fs2_handle h) { return h ? h->launches : 0; }

extern "C" int fs2_draw_noise(fs2_handle h, double sigma, uint64_t step, double *noise_dev, void *stream)
{
    if (!h) return FS2_ERR_INVALID;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    double *dst = noise_dev ? noise_dev : h->noise;
    int blocks = (int)((h->P + 255) / 256);
    if (blocks > h->sm_count * 16) blocks = h->sm_count * 16;
    fs2_noise_kernel<<<blocks, 256, 0, s>>>(dst, h->P, h->cfg.global_offset, h->pl_on ? h->logi : nullptr, sigma, step, h->cfg.seed);
    h->launches++;
    FS2_CUDA(cudaGetLastError());
    return FS2_OK;
}

// distance from v to the interval of cell c (cells -1 and G are half-infinite)
static double cell_dist(double v, int c, int G, double x0, double s)
{
    double lo = (c < 0) ? -INFINITY : x0 + c * s;
    double hi = (c >= G) ? INFINITY : x0 + (c + 1) * s;
    if (v < lo) return lo - v;
    if (v > hi) return v - hi;
    return 0.0;
}

static void build_table(unsigned *tab, int G, const Fs2ObsBatch *ob, int m, double x0, double y0, double s, double margin)
{
    const int GP = G + 2;
    for (int k = 0; k < m; ++k) {
        if (!isfinite(ob->oxf[k]) || !isfinite(ob->oyf[k])) continue;
        // candidate cell range of this observation (one cell of slack each side), then the exact distance test
        int cx0 = (int)floor((ob->oxf[k] - margin - x0) / s) - 1, cx1 = (int)floor((ob->oxf[k] + margin - x0) / s) + 1;
        int cy0 = (int)floor((ob->oyf[k] - margin - y0) / s) - 1, cy1 = (int)floor((ob->oyf[k] + margin - y0) / s) + 1;
        if (cx0 < -1) cx0 = -1; if (cy0 < -1) cy0 = -1;
        if (cx1 > G) cx1 = G; if (cy1 > G) cy1 = G;
        for (int cy = cy0; cy <= cy1; ++cy)
            for (int cx = cx0; cx <= cx1; ++cx)
                if (cell_dist(ob->oxf[k], cx, G, x0, s) <= margin && cell_dist(ob->oyf[k], cy, G, y0, s) <= margin)
                    tab[(cy + 1) * GP + cx + 1] |= 1u << k;
    }
}

// robot-frame Cartesian of every observation with the HOST libm (the same cos/sin the oracle calls), and
// the cell tables the screen uses: tabN[cell] = observations within eN (Chebyshev, plus rounding margin)
// of the cell, so a box of half-width <= eN centred anywhere in the cell can only contain those.
static void fill_batch(Fs2ObsBatch *ob, const double *obs, int k0, int m, double gate)
{
    memset(ob, 0, sizeof(*ob));
    float omax = 0.f;
    const float inf = INFINITY;
    float x0 = inf, x1 = -inf, y0 = inf, y1 = -inf;
    for (int k = 0; k < 32; ++k) {
        if (k < m) {
            double zd = obs[2 * (k0 + k)], za = obs[2 * (k0 + k) + 1];
            ob->zd[k] = zd; ob->za[k] = za;
            ob->cza[k] = cos(za); ob->sza[k] = sin(za);
            ob->ox[k] = zd * ob->cza[k];        // fast_slam_2.py:101
            ob->oy[k] = zd * ob->sza[k];        // fast_slam_2.py:102
            ob->oxf[k] = (float)ob->ox[k];
            ob->oyf[k] = (float)ob->oy[k];
            float ax = fabsf(ob->oxf[k]), ay = fabsf(ob->oyf[k]);
            if (isfinite(ax) && isfinite(ay)) {
                if (ax > omax) omax = ax;
                if (ay > omax) omax = ay;
                if (ob->oxf[k] < x0) x0 = ob->oxf[k];
                if (ob->oxf[k] > x1) x1 = ob->oxf[k];
                if (ob->oyf[k] < y0) y0 = ob->oyf[k];
                if (ob->oyf[k] > y1) y1 = ob->oyf[k];
            }
        } else {
            ob->oxf[k] = inf; ob->oyf[k] = inf;  // never inside a box
            ob->ox[k] = INFINITY; ob->oy[k] = INFINITY;
        }
    }
    ob->slack = 2.4e-7f * omax + 1e-30f;
    ob->M = m;
    ob->k0 = k0;
    ob->all_mask = (m >= 32) ? 0xffffffffu : ((1u << m) - 1u);
    ob->amax1 = ob->amax2 = -1.f; ob->xymax = 0.f;
    if (!(x0 <= x1)) {                           // no finite observation: nothing can match
        ob->gx0 = ob->gy0 = 0.f; ob->inv_s1 = ob->inv_s2 = 0.f; ob->e1 = ob->e2 = -1.f;
        return;
    }
    double ext = fmax((double)x1 - x0, (double)y1 - y0);
    if (ext < 1e-3) ext = 1e-3;
    const double coord = fmax(fmax(fabs(x0), fabs(x1)), fmax(fabs(y0), fabs(y1))) + ext;
    ob->gx0 = x0; ob->gy0 = y0;
    // margins are independent of the cell size.  e2 covers a landmark that still has the reference's default
    // covariance 0.1*I (landmark.py:13): gate * sqrt(0.1) plus the box widening; e1, a quarter of it, covers
    // landmarks that have been updated at least once.  Boxes wider than e2 are screened against all observations.
    ob->e2 = (float)(gate * sqrt(0.1) * 1.03 + 1e-3);
    ob->e1 = 0.25f * ob->e2;
    // Each level's grid covers the observations' bounding square PLUS its own margin on every side, so that the
    // half-infinite border cells lie beyond every observation's reach and stay empty: a landmark outside the covered
    // square is never a candidate at that level.
    const double pad1 = (double)ob->e1 * 1.02 + 1e-3 * ext + 1e-5 * coord, pad2 = (double)ob->e2 * 1.02 + 1e-3 * ext + 1e-5 * coord;
    const double s1 = (ext + 2.0 * pad1) / FS2_G1, s2 = (ext + 2.0 * pad2) / FS2_G2;
    ob->inv_s1 = (float)(1.0 / s1); ob->inv_s2 = (float)(1.0 / s2);
    const double gx1 = (double)x0 - pad1, gy1 = (double)y0 - pad1, gx2 = (double)x0 - pad2, gy2 = (double)y0 - pad2;
    // the device finds the cell in fp32 from inv_s as rounded above: margins cover that and the box centre's own rounding
    const double m1 = (double)ob->e1 * 1.0001 + 2e-4 * s1 + 4e-6 * coord;
    const double m2 = (double)ob->e2 * 1.0001 + 2e-4 * s2 + 4e-6 * coord;
    build_table(ob->tab1, FS2_G1, ob, m, gx1, gy1, 1.0 / (double)ob->inv_s1, m1);
    build_table(ob->tab2, FS2_G2, ob, m, gx2, gy2, 1.0 / (double)ob->inv_s2, m2);
    // table cell of a point: floor(clamp(fma(x, inv_s, c), 0, G + 1)) with c = 1 - origin * inv_s (fs2_cell)
    ob->cx1 = (float)(1.0 - gx1 * (double)ob->inv_s1); ob->cy1 = (float)(1.0 - gy1 * (double)ob->inv_s1);
    ob->cx2 = (float)(1.0 - gx2 * (double)ob->inv_s2); ob->cy2 = (float)(1.0 - gy2 * (double)ob->inv_s2);
    // level of a landmark without building its box: fs2_box gives a safe landmark the half-width
    //   rx = (gate_f * 1.00001f) * sqrt.approx(c00) + 2.4e-7 * |x| + slack   (likewise ry),
    // so max(c00, c11) <= amaxN and |x|, |y| < xymax imply rx, ry <= eN.  xymax lies beyond every observation by
    // more than e2: a safe landmark of level <= 2 out there cannot gate any observation (|dx| < gate * sqrt(c00)).
    ob->xymax = (float)(((double)omax + (double)ob->e2) * 1.001 + 1e-3);
    const double gf = (double)((float)gate * 1.0000002f) * 1.00001 * (1.0 + 1e-6);   // >= gate_f * 1.00001f as the device rounds it
    const double lim[2] = {(double)ob->e1, (double)ob->e2};
    float *amax[2] = {&ob->amax1, &ob->amax2};
    for (int l = 0; l < 2; ++l) {
        const double room = lim[l] - 2.4e-7 * (double)ob->xymax * 1.0001 - (double)ob->slack * 1.0001;
        if (room > 0.0) {
            const double q = room / (gf * (1.0 + 4e-7));        // sqrt.approx: 2^-22 relative
            *amax[l] = (float)(q * q * (1.0 - 2e-6));
            if ((double)*amax[l] > q * q * (1.0 - 1e-6)) *amax[l] = nextafterf(*amax[l], 0.f);
        }
    }
}

// host-only: the per-step observation block exactly as the update kernel receives it (tests check that the
// cell tables are conservative).  out must hold fs2_debug_obs_batch_size() bytes.
extern "C" int fs2_debug_obs_batch_size(void) { return (int)sizeof(Fs2ObsBatch); }
extern "C" int fs2_debug_obs_batch(const double *obs_host, int32_t M, void *out)
{
    if (!obs_host || !out || M < 0 || M > 32) return FS2_ERR_INVALID;
    fill_batch((Fs2ObsBatch *)out, obs_host, 0, M, 8.0);
    return FS2_OK;
}

// map copies a deferring resample left to the next update kernel: make them now (fs2_materialize_kernel) -- for every
// caller that is about to touch the maps and is not that update kernel
static int sync_maps(fs2_handle h, cudaStream_t s)
{
    if (!h->defer_pending) return FS2_OK;
    fs2_materialize_kernel<<<h->sm_count * 8, 256, 0, s>>>(h->dctl, h->leaders, h->nfolv, h->slot, h->count, h->lm, h->lcap);
    h->launches++;
    FS2_CUDA(cudaGetLastError());
    FS2_CUDA(cudaMemsetAsync(h->dctl, 0, 4 * sizeof(int32_t), s));
    h->defer_pending = 0;
    return FS2_OK;
}

extern "C" int fs2_sync_maps(fs2_handle h, void *stream)
{
    if (!h) return FS2_ERR_INVALID;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    return sync_maps(h, (cudaStream_t)stream);
}

static int launch_update(fs2_handle h, int do_motion, double rotation, double translation, const double *noise_dev,
                         const double *obs_host, int32_t M, int32_t *assoc_dev, cudaStream_t s)
{
    if (M < 0 || (M > 0 && !obs_host)) return FS2_ERR_INVALID;
    if (do_motion && !noise_dev) return FS2_ERR_INVALID;
    // a pending deferred copy rides in the first launch of this update if that launch streams the maps
    bool ride = h->defer_pending && h->use_ws && M > 0 && !(h->cfg.flags & FS2_FLAG_FORCE_SEQUENTIAL);
    if (h->defer_pending && !ride) {
        int r = sync_maps(h, s);
        if (r != FS2_OK) return r;
    }
    Fs2UpdateArgs ua;
    memset(&ua, 0, sizeof(ua));
    ua.r00 = h->cfg.measurement_noise[0]; ua.r01 = h->cfg.measurement_noise[1];
    ua.r10 = h->cfg.measurement_noise[2]; ua.r11 = h->cfg.measurement_noise[3];
    ua.gate = h->cfg.max_landmark_distance;
    ua.gate_f = (float)fabs(h->cfg.max_landmark_distance) * 1.0000002f;  // never below the fp64 gate
    ua.rotation = rotation; ua.translation = translation; ua.noise = noise_dev;
    ua.assoc = assoc_dev;
    ua.force_seq = (h->cfg.flags & FS2_FLAG_FORCE_SEQUENTIAL) ? 1 : 0;
    const int smem = (int)sizeof(Fs2UpdateSmem);
    int64_t blocks64 = (h->P + FS2_WPB - 1) / FS2_WPB;
    int blocks = (int)(blocks64 < (int64_t)h->sm_count * 2 ? blocks64 : (int64_t)h->sm_count * 2);
    const Fs2State st = make_state(h);
    int k0 = 0;
    bool first = true;
    do {   // batches of <= 32 observations; a step without observations still moves the particles
        int m = M - k0 < 32 ? M - k0 : 32;
        Fs2ObsBatch ob;
        fill_batch(&ob, obs_host, k0, m, fabs(h->cfg.max_landmark_distance));
        ua.do_motion = (do_motion && first) ? 1 : 0;
        if (h->use_ws) {
            int64_t wb64 = (h->P + FS2_SW - 1) / FS2_SW;
            int wblocks = (int)(wb64 < (int64_t)h->sm_count * FS2_WS_MINB ? wb64 : (int64_t)h->sm_count * FS2_WS_MINB);
            if (ride) {      // the device knows whether the last step resampled: the form that does not apply returns at once
                ua.dctl = h->dctl; ua.leaders = h->leaders; ua.nfol = h->nfolv;
                fs2_update_ws_kernel<true><<<wblocks, FS2_WS_THREADS, (int)sizeof(Fs2WsSmem), s>>>(st, ob, ua);
                h->launches++;
            }
            fs2_update_ws_kernel<false><<<wblocks, FS2_WS_THREADS, (int)sizeof(Fs2WsSmem), s>>>(st, ob, ua);
            if (ride) {
                FS2_CUDA(cudaMemsetAsync(h->dctl, 0, 4 * sizeof(int32_t), s));
                ua.dctl = nullptr;
                h->defer_pending = 0;
                ride = false;
            }
        } else {
            fs2_update_kernel<<<blocks, FS2_WPB * 32, smem, s>>>(st, ob, ua);
        }
        h->launches++;
        FS2_CUDA(cudaGetLastError());
        first = false;
        k0 += m;
    } while (k0 < M);
    return FS2_OK;
}

extern "C" int fs2_motion(fs2_handle h, double rotation, double translation, const double *noise_dev, void *stream)
{
    if (!h) return FS2_ERR_INVALID;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    return launch_update(h, 1, rotation, translation, noise_dev ? noise_dev : h->noise, nullptr, 0, nullptr, (cudaStream_t)stream);
}

extern "C" int fs2_update(fs2_handle h, const double *obs_host, int32_t M, int32_t *assoc_dev, void *stream)
{
    if (!h) return FS2_ERR_INVALID;
    if (M == 0) return FS2_OK;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    return launch_update(h, 0, 0.0, 0.0, nullptr, obs_host, M, assoc_dev, (cudaStream_t)stream);
}

extern "C" int fs2_motion_update(fs2_handle h, double rotation, double translation, const double *noise_dev,
                                 const double *obs_host, int32_t M, int32_t *assoc_dev, void *stream)
{
    if (!h) return FS2_ERR_INVALID;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    return launch_update(h, 1, rotation, translation, noise_dev ? noise_dev : h->noise, obs_host, M, assoc_dev, (cudaStream_t)stream);
}

extern "C" int fs2_weight_total(fs2_handle h, void *stream)
{
    if (!h) return FS2_ERR_INVALID;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    int blocks = (int)((h->P + FS2_RED_THREADS - 1) / FS2_RED_THREADS);
    if (blocks > h->red_blocks) blocks = h->red_blocks;
    fs2_weight_total_kernel<<<blocks, FS2_RED_THREADS, 0, (cudaStream_t)stream>>>(h->w, h->P, h->partial, h->counters, h->stats);
    h->launches++;
    FS2_CUDA(cudaGetLastError());
    return FS2_OK;
}

static int launch_normalize(fs2_handle h, const double *total_dev, int apply, cudaStream_t s)
{
    int blocks = (int)((h->P + FS2_RED_THREADS - 1) / FS2_RED_THREADS);
    if (blocks > h->red_blocks) blocks = h->red_blocks;
    fs2_normalize_kernel<<<blocks, FS2_RED_THREADS, 0, s>>>(h->w, h->x, h->y, h->yaw, h->P, h->Pglobal,
                                                           total_dev ? total_dev : h->stats, apply, h->partial_sq,
                                                           h->partial_best, h->counters + 1, h->stats,
                                                           h->pl_on ? h->logi : nullptr, h->cfg.global_offset);
    h->launches++;
    FS2_CUDA(cudaGetLastError());
    return FS2_OK;
}

extern "C" int fs2_normalize(fs2_handle h, const double *total_dev, void *stream)
{
    if (!h) return FS2_ERR_INVALID;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    return launch_normalize(h, total_dev, 1, (cudaStream_t)stream);
}

extern "C" int fs2_estimate(fs2_handle h, void *stream)
{
    if (!h) return FS2_ERR_INVALID;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    return launch_normalize(h, h->stats, 0, (cudaStream_t)stream);
}

extern "C" int fs2_resample_indices(fs2_handle h, const double *w_all_dev, int64_t n, double u0, int64_t m_begin,
                                    int64_t m_count, int32_t *ancestor_dev, void *stream)
{
    if (!h || !w_all_dev || n <= 0 || n > h->Pglobal || m_begin < 0 || m_count < 0 || m_begin + m_count > n) return FS2_ERR_INVALID;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    int32_t *anc = ancestor_dev ? ancestor_dev : h->ancestor;
    if (!ancestor_dev && m_count > h->P) return FS2_ERR_INVALID;
    const int nb = (int)((n + FS2_SCAN_B - 1) / FS2_SCAN_B);
    FS2_CUDA(cudaMemsetAsync(h->anomaly, 0, sizeof(int), s));
    FS2_CUDA(cudaMemsetAsync(h->stuck, 0, sizeof(int), s));
    fs2_scan_blocksum<<<nb, FS2_SCAN_T, 0, s>>>(w_all_dev, n, h->bsum, h->anomaly);
    FS2_CUDA(cudaGetLastError());
    // the anomaly flag decides the path; it is one int and the scan is short: read it back
    FS2_CUDA(cudaMemcpyAsync(h->h_flags, h->anomaly, sizeof(int), cudaMemcpyDeviceToHost, s));
    FS2_CUDA(cudaStreamSynchronize(s));
    h->launches += 1;
    if (h->h_flags[0]) {
        fs2_resample_serial<<<1, 32, 0, s>>>(w_all_dev, n, u0, m_begin, m_count, anc, h->cumsum);
        h->launches += 1;
        FS2_CUDA(cudaGetLastError());
        return FS2_OK;
    }
    fs2_scan_blockprefix<<<1, 1024, 0, s>>>(h->bsum, nb, h->bpre);
    {
        const int ng = (nb + FS2_GRP - 1) / FS2_GRP;
        Fs2ScanFused none;
        memset(&none, 0, sizeof(none));
        fs2_scan_groupfunc<<<ng, FS2_GRP_T, 0, s>>>(w_all_dev, n, nb, h->bsum, h->bpre, make_scan(h), h->counters + 2, none);
        fs2_scan_emit<<<ng, FS2_GRP_T, 0, s>>>(w_all_dev, n, nb, make_scan(h), h->cumsum, nullptr);
    }
    int sblocks = (int)((m_count + 255) / 256);
    if (sblocks > h->sm_count * 16) sblocks = h->sm_count * 16;
    if (sblocks > 0) fs2_resample_search<<<sblocks, 256, 0, s>>>(h->cumsum, n, u0, m_begin, m_count, anc, h->stuck);
    h->launches += 4;
    FS2_CUDA(cudaGetLastError());
    return FS2_OK;
}

static int commit_gather(fs2_handle h, cudaStream_t s)
{
    const size_t bd = sizeof(double) * (size_t)h->P, bi = sizeof(int32_t) * (size_t)h->P;
    FS2_CUDA(cudaMemcpyAsync(h->x, h->x2, bd, cudaMemcpyDeviceToDevice, s));
    FS2_CUDA(cudaMemcpyAsync(h->y, h->y2, bd, cudaMemcpyDeviceToDevice, s));
    FS2_CUDA(cudaMemcpyAsync(h->yaw, h->yaw2, bd, cudaMemcpyDeviceToDevice, s));
    FS2_CUDA(cudaMemcpyAsync(h->w, h->w2, bd, cudaMemcpyDeviceToDevice, s));
    FS2_CUDA(cudaMemcpyAsync(h->count, h->count2, bi, cudaMemcpyDeviceToDevice, s));
    FS2_CUDA(cudaMemcpyAsync(h->slot, h->slot2, bi, cudaMemcpyDeviceToDevice, s));
    return FS2_OK;
}

// anc: ancestors of this shard's P slots (encoding: see fs2_resample.cuh).  anc_all != nullptr selects peer mode
// (anc = anc_all + rank * P, global indices).  Returns FS2_ERR_NOMEM without touching the store if the pool has
// fewer free slots than copies are needed (peer mode only; the caller then falls back to the staged path).
static int launch_gather(fs2_handle h, const int32_t *anc, const double *rec, const int32_t *anc_all, int commit, cudaStream_t s)
{
    const int64_t P = h->P, S = h->S;
    const int64_t rstride = 8 + 6 * (int64_t)h->lcap;
    const Fs2Peers *peers = anc_all ? (const Fs2Peers *)h->peers_dev : nullptr;
    int blocks = (int)((P + 255) / 256);
    if (blocks > h->sm_count * 16) blocks = h->sm_count * 16;
    { int r = sync_maps(h, s); if (r != FS2_OK) return r; }
    FS2_CUDA(cudaMemsetAsync(h->alive, 0, sizeof(int32_t) * (size_t)S, s));
    fs2_gather_mark<<<blocks, 256, 0, s>>>(anc, P, peers, h->slot, h->alive, h->extra);
    if (anc_all) {
        int gb = (int)((h->Pglobal + 255) / 256);
        if (gb > h->sm_count * 16) gb = h->sm_count * 16;
        fs2_gather_mark_global<<<gb, 256, 0, s>>>(anc_all, h->Pglobal, P, (int)(h->cfg.global_offset / P), h->slot, h->alive);
        h->launches++;
    }
    fs2_iscan_sums<<<h->iscan_nb, 256, 0, s>>>(h->extra, h->alive, P, S, h->iscan_bs);
    fs2_iscan_prefix<<<1, 1024, 0, s>>>(h->iscan_bs, h->iscan_nb, h->ncopies);
    h->launches += 3;
    if (anc_all) {   // enough free slots?  (always true without peers: free = P - survivors = copies)
        FS2_CUDA(cudaMemcpyAsync(h->h_flags, h->ncopies, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
        FS2_CUDA(cudaStreamSynchronize(s));
        if (h->h_flags[0] > h->h_flags[1]) return FS2_ERR_NOMEM;
    }
    fs2_iscan_apply<<<h->iscan_nb, 256, 0, s>>>(h->extra, h->alive, P, S, h->iscan_bs, h->tasks, h->freeslot);
    fs2_gather_pose<<<blocks, 256, 0, s>>>(anc, h->extra, P, h->x, h->y, h->yaw, h->w, h->count, h->slot, rec, rstride, peers,
                                          h->x2, h->y2, h->yaw2, h->w2, h->count2, h->slot2);
    int cblocks = h->sm_count * 8;
    int64_t need = (P + 7) / 8;
    if ((int64_t)cblocks > need) cblocks = (int)need;
    fs2_gather_copy<<<cblocks, 256, 0, s>>>(h->tasks, h->freeslot, h->ncopies, anc, P, h->slot, h->count, rec, rstride, peers,
                                           h->lm, h->lcap, h->slot2);
    h->launches += 3;
    FS2_CUDA(cudaGetLastError());
    return commit ? commit_gather(h, s) : FS2_OK;
}

extern "C" int fs2_gather(fs2_handle h, const int32_t *ancestor_dev, void *stream)
{
    if (!h) return FS2_ERR_INVALID;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    return launch_gather(h, ancestor_dev ? ancestor_dev : h->ancestor, nullptr, nullptr, 1, (cudaStream_t)stream);
}

extern "C" int fs2_gather_ext(fs2_handle h, const int32_t *ancestor_dev, const double *records_dev, int64_t n_staged,
                              void *stream)
{
    if (!h || !ancestor_dev || n_staged < 0 || (n_staged > 0 && !records_dev)) return FS2_ERR_INVALID;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    return launch_gather(h, ancestor_dev, records_dev, nullptr, 1, (cudaStream_t)stream);
}

// peer mode: ancestors_all_dev = GLOBAL ancestor of every global slot (int32[global_particles]).  New poses land in
// the second buffer; fs2_gather_commit publishes them once every rank has finished reading (barrier in between).
extern "C" int fs2_gather_p2p(fs2_handle h, const int32_t *ancestors_all_dev, void *stream)
{
    if (!h || !ancestors_all_dev || !h->peers_dev) return FS2_ERR_INVALID;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    return launch_gather(h, ancestors_all_dev + h->cfg.global_offset, nullptr, ancestors_all_dev, 0, (cudaStream_t)stream);
}

extern "C" int fs2_gather_commit(fs2_handle h, void *stream)
{
    if (!h) return FS2_ERR_INVALID;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    return commit_gather(h, (cudaStream_t)stream);
}

// ---- peer-memory migration (one node, NVLink): the stores of all shards are mapped into every process ----
extern "C" int fs2_ipc_export(fs2_handle h, void *handles_out /* 7 x cudaIpcMemHandle_t = 448 bytes */)
{
    if (!h || !handles_out) return FS2_ERR_INVALID;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    cudaIpcMemHandle_t *o = (cudaIpcMemHandle_t *)handles_out;
    void *ptrs[7] = {h->x, h->y, h->yaw, h->w, h->lm, h->count, h->slot};
    for (int i = 0; i < 7; ++i) FS2_CUDA(cudaIpcGetMemHandle(&o[i], ptrs[i]));
    return FS2_OK;
}

extern "C" int fs2_ipc_open_peers(fs2_handle h, const void *all_handles /* world x 448 bytes, rank order */, int32_t world, int32_t rank)
{
    if (!h || !all_handles || world < 1 || world > 16 || rank < 0 || rank >= world) return FS2_ERR_INVALID;
    if (h->Pglobal != h->P * world) return FS2_ERR_INVALID;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    const cudaIpcMemHandle_t *a = (const cudaIpcMemHandle_t *)all_handles;
    Fs2Peers hp;
    memset(&hp, 0, sizeof(hp));
    hp.rank = rank; hp.world = world;
    // a second call re-maps: close what the first one opened
    for (int r = 0; r < h->peer_world; ++r)
        for (int i = 0; i < 7; ++i)
            if (h->peer_bases[r][i]) { cudaIpcCloseMemHandle(h->peer_bases[r][i]); h->peer_bases[r][i] = nullptr; }
    h->peer_world = world;      // set first: whatever is opened below is recorded as it is opened, fs2_destroy closes it
    void *own[7] = {h->x, h->y, h->yaw, h->w, h->lm, h->count, h->slot};
    for (int r = 0; r < world; ++r) {
        void *p[7];
        for (int i = 0; i < 7; ++i) {
            if (r == rank) p[i] = own[i];
            else {
                FS2_CUDA(cudaIpcOpenMemHandle(&p[i], a[r * 7 + i], cudaIpcMemLazyEnablePeerAccess));
                h->peer_bases[r][i] = p[i];
            }
        }
        hp.x[r] = (const double *)p[0]; hp.y[r] = (const double *)p[1]; hp.yaw[r] = (const double *)p[2];
        hp.w[r] = (const double *)p[3]; hp.lm[r] = (const double *)p[4];
        hp.count[r] = (const int32_t *)p[5]; hp.slot[r] = (const int32_t *)p[6];
    }
    if (!h->peers_dev) FS2_CUDA(cudaMalloc((void **)&h->peers_dev, sizeof(Fs2Peers)));
    FS2_CUDA(cudaMemcpy(h->peers_dev, &hp, sizeof(hp), cudaMemcpyHostToDevice));
    return FS2_OK;
}

// one block per record: read particle ids[r] (LOCAL index on the source) of shard `src` straight out of that GPU's
// store (peer loads over NVLink) and lay it down as a record in this GPU's staging buffer
__global__ void fs2_pull_records_kernel(const Fs2Peers *peers, int src, const int64_t *ids, int64_t off, int64_t n,
                                        int lcap, double *rec)
{
    const int64_t per = 6 * (int64_t)lcap, stride = per + 8;
    const double *px = peers->x[src], *py = peers->y[src], *pyaw = peers->yaw[src], *pw = peers->w[src], *plm = peers->lm[src];
    const int32_t *pc = peers->count[src], *ps = peers->slot[src];
    for (int64_t r = blockIdx.x; r < n; r += gridDim.x) {
        const int64_t i = ids[r] - off;
        double *dst = rec + (size_t)r * stride;
        const int cnt = pc[i];
        if (threadIdx.x == 0) { dst[0] = px[i]; dst[1] = py[i]; dst[2] = pyaw[i]; dst[3] = pw[i]; dst[4] = (double)cnt; }
        const int4 *srcm = reinterpret_cast<const int4 *>(plm + (size_t)ps[i] * per);
        int4 *dstm = reinterpret_cast<int4 *>(dst + 8);
        for (int g = threadIdx.x; g < cnt * 3; g += blockDim.x) dstm[g] = srcm[g];
    }
}

extern "C" int fs2_pull_records(fs2_handle h, int32_t src_rank, const int64_t *global_ids_dev, int64_t n, double *records_dev,
                                void *stream)
{
    if (!h || !h->peers_dev || src_rank < 0 || src_rank >= h->peer_world || n < 0 || (n > 0 && (!global_ids_dev || !records_dev)))
        return FS2_ERR_INVALID;
    if (n == 0) return FS2_OK;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    { int r_ = sync_maps(h, (cudaStream_t)stream); if (r_ != FS2_OK) return r_; }
    int blocks = (int)(n < (int64_t)h->sm_count * 16 ? n : (int64_t)h->sm_count * 16);
    fs2_pull_records_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const Fs2Peers *)h->peers_dev, src_rank, global_ids_dev,
                                                                      (int64_t)src_rank * h->P, n, h->lcap, records_dev);
    h->launches++;
    FS2_CUDA(cudaGetLastError());
    return FS2_OK;
}

// pack selected LOCAL particles (pose, weight, count, map rows) for sending to another GPU
__global__ void fs2_pack_records_kernel(const double *x, const double *y, const double *yaw, const double *w,
                                        const int32_t *count, const int32_t *slot, const double *lm, int lcap,
                                        const int64_t *sel, int64_t nsel, double *rec)
{
    const int64_t per = 6 * (int64_t)lcap, stride = per + 8;
    for (int64_t r = blockIdx.x; r < nsel; r += gridDim.x) {
        const int64_t p = sel[r];
        double *dst = rec + (size_t)r * stride;
        const int n = count[p];
        if (threadIdx.x == 0) { dst[0] = x[p]; dst[1] = y[p]; dst[2] = yaw[p]; dst[3] = w[p]; dst[4] = (double)n; }
        const double *src = lm + (size_t)slot[p] * per;
        for (int64_t j = threadIdx.x; j < 6 * (int64_t)n; j += blockDim.x) dst[8 + j] = src[j];
    }
}

extern "C" int fs2_pack_records(fs2_handle h, const int64_t *sel_dev, int64_t nsel, double *records_dev, void *stream)
{
    if (!h || nsel < 0 || (nsel > 0 && (!sel_dev || !records_dev))) return FS2_ERR_INVALID;
    if (nsel == 0) return FS2_OK;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    { int r_ = sync_maps(h, (cudaStream_t)stream); if (r_ != FS2_OK) return r_; }
    int blocks = (int)(nsel < (int64_t)h->sm_count * 16 ? nsel : (int64_t)h->sm_count * 16);
    fs2_pack_records_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(h->x, h->y, h->yaw, h->w, h->count, h->slot, h->lm,
                                                                      h->lcap, sel_dev, nsel, records_dev);
    h->launches++;
    FS2_CUDA(cudaGetLastError());
    return FS2_OK;
}

// ---- decoupled placement of a sharded filter (fs2_place.cuh) ------------------------------------------------------
__global__ void fs2_iota_kernel(int32_t *a, int64_t n, int64_t off)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a[i] = (int32_t)(off + i);
}

extern "C" int fs2_place_enable(fs2_handle h, void *stream)
{
    if (!h) return FS2_ERR_INVALID;
    if (h->pl_on) return FS2_OK;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    const size_t P = (size_t)h->P, N = (size_t)h->Pglobal;
    if (h->Pglobal % h->P != 0 || h->Pglobal / h->P > PL_MAXR) return FS2_ERR_UNSUPPORTED;
    int r;
    if ((r = dev_alloc(&h->logi, P)) != FS2_OK || (r = dev_alloc(&h->logi_new, P)) != FS2_OK ||
        (r = dev_alloc(&h->src, P)) != FS2_OK || (r = dev_alloc(&h->ancl, P)) != FS2_OK)
        return r;
    PlPlan &pl = h->pl;
    memset(&pl, 0, sizeof(pl));
    pl.N = h->Pglobal; pl.P = h->P;
    pl.world = (int)(h->Pglobal / h->P); pl.rank = (int)(h->cfg.global_offset / h->P);
    const size_t ntiles = (N + PL_TILE - 1) / PL_TILE;
    const size_t a16 = 256;
    auto up = [&](size_t b) { return (b + a16 - 1) / a16 * a16; };
    const size_t bytes = 3 * up(N) + 2 * up(4 * N) + up(4 * ntiles * PL_MAXR) + up(4 * PL_MAXR) + up(4 * PL_MAXR * PL_HB) +
                         4 * up(4 * PL_MAXR) + up(8 * 4);
    unsigned char *b = nullptr;
    if ((r = dev_alloc(&b, bytes)) != FS2_OK) return r;
    h->pl_block = b;
    size_t o = 0;
    auto take = [&](size_t n) { unsigned char *q = b + o; o += up(n); return q; };
    pl.home = take(N); pl.dest = take(N); pl.cls = take(N);
    pl.ocnt = (int32_t *)take(4 * N); pl.prefix = (int32_t *)take(4 * N);
    pl.tilecnt = (unsigned *)take(4 * ntiles * PL_MAXR);
    pl.cnt = (unsigned *)take(4 * PL_MAXR);
    pl.hist = (unsigned *)take(4 * PL_MAXR * PL_HB);
    pl.surplus = (int *)take(4 * PL_MAXR); pl.thresh = (int *)take(4 * PL_MAXR);
    pl.expoff = (unsigned *)take(4 * PL_MAXR); pl.impend = (unsigned *)take(4 * PL_MAXR);
    pl.info = (unsigned long long *)take(8 * 4);
    cudaStream_t s = (cudaStream_t)stream;
    int blocks = (int)((h->P + 255) / 256);
    if (blocks > h->sm_count * 16) blocks = h->sm_count * 16;
    fs2_iota_kernel<<<blocks, 256, 0, s>>>(h->logi, h->P, h->cfg.global_offset);
    FS2_CUDA(cudaGetLastError());
    h->pl_on = 1;
    {   // the copies of a placed resample ride in the next update launch too (fs2_place.cuh, pl_mark_kernel)
        const char *d = getenv("FS2_DEFER");
        h->defer = h->use_ws && !(h->cfg.flags & FS2_FLAG_FORCE_SEQUENTIAL) && !(d && atoi(d) == 0);
    }
    return FS2_OK;
}

extern "C" void *fs2_place_logical_ids(fs2_handle h) { return (h && h->pl_on) ? (void *)h->logi : nullptr; }

// anc_all_dev: LOGICAL ancestor of every new logical particle (int32[N], non-decreasing); place_dev: physical position
// (rank * P + local index) of every current logical particle; place_new_dev: receives the new table.  Runs the plan, then
// this rank's gather up to (not including) the publication of the new poses: other ranks are reading this store
// meanwhile.  info_host (optional, 2 words): offspring that changed GPU (all ranks), maps this rank pulled over NVLink.
extern "C" int fs2_place_resample(fs2_handle h, const int32_t *anc_all_dev, const int32_t *place_dev, int32_t *place_new_dev,
                                  int64_t *info_host, void *stream)
{
    if (!h || !h->pl_on || !h->peers_dev || !anc_all_dev || !place_dev || !place_new_dev) return FS2_ERR_INVALID;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    { int r_ = sync_maps(h, (cudaStream_t)stream); if (r_ != FS2_OK) return r_; }
    cudaStream_t s = (cudaStream_t)stream;
    PlPlan &pl = h->pl;
    const int64_t N = pl.N, P = pl.P, S = h->S;
    const int ntiles = (int)((N + PL_TILE - 1) / PL_TILE);
    int gb = (int)((N + 255) / 256);
    if (gb > h->sm_count * 16) gb = h->sm_count * 16;
    FS2_CUDA(cudaMemsetAsync(pl.ocnt, 0, sizeof(int32_t) * (size_t)N, s));
    FS2_CUDA(cudaMemsetAsync(pl.cnt, 0, sizeof(unsigned) * PL_MAXR, s));
    FS2_CUDA(cudaMemsetAsync(pl.hist, 0, sizeof(unsigned) * PL_MAXR * PL_HB, s));
    FS2_CUDA(cudaMemsetAsync(pl.info, 0, sizeof(unsigned long long) * 4, s));
    FS2_CUDA(cudaMemsetAsync(h->alive, 0, sizeof(int32_t) * (size_t)S, s));
    pl_home_kernel<<<gb, 256, 0, s>>>(pl, anc_all_dev, place_dev);
    pl_hist_kernel<<<gb, 256, 0, s>>>(pl, anc_all_dev);
    pl_decide_kernel<<<1, 32, 0, s>>>(pl);
    pl_eligible_kernel<<<gb, 256, 0, s>>>(pl, anc_all_dev);
    pl_mcscan_count<<<ntiles, PL_T, 0, s>>>(pl.cls, N, pl.tilecnt);
    pl_mcscan_prefix<<<1, 1024, 0, s>>>(pl.tilecnt, ntiles);
    pl_mcscan_apply<<<ntiles, PL_T, 0, s>>>(pl.cls, N, pl.tilecnt, pl.prefix);
    pl_dest_kernel<<<gb, 256, 0, s>>>(pl);
    pl_mcscan_count<<<ntiles, PL_T, 0, s>>>(pl.dest, N, pl.tilecnt);
    pl_mcscan_prefix<<<1, 1024, 0, s>>>(pl.tilecnt, ntiles);
    pl_mcscan_apply<<<ntiles, PL_T, 0, s>>>(pl.dest, N, pl.tilecnt, pl.prefix);
    pl_finish_kernel<<<gb, 256, 0, s>>>(pl, anc_all_dev, place_dev, place_new_dev, h->logi_new, h->src, h->ancl, h->slot, h->alive);
    // ---- this rank's gather ----
    int blocks = (int)((P + 255) / 256);
    if (blocks > h->sm_count * 16) blocks = h->sm_count * 16;
    pl_mark_kernel<<<blocks, 256, 0, s>>>(h->src, h->ancl, P, pl.rank, h->slot, h->alive, h->extra, h->defer, h->nfolv, h->leaders,
                                          h->dctl);
    fs2_iscan_sums<<<h->iscan_nb, 256, 0, s>>>(h->extra, h->alive, P, S, h->iscan_bs);
    fs2_iscan_prefix<<<1, 1024, 0, s>>>(h->iscan_bs, h->iscan_nb, h->ncopies);
    h->launches += 15;
    FS2_CUDA(cudaGetLastError());
    // enough free slots?  (spare slots = P make this certain; a smaller pool may fall short)
    FS2_CUDA(cudaMemcpyAsync(h->h_flags, h->ncopies, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
    unsigned long long info[4] = {0, 0, 0, 0};
    FS2_CUDA(cudaMemcpyAsync(info, pl.info, sizeof(info), cudaMemcpyDeviceToHost, s));
    FS2_CUDA(cudaStreamSynchronize(s));
    if (info_host) { info_host[0] = (int64_t)info[0]; info_host[1] = (int64_t)info[1]; }
    if (h->h_flags[0] > h->h_flags[1]) return FS2_ERR_NOMEM;
    fs2_iscan_apply<<<h->iscan_nb, 256, 0, s>>>(h->extra, h->alive, P, S, h->iscan_bs, h->tasks, h->freeslot);
    const Fs2Peers *peers = (const Fs2Peers *)h->peers_dev;
    pl_pose_kernel<<<blocks, 256, 0, s>>>(h->src, h->extra, P, peers, h->x2, h->y2, h->yaw2, h->w2, h->count2, h->slot2);
    int cblocks = h->sm_count * 8;
    int64_t need = (P + 7) / 8;
    if ((int64_t)cblocks > need) cblocks = (int)need;
    const int32_t *dctl = h->defer ? h->dctl : nullptr;
    pl_copy_kernel<<<cblocks, 256, 0, s>>>(0, h->tasks, h->freeslot, h->ncopies, h->src, h->ancl, P, pl.rank, peers, h->lm, h->lcap,
                                          h->slot2, h->count2, dctl, h->nfolv);
    pl_copy_kernel<<<cblocks, 256, 0, s>>>(1, h->tasks, h->freeslot, h->ncopies, h->src, h->ancl, P, pl.rank, peers, h->lm, h->lcap,
                                          h->slot2, h->count2, dctl, h->nfolv);
    if (h->defer) h->defer_pending = 1;        // (the followers' maps: written by the next update launch, or by sync_maps)
    h->launches += 4;
    FS2_CUDA(cudaGetLastError());
    return FS2_OK;
}

// after every rank has finished reading (barrier by the caller): publish the new poses and the new logical ids
extern "C" int fs2_place_commit(fs2_handle h, void *stream)
{
    if (!h || !h->pl_on) return FS2_ERR_INVALID;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    int r = commit_gather(h, s);
    if (r != FS2_OK) return r;
    FS2_CUDA(cudaMemcpyAsync(h->logi, h->logi_new, sizeof(int32_t) * (size_t)h->P, cudaMemcpyDeviceToDevice, s));
    return FS2_OK;
}

// A7-A10 as one asynchronous chain with the resampling decision on the device (fs2_step.cuh)
static int launch_finish(fs2_handle h, double u0, int32_t *anc, cudaStream_t s)
{
    const int64_t P = h->P, S = h->S;
    const int nb = (int)((P + FS2_SCAN_B - 1) / FS2_SCAN_B);
    int rb = (int)((P + FS2_RED_THREADS - 1) / FS2_RED_THREADS);
    if (rb > h->red_blocks) rb = h->red_blocks;
    { int r = sync_maps(h, s); if (r != FS2_OK) return r; }      // (two resamples without an update in between)
    fs2_weight_total_kernel<<<rb, FS2_RED_THREADS, 0, s>>>(h->w, P, h->partial, h->counters, h->stats, h->ctl);
    int nbk = nb < h->red_blocks ? nb : h->red_blocks;
    fs2_normalize_scan_kernel<<<nbk, FS2_SCAN_T, 0, s>>>(h->w, h->x, h->y, h->yaw, P, h->Pglobal, h->stats, h->partial_sq,
                                                        h->partial_best, h->counters + 1, h->stats, h->bsum, h->bpre, nb, h->ctl, 1);
    // ---- everything below returns at once unless the device decided to resample (fast_slam_2.py:62) ----
    {
        const int ng = (nb + FS2_GRP - 1) / FS2_GRP;
        Fs2ScanFused fu;
        fu.ctl = h->ctl; fu.u0 = u0; fu.ancestor = anc; fu.cum = h->cumsum; fu.used = h->alive; fu.S = S;
        fu.is_state = h->is_state; fu.is_tiles = h->is_tiles; fu.is_ticket = h->is_ticket;
        fs2_scan_groupfunc<<<ng, FS2_GRP_T, 0, s>>>(h->w, P, nb, h->bsum, h->bpre, make_scan(h), h->counters + 2, fu);
        fs2_scan_emit<<<ng, FS2_GRP_T, 0, s>>>(h->w, P, nb, make_scan(h), h->cumsum, h->ctl);
    }
    int sblocks = (int)((P + 255) / 256);
    if (sblocks > h->sm_count * 16) sblocks = h->sm_count * 16;
    fs2_search_mark_kernel<<<sblocks, 256, 0, s>>>(h->cumsum, P, u0, anc, h->slot, h->alive, h->extra, h->ctl, h->defer, h->leaders,
                                                  h->nfolv, h->dctl);
    fs2_iscan_kernel<<<h->is_tiles, FS2_IS_T, 0, s>>>(h->extra, h->alive, P, S, h->is_state, h->is_ticket, h->tasks, h->freeslot,
                                                     h->ncopies, h->is_tiles, h->ctl, h->stats);
    int cblocks = h->sm_count * 8;
    int64_t need = (P + 7) / 8;
    if ((int64_t)cblocks > need) cblocks = (int)need;
    fs2_gather_kernel<<<cblocks, 256, 0, s>>>(anc, h->extra, P, h->x, h->y, h->yaw, h->w, h->count, h->slot, h->x2, h->y2, h->yaw2,
                                             h->w2, h->count2, h->slot2, h->tasks, h->freeslot, h->ncopies, h->lm, h->lcap, h->ctl,
                                             h->defer ? h->dctl : nullptr);
    if (h->defer) h->defer_pending = 1;
    fs2_commit_estimate_kernel<<<rb, FS2_RED_THREADS, 0, s>>>(h->x, h->y, h->yaw, h->w, h->count, h->slot, h->x2, h->y2, h->yaw2,
                                                             h->w2, h->count2, h->slot2, P, h->partial_best, h->counters + 3,
                                                             h->stats, h->ctl, h->defer ? h->dctl : nullptr);
    h->launches += 8;
    FS2_CUDA(cudaGetLastError());
    return FS2_OK;
}

extern "C" int fs2_finish_step(fs2_handle h, double u0, int32_t *ancestor_dev, void *stream)
{
    if (!h) return FS2_ERR_INVALID;
    if (h->Pglobal != h->P) return FS2_ERR_UNSUPPORTED;   // sharded filters are driven stage-wise (fast_slam_b200/dist.py)
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    return launch_finish(h, u0, ancestor_dev ? ancestor_dev : h->ancestor, (cudaStream_t)stream);
}

extern "C" int fs2_step_host(fs2_handle h, double rotation, double translation, const double *obs_host, int32_t M,
                             const double *noise_host, uint64_t step, double u0, int32_t *assoc_dev,
                             int32_t *ancestor_dev, fs2_step_result *out, void *stream)
{
    if (!h || !out) return FS2_ERR_INVALID;
    if (h->Pglobal != h->P) return FS2_ERR_UNSUPPORTED;   // sharded filters are driven stage-wise (fast_slam_b200/dist.py)
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    int r;
    if (noise_host) {
        FS2_CUDA(cudaMemcpyAsync(h->noise, noise_host, sizeof(double) * (size_t)h->P, cudaMemcpyHostToDevice, s));
    } else {
        double sigma = (rotation != 0.0) ? h->cfg.rotation_noise : h->cfg.translation_noise;  // fast_slam_2.py:77-82
        if ((r = fs2_draw_noise(h, sigma, step, h->noise, s)) != FS2_OK) return r;
    }
    if ((r = launch_update(h, 1, rotation, translation, h->noise, obs_host, M, assoc_dev, s)) != FS2_OK) return r;
    if ((r = launch_finish(h, u0, ancestor_dev ? ancestor_dev : h->ancestor, s)) != FS2_OK) return r;
    // the one read-back and the one synchronisation of the step
    FS2_CUDA(cudaMemcpyAsync(h->h_stats, h->stats, sizeof(double) * FS2_STATS_LEN, cudaMemcpyDeviceToHost, s));
    FS2_CUDA(cudaStreamSynchronize(s));
    out->total = h->h_stats[FS2_STAT_TOTAL];
    out->neff = h->h_stats[FS2_STAT_NEFF];                 // before the resample, as fast_slam_2.py:59 computes it
    out->resampled = h->h_stats[FS2_STAT_RESAMPLED] != 0.0 ? 1 : 0;
    out->status_or = 0;
    out->x = h->h_stats[FS2_STAT_EST_X];
    out->y = h->h_stats[FS2_STAT_EST_Y];
    out->yaw = h->h_stats[FS2_STAT_EST_YAW];
    return FS2_OK;
}

extern "C" int fs2_upload_state(fs2_handle h, const double *x, const double *y, const double *yaw, const double *w,
                                const int32_t *count, const double *lm, void *stream)
{
    if (!h) return FS2_ERR_INVALID;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    { int r_ = sync_maps(h, (cudaStream_t)stream); if (r_ != FS2_OK) return r_; }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t P = (size_t)h->P;
    // identity slots first (reset also clears status), then overwrite what the caller provides
    int r = fs2_reset(h, stream);
    if (r != FS2_OK) return r;
    if (x) FS2_CUDA(cudaMemcpyAsync(h->x, x, sizeof(double) * P, cudaMemcpyHostToDevice, s));
    if (y) FS2_CUDA(cudaMemcpyAsync(h->y, y, sizeof(double) * P, cudaMemcpyHostToDevice, s));
    if (yaw) FS2_CUDA(cudaMemcpyAsync(h->yaw, yaw, sizeof(double) * P, cudaMemcpyHostToDevice, s));
    if (w) FS2_CUDA(cudaMemcpyAsync(h->w, w, sizeof(double) * P, cudaMemcpyHostToDevice, s));
    if (count) FS2_CUDA(cudaMemcpyAsync(h->count, count, sizeof(int32_t) * P, cudaMemcpyHostToDevice, s));
    if (lm) FS2_CUDA(cudaMemcpyAsync(h->lm, lm, sizeof(double) * P * 6 * (size_t)h->lcap, cudaMemcpyHostToDevice, s));
    FS2_CUDA(cudaStreamSynchronize(s));
    return FS2_OK;
}

extern "C" int fs2_download_state(fs2_handle h, double *x, double *y, double *yaw, double *w, int32_t *count,
                                  double *lm, int32_t *status, void *stream)
{
    if (!h) return FS2_ERR_INVALID;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    { int r_ = sync_maps(h, (cudaStream_t)stream); if (r_ != FS2_OK) return r_; }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t P = (size_t)h->P;
    if (x) FS2_CUDA(cudaMemcpyAsync(x, h->x, sizeof(double) * P, cudaMemcpyDeviceToHost, s));
    if (y) FS2_CUDA(cudaMemcpyAsync(y, h->y, sizeof(double) * P, cudaMemcpyDeviceToHost, s));
    if (yaw) FS2_CUDA(cudaMemcpyAsync(yaw, h->yaw, sizeof(double) * P, cudaMemcpyDeviceToHost, s));
    if (w) FS2_CUDA(cudaMemcpyAsync(w, h->w, sizeof(double) * P, cudaMemcpyDeviceToHost, s));
    if (count) FS2_CUDA(cudaMemcpyAsync(count, h->count, sizeof(int32_t) * P, cudaMemcpyDeviceToHost, s));
    if (status) FS2_CUDA(cudaMemcpyAsync(status, h->status, sizeof(int32_t) * P, cudaMemcpyDeviceToHost, s));
    if (lm) {
        // resolve the slot indirection on the device into a canonical-order staging buffer
        double *tmp = nullptr;
        size_t bytes = sizeof(double) * P * 6 * (size_t)h->lcap;
        cudaError_t e = cudaMalloc((void **)&tmp, bytes);
        if (e != cudaSuccess) { (void)cudaGetLastError(); return FS2_ERR_NOMEM; }
        int blocks = (int)(P < 4096 ? P : 4096);
        fs2_pack_maps_kernel<<<blocks, 256, 0, s>>>(h->lm, h->slot, nullptr, (int64_t)P, h->lcap, tmp);
        h->launches++;
        cudaError_t e2 = cudaMemcpyAsync(lm, tmp, bytes, cudaMemcpyDeviceToHost, s);
        cudaError_t e3 = cudaStreamSynchronize(s);
        cudaFree(tmp);
        if (e2 != cudaSuccess || e3 != cudaSuccess) { snprintf(g_cuda_err, sizeof(g_cuda_err), "download maps failed"); return FS2_ERR_CUDA; }
    }
    FS2_CUDA(cudaStreamSynchronize(s));
    return FS2_OK;
}

// selected particles only (full-size parity checks sample a few thousand particles of a 2^20 set)
extern "C" int fs2_download_particles(fs2_handle h, const int64_t *sel_host, int64_t nsel, double *x, double *y,
                                      double *yaw, double *w, int32_t *count, double *lm, void *stream)
{
    if (!h || !sel_host || nsel <= 0) return FS2_ERR_INVALID;
    FS2_CUDA(cudaSetDevice(h->cfg.device));
    { int r_ = sync_maps(h, (cudaStream_t)stream); if (r_ != FS2_OK) return r_; }
    cudaStream_t s = (cudaStream_t)stream;
    for (int64_t r = 0; r < nsel; ++r) {
        int64_t p = sel_host[r];
        if (p < 0 || p >= h->P) return FS2_ERR_INVALID;
        if (x) FS2_CUDA(cudaMemcpyAsync(x + r, h->x + p, sizeof(double), cudaMemcpyDeviceToHost, s));
        if (y) FS2_CUDA(cudaMemcpyAsync(y + r, h->y + p, sizeof(double), cudaMemcpyDeviceToHost, s));
        if (yaw) FS2_CUDA(cudaMemcpyAsync(yaw + r, h->yaw + p, sizeof(double), cudaMemcpyDeviceToHost, s));
        if (w) FS2_CUDA(cudaMemcpyAsync(w + r, h->w + p, sizeof(double), cudaMemcpyDeviceToHost, s));
        if (count) FS2_CUDA(cudaMemcpyAsync(count + r, h->count + p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    }
    if (lm) {
        int64_t *sel_dev = nullptr;
        double *tmp = nullptr;
        size_t bytes = sizeof(double) * (size_t)nsel * 6 * (size_t)h->lcap;
        if (cudaMalloc((void **)&sel_dev, sizeof(int64_t) * (size_t)nsel) != cudaSuccess) { (void)cudaGetLastError(); return FS2_ERR_NOMEM; }
        if (cudaMalloc((void **)&tmp, bytes) != cudaSuccess) { (void)cudaGetLastError(); cudaFree(sel_dev); return FS2_ERR_NOMEM; }
        cudaMemcpyAsync(sel_dev, sel_host, sizeof(int64_t) * (size_t)nsel, cudaMemcpyHostToDevice, s);
        int blocks = (int)(nsel < 4096 ? nsel : 4096);
        fs2_pack_maps_kernel<<<blocks, 256, 0, s>>>(h->lm, h->slot, sel_dev, nsel, h->lcap, tmp);
        h->launches++;
        cudaError_t e2 = cudaMemcpyAsync(lm, tmp, bytes, cudaMemcpyDeviceToHost, s);
        cudaError_t e3 = cudaStreamSynchronize(s);
        cudaFree(tmp);
        cudaFree(sel_dev);
        if (e2 != cudaSuccess || e3 != cudaSuccess) { snprintf(g_cuda_err, sizeof(g_cuda_err), "download particles failed"); return FS2_ERR_CUDA; }
    }
    FS2_CUDA(cudaStreamSynchronize(s));
    return FS2_OK;
}

// ------------------------------------------------------------------------------------------------------
// scan front-end (rows A11-A15): LandmarkUtils.get_measurements_to_landmarks for a batch of scans
// ------------------------------------------------------------------------------------------------------
static bool g_fe_tables_ready[64] = {false};

static int fe_prepare_tables(int device)
{
    if (device >= 0 && device < 64 && g_fe_tables_ready[device]) return FS2_OK;
    // cv::createTrigTable: float angle accumulated in float, sin/cos evaluated in double (irho = 1)
    float ts[FE_NUMANGLE], tc[FE_NUMANGLE];
    float ang = 0.0f;
    const float theta = (float)(3.14159265358979323846 / 180.0);
    for (int n = 0; n < FE_NUMANGLE; ++n) {
        ts[n] = (float)(sin((double)ang) * 1.0);
        tc[n] = (float)(cos((double)ang) * 1.0);
        ang += theta;
    }
    FS2_CUDA(cudaMemcpyToSymbol(fe_tab_sin, ts, sizeof(ts)));
    FS2_CUDA(cudaMemcpyToSymbol(fe_tab_cos, tc, sizeof(tc)));
    if (device >= 0 && device < 64) g_fe_tables_ready[device] = true;
    return FS2_OK;
}

extern "C" int fs2_frontend_max_measurements(void) { return FE_MAX_K; }

// scans_host: [B][N][2] points, or (ranges_host != null) [B][N] beam ranges + [N] beam angles
// Scratch of the front-end, kept per device between calls and only ever grown: a batch of 256 scans needs ~0.7 GB of
// Hough accumulators, and allocating and freeing that on every call cost several times the kernels' own time.
#include <mutex>
static std::mutex g_fe_mutex;
#define FE_SLOTS 26            // 0-15 and 21-24: front-end, 16-20: fs2_icp
static void *g_fe_ptr[64][FE_SLOTS];
static size_t g_fe_cap[64][FE_SLOTS];

static cudaError_t fe_buf(int device, int slot, void **out, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    if (g_fe_cap[device][slot] < bytes) {
        if (g_fe_ptr[device][slot]) cudaFree(g_fe_ptr[device][slot]);
        g_fe_ptr[device][slot] = nullptr;
        g_fe_cap[device][slot] = 0;
        const size_t want = bytes + bytes / 4;            // head-room: batches of similar scans differ a little
        cudaError_t e = cudaMalloc(&g_fe_ptr[device][slot], want);
        if (e != cudaSuccess) { (void)cudaGetLastError(); e = cudaMalloc(&g_fe_ptr[device][slot], bytes); if (e != cudaSuccess) return e; g_fe_cap[device][slot] = bytes; }
        else g_fe_cap[device][slot] = want;
    }
    *out = g_fe_ptr[device][slot];
    return cudaSuccess;
}

extern "C" int fs2_frontend_release(int32_t device)
{
    if (device < 0 || device >= 64) return FS2_ERR_INVALID;
    std::lock_guard<std::mutex> guard(g_fe_mutex);
    cudaSetDevice(device);
    for (int i = 0; i < FE_SLOTS; ++i) {
        if (g_fe_ptr[device][i]) cudaFree(g_fe_ptr[device][i]);
        g_fe_ptr[device][i] = nullptr;
        g_fe_cap[device][i] = 0;
    }
    return FS2_OK;
}

// largest double s with sqrt(s) <= eps (sqrt is correctly rounded and monotone, so "sqrt(s) <= eps" <=> "s <= this"):
// lets the kernels decide "distance <= eps" exactly as the reference's np.sqrt(...) <= eps without a square root
static double fe_sq_threshold(double eps)
{
    double s = eps * eps;
    while (sqrt(s) > eps) s = nextafter(s, 0.0);
    while (sqrt(nextafter(s, INFINITY)) <= eps) s = nextafter(s, INFINITY);
    return s;
}

extern "C" double fs2_frontend_sq_threshold(double eps) { return fe_sq_threshold(eps); }

static int fe_run(const double *scans_host, const double *ranges_host, const double *angles_host, double min_range,
                  double max_range, int32_t B, int32_t N, double sigma, int32_t device, double *meas_host, int32_t *k_host,
                  int32_t *status_host, void *stream, float *inter_host = nullptr, int32_t *ninter_host = nullptr)
{
    if ((!scans_host && !ranges_host) || !meas_host || !k_host || B <= 0 || N <= 0 || !(sigma > 0.0)) return FS2_ERR_INVALID;
    if (device < 0 || device >= 64) return FS2_ERR_INVALID;
    if (B > 65535) return FS2_ERR_UNSUPPORTED;            // the scan index is a grid's y dimension: split larger batches
    std::lock_guard<std::mutex> guard(g_fe_mutex);        // the scratch buffers below are shared by all callers
    const int radius = (int)(4.0 * sigma + 0.5);          // scipy: int(truncate * sd + 0.5)
    if (radius > 32) return FS2_ERR_UNSUPPORTED;
    FS2_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    int r = fe_prepare_tables(device);
    if (r != FS2_OK) return r;
    {   // scipy.ndimage._gaussian_kernel1d, order 0
        double w[65], sum = 0.0;
        for (int k = -radius; k <= radius; ++k) { w[k + radius] = exp(-0.5 / (sigma * sigma) * (double)(k * k)); sum += w[k + radius]; }
        for (int k = 0; k <= 2 * radius; ++k) w[k] /= sum;
        FS2_CUDA(cudaMemcpyToSymbolAsync(fe_kernel, w, sizeof(double) * (2 * radius + 1), 0, cudaMemcpyHostToDevice, s));
    }
    double *scans = nullptr, *filtered = nullptr, *meas = nullptr;
    FeGeo *geo = nullptr;
    unsigned *bitmap = nullptr;
    int *acc = nullptr, *nlines = nullptr, *kcount = nullptr, *status = nullptr, *nvalid = nullptr;
    double *ranges = nullptr, *trig = nullptr;
    float2 *lines = nullptr, *inter = nullptr;
    int *ninter = nullptr;
    const size_t pts_bytes = sizeof(double) * (size_t)B * N * 2;
    int rc = FS2_OK;
    FeGeo *hgeo = (FeGeo *)malloc(sizeof(FeGeo) * (size_t)B);
    int32_t *hstatus = (int32_t *)malloc(sizeof(int32_t) * (size_t)B);
    if (!hgeo || !hstatus) { free(hgeo); free(hstatus); return FS2_ERR_NOMEM; }
    // the fused Hough path (shared-memory accumulators, csrc/fs2_frontend.cuh) unless the scans are too long for its
    // hash set and 16-bit cells, or FS2_FE_LEGACY=1 asks for the global accumulators (kept for cross-checking)
    const char *fe_legacy = getenv("FS2_FE_LEGACY");
    const bool fused = N <= FE_FUSED_MAX_POINTS && !(fe_legacy && fe_legacy[0] == '1');
#define FE_TRY(call) do { if ((call) != cudaSuccess) { snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", #call, cudaGetErrorString(cudaGetLastError())); rc = FS2_ERR_CUDA; goto done; } } while (0)
    FE_TRY(fe_buf(device, 0, (void **)&scans, pts_bytes));
    FE_TRY(fe_buf(device, 1, (void **)&filtered, pts_bytes));
    FE_TRY(fe_buf(device, 2, (void **)&geo, sizeof(FeGeo) * (size_t)B));
    FE_TRY(fe_buf(device, 3, (void **)&lines, sizeof(float2) * (size_t)B * FE_MAX_LINES));
    FE_TRY(fe_buf(device, 4, (void **)&nlines, sizeof(int) * (size_t)B));
    FE_TRY(fe_buf(device, 5, (void **)&kcount, sizeof(int) * (size_t)B));
    FE_TRY(fe_buf(device, 6, (void **)&status, sizeof(int) * (size_t)B));
    FE_TRY(fe_buf(device, 7, (void **)&meas, sizeof(double) * (size_t)B * FE_MAX_K * 2));
    FE_TRY(cudaMemsetAsync(status, 0, sizeof(int) * (size_t)B, s));
    FE_TRY(cudaMemsetAsync(meas, 0, sizeof(double) * (size_t)B * FE_MAX_K * 2, s));
    if (ranges_host) {
        double *htrig = (double *)malloc(sizeof(double) * 2 * (size_t)N);
        if (!htrig) { rc = FS2_ERR_NOMEM; goto done; }
        for (int i = 0; i < N; ++i) { htrig[i] = cos(angles_host[i]); htrig[N + i] = sin(angles_host[i]); }   // robot.py:55-56
        cudaError_t e1 = fe_buf(device, 12, (void **)&trig, sizeof(double) * 2 * (size_t)N);
        if (e1 == cudaSuccess) e1 = cudaMemcpy(trig, htrig, sizeof(double) * 2 * (size_t)N, cudaMemcpyHostToDevice);
        free(htrig);
        FE_TRY(e1);
        FE_TRY(fe_buf(device, 8, (void **)&ranges, sizeof(double) * (size_t)B * N));
        FE_TRY(fe_buf(device, 9, (void **)&nvalid, sizeof(int) * (size_t)B));
        FE_TRY(cudaMemcpyAsync(ranges, ranges_host, sizeof(double) * (size_t)B * N, cudaMemcpyHostToDevice, s));
        fe_polar_points<<<B, FE_THREADS, 0, s>>>(ranges, trig, trig + N, N, min_range, max_range, scans, nvalid, status);
    } else {
        FE_TRY(cudaMemcpyAsync(scans, scans_host, pts_bytes, cudaMemcpyHostToDevice, s));
    }
    fe_filter_geometry<<<B, FE_THREADS, 0, s>>>(scans, N, nvalid, radius, filtered, geo);
    {
        int2 *cand = nullptr;
        int *ncand = nullptr;
        FE_TRY(fe_buf(device, 21, (void **)&cand, sizeof(int2) * (size_t)B * FE_MAX_LINES));
        FE_TRY(fe_buf(device, 22, (void **)&ncand, sizeof(int) * (size_t)B));
        FE_TRY(cudaMemsetAsync(ncand, 0, sizeof(int) * (size_t)B, s));
        if (fused) {
            // pixel list + votes and peaks in shared memory: no accumulator in global memory, no host round trip
            int2 *pix = nullptr, *pix_t = nullptr;
            int *npix = nullptr;
            FE_TRY(fe_buf(device, 15, (void **)&pix, sizeof(int2) * (size_t)B * N * 13));
            FE_TRY(fe_buf(device, 24, (void **)&pix_t, sizeof(int2) * (size_t)B * (N * 13 + 32)));
            FE_TRY(fe_buf(device, 23, (void **)&npix, sizeof(int) * (size_t)B));
            FE_TRY(cudaFuncSetAttribute(fe_raster_list, cudaFuncAttributeMaxDynamicSharedMemorySize, FE_HASH_SIZE * (int)sizeof(int)));
            FE_TRY(cudaFuncSetAttribute(fe_vote_peaks, cudaFuncAttributeMaxDynamicSharedMemorySize, FE_VP_SMEM));
            fe_raster_list<<<B, FE_LIST_THREADS, FE_HASH_SIZE * sizeof(int), s>>>(filtered, N, geo, pix, pix_t, npix, status);
            fe_vote_peaks<<<dim3(FE_VBANDS, (unsigned)B), FE_VP_THREADS, FE_VP_SMEM, s>>>(geo, pix_t, npix, N, 80, cand, ncand);   // hough_transformation.py:24
        } else {
            FE_TRY(cudaMemcpyAsync(hgeo, geo, sizeof(FeGeo) * (size_t)B, cudaMemcpyDeviceToHost, s));
            FE_TRY(cudaStreamSynchronize(s));
            long long bw = 0, ac = 0;
            for (int b = 0; b < B; ++b) {
                if (hgeo[b].width <= 0 || hgeo[b].height <= 0 || (long long)hgeo[b].width * hgeo[b].height > (1ll << 28)) { rc = FS2_ERR_INVALID; goto done; }
                hgeo[b].bitmap_off = bw;
                hgeo[b].acc_off = ac;
                bw += ((long long)hgeo[b].width * hgeo[b].height + 31) / 32;
                ac += (long long)(FE_NUMANGLE + 2) * (hgeo[b].numrho + 2);
            }
            FE_TRY(fe_buf(device, 10, (void **)&bitmap, sizeof(unsigned) * (size_t)bw));
            FE_TRY(fe_buf(device, 11, (void **)&acc, sizeof(int) * (size_t)ac));
            FE_TRY(cudaMemsetAsync(bitmap, 0, sizeof(unsigned) * (size_t)bw, s));
            FE_TRY(cudaMemsetAsync(acc, 0, sizeof(int) * (size_t)ac, s));
            FE_TRY(cudaMemcpyAsync(geo, hgeo, sizeof(FeGeo) * (size_t)B, cudaMemcpyHostToDevice, s));
            dim3 grid((unsigned)((N * 13 + FE_THREADS - 1) / FE_THREADS), (unsigned)B);
            fe_raster_vote<<<grid, FE_THREADS, 0, s>>>(filtered, N, geo, bitmap, acc);
            fe_peaks_find<<<dim3(FE_PSPLIT, (unsigned)B), FE_THREADS, 0, s>>>(geo, acc, 80, cand, ncand);    // hough_transformation.py:24
        }
        fe_peaks_rank<<<B, FE_MAX_LINES, 0, s>>>(geo, cand, ncand, lines, nlines, status);
        if (inter_host) {
            FE_TRY(fe_buf(device, 13, (void **)&inter, sizeof(float2) * (size_t)B * FE_MAX_INTER));
            FE_TRY(fe_buf(device, 14, (void **)&ninter, sizeof(int) * (size_t)B));
        }
        fe_intersect_cluster<<<B, FE_IC_THREADS, 0, s>>>(filtered, N, geo, lines, nlines, fe_sq_threshold(0.5), fe_sq_threshold(0.1),
                                                      meas, kcount, status, inter, ninter);  // landmark_utils.py:57,63
        FE_TRY(cudaGetLastError());
    }
    FE_TRY(cudaMemcpyAsync(meas_host, meas, sizeof(double) * (size_t)B * FE_MAX_K * 2, cudaMemcpyDeviceToHost, s));
    FE_TRY(cudaMemcpyAsync(k_host, kcount, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost, s));
    FE_TRY(cudaMemcpyAsync(hstatus, status, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost, s));
    if (inter_host) {
        FE_TRY(cudaMemcpyAsync(inter_host, inter, sizeof(float2) * (size_t)B * FE_MAX_INTER, cudaMemcpyDeviceToHost, s));
        FE_TRY(cudaMemcpyAsync(ninter_host, ninter, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost, s));
    }
    FE_TRY(cudaStreamSynchronize(s));
    for (int b = 0; b < B; ++b) {
        if (hstatus[b] & FE_ST_TOO_LARGE) rc = FS2_ERR_INVALID;          // an image of more than 2^28 pixels
        if (status_host) status_host[b] = hstatus[b];
    }
done:
#undef FE_TRY
    free(hgeo);
    free(hstatus);
    return rc;
}

// LineFilter.filter (algorithms/line_filter.py:12-21) on its own: B scans of N points, filtered points back
extern "C" int fs2_line_filter(const double *scans_host, int32_t B, int32_t N, double sigma, int32_t device,
                               double *filtered_host, void *stream)
{
    if (!scans_host || !filtered_host || B <= 0 || N <= 0 || !(sigma > 0.0) || device < 0 || device >= 64) return FS2_ERR_INVALID;
    const int radius = (int)(4.0 * sigma + 0.5);          // scipy: int(truncate * sd + 0.5)
    if (radius > 32) return FS2_ERR_UNSUPPORTED;
    std::lock_guard<std::mutex> guard(g_fe_mutex);
    FS2_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    double w[65], sum = 0.0;
    for (int k = -radius; k <= radius; ++k) { w[k + radius] = exp(-0.5 / (sigma * sigma) * (double)(k * k)); sum += w[k + radius]; }
    for (int k = 0; k <= 2 * radius; ++k) w[k] /= sum;
    FS2_CUDA(cudaMemcpyToSymbolAsync(fe_kernel, w, sizeof(double) * (2 * radius + 1), 0, cudaMemcpyHostToDevice, s));
    double *scans = nullptr, *filtered = nullptr;
    FeGeo *geo = nullptr;
    const size_t bytes = sizeof(double) * (size_t)B * N * 2;
    FS2_CUDA(fe_buf(device, 0, (void **)&scans, bytes));
    FS2_CUDA(fe_buf(device, 1, (void **)&filtered, bytes));
    FS2_CUDA(fe_buf(device, 2, (void **)&geo, sizeof(FeGeo) * (size_t)B));
    FS2_CUDA(cudaMemcpyAsync(scans, scans_host, bytes, cudaMemcpyHostToDevice, s));
    fe_filter_geometry<<<B, FE_THREADS, 0, s>>>(scans, N, nullptr, radius, filtered, geo);
    FS2_CUDA(cudaGetLastError());
    FS2_CUDA(cudaMemcpyAsync(filtered_host, filtered, bytes, cudaMemcpyDeviceToHost, s));
    FS2_CUDA(cudaStreamSynchronize(s));
    return FS2_OK;
}

extern "C" int fs2_frontend(const double *scans_host, int32_t B, int32_t N, double sigma, int32_t device,
                            double *meas_host, int32_t *k_host, int32_t *status_host, void *stream)
{
    if (!scans_host) return FS2_ERR_INVALID;
    return fe_run(scans_host, nullptr, nullptr, 0.0, 0.0, B, N, sigma, device, meas_host, k_host, status_host, stream);
}

extern "C" int fs2_hough_max_intersections(void) { return FE_MAX_INTER; }

extern "C" int fs2_hough_intersections(const double *points_host, int32_t B, int32_t N, int32_t device, float *inter_host,
                                       int32_t *n_inter_host, int32_t *status_host, void *stream)
{
    if (!points_host || !inter_host || !n_inter_host || B <= 0) return FS2_ERR_INVALID;
    double *meas = (double *)malloc(sizeof(double) * (size_t)B * FE_MAX_K * 2);
    int32_t *k = (int32_t *)malloc(sizeof(int32_t) * (size_t)B);
    if (!meas || !k) { free(meas); free(k); return FS2_ERR_NOMEM; }
    // the reference's method takes points that went through LineFilter already: no filter here (radius 0)
    int rc = fe_run(points_host, nullptr, nullptr, 0.0, 0.0, B, N, 0.1, device, meas, k, status_host, stream, inter_host, n_inter_host);
    free(meas); free(k);
    return rc;
}

extern "C" int fs2_frontend_polar(const double *ranges_host, const double *angles_host, int32_t B, int32_t N,
                                  double min_range, double max_range, double sigma, int32_t device, double *meas_host,
                                  int32_t *k_host, int32_t *status_host, void *stream)
{
    if (!ranges_host || !angles_host) return FS2_ERR_INVALID;
    return fe_run(nullptr, ranges_host, angles_host, min_range, max_range, B, N, sigma, device, meas_host, k_host,
                  status_host, stream);
}


// ======================================================================================================
// map clustering (row N1): LandmarkUtils.update_known_landmarks / GeometryUtils.cluster_points
// ======================================================================================================
struct KlWork {
    KlGrid g;
    KlPts pts;
    KlAcc acc;
    kl_u64 *bsum, *scan_total, *pbase;
    unsigned *counter;            // export / extract cursors
    int64_t pbase_cap;
    // sharded run (fs2_kl_shard_*): state between the stages
    kl_u64 shard_base0, shard_n_local;
    unsigned shard_total;
    int shard_stage;
    unsigned tcap;
    void *allocs[48];
    int nallocs;
};

static void kl_work_destroy(KlWork *w)
{
    if (!w) return;
    for (int i = 0; i < w->nallocs; ++i) cudaFree(w->allocs[i]);
    free(w);
}

template <typename T>
static int kl_alloc(KlWork *w, T **p, size_t n)
{
    int rc = dev_alloc(p, n ? n : 1);
    if (rc == FS2_OK) w->allocs[w->nallocs++] = (void *)*p;
    return rc;
}

static unsigned kl_env_u32(const char *name, unsigned dflt)
{
    const char *v = getenv(name);
    if (!v || !*v) return dflt;
    long long x = atoll(v);
    return (x > 0 && x < (1ll << 31)) ? (unsigned)x : dflt;
}

// tcap: tiles (rounded up to a power of two), ccap: points the exact part can hold, kcap: clusters
static int kl_work_create(KlWork **out, unsigned tcap, unsigned ccap, unsigned kcap, int64_t particles)
{
    KlWork *w = (KlWork *)calloc(1, sizeof(KlWork));
    if (!w) return FS2_ERR_NOMEM;
    unsigned t = 64;
    while (t < tcap) t <<= 1;
    w->tcap = t;
    const size_t nc = (size_t)t * KL_TC;
    int rc = FS2_OK;
#define KL_A(ptr, n) do { if (rc == FS2_OK) rc = kl_alloc(w, &(ptr), (n)); } while (0)
    KL_A(w->g.hkeys, t); KL_A(w->g.nbr, (size_t)t * KL_NB * KL_NB);
    KL_A(w->g.cnt, nc); KL_A(w->g.minidx, nc); KL_A(w->g.sx, nc); KL_A(w->g.sy, nc);
    KL_A(w->g.status, nc); KL_A(w->g.inv, nc); KL_A(w->g.parent, nc); KL_A(w->g.ncore, nc);
    KL_A(w->g.mincore, nc); KL_A(w->g.rootmin, nc); KL_A(w->g.off, nc); KL_A(w->g.cursor, nc); KL_A(w->g.cid, nc);
    KL_A(w->g.err, 4);
    KL_A(w->g.work, 2);
    KL_A(w->pts.x, ccap); KL_A(w->pts.y, ccap); KL_A(w->pts.idx, ccap); KL_A(w->pts.cell, ccap);
    KL_A(w->pts.flag, ccap); KL_A(w->pts.label, ccap);
    KL_A(w->acc.n, kcap); KL_A(w->acc.ax, kcap); KL_A(w->acc.ay, kcap); KL_A(w->acc.bxl, kcap); KL_A(w->acc.byl, kcap);
    KL_A(w->acc.bxh, kcap); KL_A(w->acc.byh, kcap); KL_A(w->acc.minidx, kcap); KL_A(w->acc.count, 4);
    const size_t nb_cells = (nc + 1023) / 1024, nb_part = ((size_t)(particles > 0 ? particles : 1) + 1023) / 1024;
    KL_A(w->bsum, nb_cells > nb_part ? nb_cells : nb_part);
    KL_A(w->scan_total, 2);
    KL_A(w->counter, 4);
    KL_A(w->pbase, (size_t)(particles > 0 ? particles : 1));
#undef KL_A
    if (rc != FS2_OK) { kl_work_destroy(w); return rc; }
    w->pbase_cap = particles;
    w->g.hmask = t - 1;
    w->pts.cap = ccap;
    w->acc.cap = kcap;
    *out = w;
    return FS2_OK;
}

static void kl_class_table(unsigned char *tab)
{
    for (int dy = -KL_WR; dy <= KL_WR; ++dy)
        for (int dx = -KL_WR; dx <= KL_WR; ++dx) {
            const int ax = abs(dx), ay = abs(dy);
            const int far2 = (ax + 1) * (ax + 1) + (ay + 1) * (ay + 1);
            const int nx = ax > 0 ? ax - 1 : 0, ny = ay > 0 ? ay - 1 : 0;
            const int near2 = nx * nx + ny * ny;
            const int r2 = KL_TS * KL_TS;                 // eps^2 in cells
            tab[(dy + KL_WR) * KL_WD + dx + KL_WR] = (far2 < r2) ? 1 : ((near2 > r2) ? 0 : 2);
        }
}

struct KlHostOut {
    double *centroids;      // [max_clusters][2]
    int64_t *members;       // [max_clusters] or null
    int32_t max_clusters;
    int32_t *n_clusters;
    fs2_kl_info *info;
};

#define KL_TRY(call)                                                                              \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", #call, cudaGetErrorString(e__));   \
            (void)cudaGetLastError();                                                             \
            return FS2_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

// everything after the source has been described; `launch_pass(op)` runs one pass over the points
// grid geometry for this eps, empty grid
static int kl_prepare(KlWork *w, double eps, cudaStream_t s)
{
    KlGrid &g = w->g;
    const unsigned T = w->tcap;
    const size_t nc = (size_t)T * KL_TC;
    g.h = eps / (double)KL_TS;
    g.inv_h = 1.0 / g.h;
    {
        int ex = 0;
        g.pow2 = (frexp(g.h, &ex) == 0.5) ? 1 : 0;
    }
    g.eps2 = eps * eps;
    unsigned char tab[KL_WD * KL_WD];
    kl_class_table(tab);
    KL_TRY(cudaMemcpyToSymbolAsync(kl_cls, tab, sizeof(tab), 0, cudaMemcpyHostToDevice, s));
    KL_TRY(cudaMemsetAsync(g.hkeys, 0xff, sizeof(kl_u64) * T, s));
    KL_TRY(cudaMemsetAsync(g.cnt, 0, sizeof(unsigned) * nc, s));
    KL_TRY(cudaMemsetAsync(g.minidx, 0xff, sizeof(kl_u64) * nc, s));
    KL_TRY(cudaMemsetAsync(g.sx, 0, sizeof(kl_u64) * nc, s));
    KL_TRY(cudaMemsetAsync(g.sy, 0, sizeof(kl_u64) * nc, s));
    KL_TRY(cudaMemsetAsync(g.err, 0, sizeof(int) * 4, s));
    KL_TRY(cudaMemsetAsync(g.work, 0, sizeof(kl_u64) * 2, s));
    KL_TRY(cudaMemsetAsync(w->acc.count, 0, sizeof(unsigned) * 4, s));
    {
        const char *v = getenv("FS2_KL_WORK");                 // exact distance tests the point-level part may spend
        const double lim = (v && *v) ? atof(v) : 3.0e10;       // (~1 s on a B200)
        g.work_limit = (kl_u64)(lim > 1.0 ? lim : 1.0);
    }
    return FS2_OK;
}

// everything that works on cells; *total = points that have to go through the point-level part
static int kl_cell_level(KlWork *w, long long min_samples, fs2_kl_info *info, unsigned *total, int *nl, cudaStream_t s)
{
    KlGrid &g = w->g;
    const unsigned T = w->tcap;
    const size_t nc = (size_t)T * KL_TC;
    g.min_samples = min_samples;
    kl_nbr_kernel<<<(T * KL_NB * KL_NB + 255) / 256, 256, 0, s>>>(g);
    kl_classify_kernel<<<T, KL_TC, 0, s>>>(g);
    kl_union_adjacent_kernel<<<T, KL_TC, 0, s>>>(g);
    kl_flatten_kernel<<<T, KL_TC, 0, s>>>(g);
    kl_mark_kernel<<<T, KL_TC, 0, s>>>(g);
    kl_flatten_kernel<<<T, KL_TC, 0, s>>>(g);
    const int nbc = (int)((nc + 1023) / 1024);
    KlInInvolved inv_in{g.cnt, g.inv, g.hkeys};
    kl_scan_sums<<<nbc, 256, 0, s>>>(inv_in, (long long)nc, w->bsum);
    kl_scan_prefix<<<1, 1024, 0, s>>>(w->bsum, nbc, w->scan_total);
    kl_scan_apply<<<nbc, 256, 0, s>>>(inv_in, (long long)nc, w->bsum, g.off);
    *nl += 9;
    KL_TRY(cudaGetLastError());
    kl_u64 h_inv = 0;
    int h_err2[2] = {0, 0};
    KL_TRY(cudaMemcpyAsync(&h_inv, w->scan_total, sizeof(kl_u64), cudaMemcpyDeviceToHost, s));
    KL_TRY(cudaMemcpyAsync(h_err2, g.err, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
    KL_TRY(cudaStreamSynchronize(s));
    const int h_err = h_err2[0];
    if (info) { info->involved_points = (int64_t)h_inv; info->err_bits = h_err; info->tiles_used = h_err2[1]; }
    if (h_err & (KL_ERR_NONFINITE | KL_ERR_RANGE)) return FS2_ERR_INVALID;      // sklearn raises on NaN / inf as well
    if (h_err) return FS2_ERR_NOMEM;
    if (h_inv > (kl_u64)w->pts.cap) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "map clustering: %llu points need the exact path, workspace holds %u "
                 "(FS2_KL_POINTS)", h_inv, w->pts.cap);
        return FS2_ERR_NOMEM;
    }
    *total = (unsigned)h_inv;
    return FS2_OK;
}

// point level (`pass(op)` runs one pass over the points that can be involved), cluster sums, centroids on the host
template <class Pass>
static int kl_finish(KlWork *w, Pass &&pass, unsigned total, int64_t n_points, int sm_count, const KlHostOut &out, int *nl,
                     cudaStream_t s)
{
    KlGrid &g = w->g;
    const unsigned T = w->tcap;
    int h_err = 0;
    const int wide = sm_count * 8;
    if (total) {
        pass(KlCompactOp{g, w->pts});
        kl_exact_core_kernel<<<(int)((total + 7) / 8 < (unsigned)wide ? (total + 7) / 8 : (unsigned)wide), 256, 0, s>>>(g, w->pts, total);
        kl_union_exact_kernel<<<T < (unsigned)wide ? T : (unsigned)wide, 256, 0, s>>>(g, w->pts);
        kl_flatten_kernel<<<T, KL_TC, 0, s>>>(g);
        *nl += 4;
    }
    kl_rootmin_kernel<<<T, KL_TC, 0, s>>>(g);
    if (total) {
        kl_border_kernel<<<(int)((total + 7) / 8 < (unsigned)wide ? (total + 7) / 8 : (unsigned)wide), 256, 0, s>>>(g, w->pts, total);
        ++*nl;
    }
    const KlAcc &a = w->acc;
    const size_t kb = sizeof(kl_u64) * a.cap;
    KL_TRY(cudaMemsetAsync(a.n, 0, kb, s)); KL_TRY(cudaMemsetAsync(a.ax, 0, kb, s)); KL_TRY(cudaMemsetAsync(a.ay, 0, kb, s));
    KL_TRY(cudaMemsetAsync(a.bxl, 0, kb, s)); KL_TRY(cudaMemsetAsync(a.byl, 0, kb, s));
    KL_TRY(cudaMemsetAsync(a.bxh, 0, kb, s)); KL_TRY(cudaMemsetAsync(a.byh, 0, kb, s));
    kl_cluster_ids_kernel<<<T, KL_TC, 0, s>>>(g, a);
    kl_acc_cells_kernel<<<T, KL_TC, 0, s>>>(g, a);
    *nl += 3;
    if (total) { kl_acc_points_kernel<<<(total + 255) / 256, 256, 0, s>>>(g, w->pts, a, total); ++*nl; }
    KL_TRY(cudaGetLastError());
    unsigned K = 0;
    KL_TRY(cudaMemcpyAsync(&K, a.count, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    KL_TRY(cudaMemcpyAsync(&h_err, g.err, sizeof(int), cudaMemcpyDeviceToHost, s));
    KL_TRY(cudaStreamSynchronize(s));
    if (h_err & KL_ERR_WORK) {
        if (out.info) out.info->err_bits = h_err;
        snprintf(g_cuda_err, sizeof(g_cuda_err), "map clustering: the exact point-level part needs more than %llu distance tests "
                 "(dense cells about eps apart); raise FS2_KL_WORK to let it run", (unsigned long long)g.work_limit);
        return FS2_ERR_UNSUPPORTED;
    }
    if (h_err || K > a.cap) {
        if (out.info) out.info->err_bits = h_err;
        return FS2_ERR_NOMEM;
    }
    struct Rec { kl_u64 minidx, n; long long ax, ay; kl_u64 bxl, byl; long long bxh, byh; };
    Rec *rec = (Rec *)malloc(sizeof(Rec) * (K ? K : 1));
    kl_u64 *tmp = (kl_u64 *)malloc(sizeof(kl_u64) * (K ? K : 1) * 8);
    if (!rec || !tmp) { free(rec); free(tmp); return FS2_ERR_NOMEM; }
    const void *srcs[8] = {a.minidx, a.n, a.ax, a.ay, a.bxl, a.byl, a.bxh, a.byh};
    if (K) {
        for (int f = 0; f < 8; ++f) {
            cudaError_t e = cudaMemcpyAsync(tmp + (size_t)f * K, srcs[f], sizeof(kl_u64) * K, cudaMemcpyDeviceToHost, s);
            if (e != cudaSuccess) { free(rec); free(tmp); snprintf(g_cuda_err, sizeof(g_cuda_err), "cluster download: %s", cudaGetErrorString(e)); return FS2_ERR_CUDA; }
        }
        if (cudaStreamSynchronize(s) != cudaSuccess) { free(rec); free(tmp); return FS2_ERR_CUDA; }
    }
    for (unsigned k = 0; k < K; ++k) {
        rec[k].minidx = tmp[k]; rec[k].n = tmp[(size_t)K + k];
        rec[k].ax = (long long)tmp[(size_t)2 * K + k]; rec[k].ay = (long long)tmp[(size_t)3 * K + k];
        rec[k].bxl = tmp[(size_t)4 * K + k]; rec[k].byl = tmp[(size_t)5 * K + k];
        rec[k].bxh = (long long)tmp[(size_t)6 * K + k]; rec[k].byh = (long long)tmp[(size_t)7 * K + k];
    }
    // label order = order of the clusters' lowest core point (dbscan_inner walks the points in index order)
    qsort(rec, K, sizeof(Rec), [](const void *p, const void *q) {
        const kl_u64 a1 = ((const Rec *)p)->minidx, b1 = ((const Rec *)q)->minidx;
        return a1 < b1 ? -1 : (a1 > b1 ? 1 : 0);
    });
    kl_u64 members = 0;
    for (unsigned k = 0; k < K; ++k) {
        members += rec[k].n;
        if ((int32_t)k >= out.max_clusters) continue;
        const __int128 bx = ((__int128)rec[k].bxh << 64) | (__int128)rec[k].bxl;
        const __int128 by = ((__int128)rec[k].byh << 64) | (__int128)rec[k].byl;
        const __int128 tx = ((__int128)rec[k].ax << KL_FIX) + bx, ty = ((__int128)rec[k].ay << KL_FIX) + by;
        const long double scale = (long double)g.h / (long double)(1ull << KL_FIX) / (long double)rec[k].n;
        out.centroids[2 * k] = (double)((long double)tx * scale);
        out.centroids[2 * k + 1] = (double)((long double)ty * scale);
        if (out.members) out.members[k] = (int64_t)rec[k].n;
    }
    free(rec); free(tmp);
    *out.n_clusters = (int32_t)K;
    if (out.info) {
        out.info->n_points = n_points;
        out.info->min_samples = g.min_samples;
        out.info->noise_points = n_points - (int64_t)members;
        out.info->clusters = (int32_t)K;
        out.info->tiles = (int32_t)T;
    }
    return (int32_t)K > out.max_clusters ? FS2_ERR_NOMEM : FS2_OK;
}

// one device, all points local: `count(g)` runs pass 1, `pass(op)` any later pass over the points
template <class Count, class Pass>
static int kl_run(KlWork *w, Count &&count, Pass &&pass, int64_t n_points, double eps, long long min_samples, int sm_count,
                  const KlHostOut &out, int64_t *launches, cudaStream_t s)
{
    int rc = kl_prepare(w, eps, s);
    if (rc != FS2_OK) return rc;
    int nl = 0;
    const bool prof = getenv("FS2_KL_PROFILE") != nullptr;      // stage times on stderr (diagnostics only)
    cudaEvent_t ev[4];
    int nev = 0;
    auto mark = [&]() { if (prof && nev < 4) { cudaEventCreate(&ev[nev]); cudaEventRecord(ev[nev], s); ++nev; } };
    mark();
    count(w->g); ++nl;
    mark();
    unsigned total = 0;
    rc = kl_cell_level(w, min_samples, out.info, &total, &nl, s);
    mark();
    if (rc == FS2_OK) rc = kl_finish(w, pass, total, n_points, sm_count, out, &nl, s);
    mark();
    if (launches) *launches += nl;
    if (prof && nev == 4) {
        float t01, t12, t23;
        cudaEventSynchronize(ev[3]);
        cudaEventElapsedTime(&t01, ev[0], ev[1]); cudaEventElapsedTime(&t12, ev[1], ev[2]); cudaEventElapsedTime(&t23, ev[2], ev[3]);
        fprintf(stderr, "fs2_kl: count pass %.3f ms, cell level %.3f ms, point level + sums %.3f ms (%u involved points)\n",
                t01, t12, t23, total);
    }
    for (int i = 0; i < nev; ++i) cudaEventDestroy(ev[i]);
    return rc;
}

extern "C" int fs2_known_landmarks(fs2_handle h, double eps, double min_samples_frac, int64_t min_samples, int32_t max_clusters,
                                   double *centroids_host, int64_t *members_host, int32_t *n_clusters, fs2_kl_info *info,
                                   void *stream)
{
    if (!h || !(eps > 0.0) || max_clusters < 0 || !n_clusters || (max_clusters > 0 && !centroids_host)) return FS2_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    KL_TRY(cudaSetDevice(h->cfg.device));
    { int r_ = sync_maps(h, (cudaStream_t)stream); if (r_ != FS2_OK) return r_; }
    fs2_kl_info local;
    if (!info) info = &local;
    // The grid starts small (2048 tiles of eps x eps: every per-call clear and per-tile kernel scales with it) and is
    // rebuilt larger when the maps fill more than a quarter of it (or overflow it); FS2_KL_TILES pins the size.
    const unsigned pinned = kl_env_u32("FS2_KL_TILES", 0);
    for (;;) {
        if (!h->kl) {
            int rc = kl_work_create(&h->kl, pinned ? pinned : 2048u, kl_env_u32("FS2_KL_POINTS", 1u << 23),
                                    kl_env_u32("FS2_KL_CLUSTERS", 1u << 16), h->P);
            if (rc != FS2_OK) { h->kl = nullptr; return rc; }
        }
        KlWork *w = h->kl;
        memset(info, 0, sizeof(*info));
        *n_clusters = 0;
        // point index of landmark j of particle p = (landmarks of the particles before p) + j   (landmark_utils.py:125-128)
        const int nbp = (int)((h->P + 1023) / 1024);
        KlInCount cin{h->count};
        kl_scan_sums<<<nbp, 256, 0, s>>>(cin, (long long)h->P, w->bsum);
        kl_scan_prefix<<<1, 1024, 0, s>>>(w->bsum, nbp, w->scan_total + 1);
        kl_scan_apply<<<nbp, 256, 0, s>>>(cin, (long long)h->P, w->bsum, w->pbase);
        h->launches += 3;
        kl_u64 N = 0;
        KL_TRY(cudaMemcpyAsync(&N, w->scan_total + 1, sizeof(kl_u64), cudaMemcpyDeviceToHost, s));
        KL_TRY(cudaStreamSynchronize(s));
        long long ms = min_samples;
        if (ms <= 0) {
            const double avg = (double)N / (double)h->Pglobal;      // len(all_landmarks) / len(particles)
            ms = (long long)(avg * min_samples_frac);               // int(avg_landmarks * 0.7)
            info->n_points = (int64_t)N; info->min_samples = ms;
            if (ms < 1) {                                           // landmark_utils.py:133-134: leave known_landmarks alone
                info->skipped = 1;
                *n_clusters = -1;
                return FS2_OK;
            }
        }
        KlSrcState src{h->lm, h->slot, h->count, w->pbase, 0ull, h->P, h->lcap};
        const int blocks = h->sm_count * 8;
        KlHostOut out{centroids_host, members_host, max_clusters, n_clusters, info};
        KL_TRY(cudaFuncSetAttribute(kl_count_state_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KL_CACHE_BYTES));
        int rc = kl_run(w, [&](const KlGrid &g) { kl_count_state_kernel<<<h->sm_count, KL_COUNT_THREADS, KL_CACHE_BYTES, s>>>(src, g); },
                        [&](auto op) { kl_pass_state<<<blocks, 256, 0, s>>>(src, op); }, (int64_t)N, eps, ms, h->sm_count, out,
                        &h->launches, s);
        const bool full = rc == FS2_ERR_NOMEM && (info->err_bits & KL_ERR_TILES);
        const bool crowded = rc == FS2_OK && 4u * (unsigned)info->tiles_used > w->tcap;     // keep the hash under a quarter full
        if ((full || crowded) && !pinned && w->tcap < 65536u) {
            unsigned bigger = w->tcap * (full ? 4u : 2u);
            while (crowded && bigger < 65536u && 4u * (unsigned)info->tiles_used > bigger) bigger *= 2u;
            const unsigned ccap = w->pts.cap, kcap = w->acc.cap;
            kl_work_destroy(w);
            h->kl = nullptr;
            const int rc2 = kl_work_create(&h->kl, bigger, ccap, kcap, h->P);
            if (rc2 != FS2_OK) { h->kl = nullptr; return full ? rc2 : rc; }
            if (full) continue;                                    // this call has no result yet: run again
        }
        return rc;
    }
}

extern "C" int fs2_cluster_points(const double *xy_host, int64_t n, double eps, int64_t min_samples, int32_t device,
                                  int32_t max_clusters, double *centroids_host, int64_t *members_host, int32_t *n_clusters,
                                  fs2_kl_info *info)
{
    if (n < 0 || (n > 0 && !xy_host) || !(eps > 0.0) || min_samples < 1 || max_clusters < 0 || !n_clusters ||
        (max_clusters > 0 && !centroids_host) || n >= (1ll << 31))
        return FS2_ERR_INVALID;
    if (info) memset(info, 0, sizeof(*info));
    *n_clusters = 0;
    if (n == 0) return FS2_OK;
    KL_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    KL_TRY(cudaGetDeviceProperties(&prop, device));
    KlWork *w = nullptr;
    unsigned tcap = 2 * (unsigned)(n < 32768 ? n : 32768);
    tcap = kl_env_u32("FS2_KL_TILES", tcap);
    int rc = kl_work_create(&w, tcap, (unsigned)n, (unsigned)(n < (1 << 16) ? n : (1 << 16)), 0);
    if (rc != FS2_OK) return rc;
    double *xy = nullptr;
    rc = kl_alloc(w, &xy, (size_t)n * 2);
    if (rc == FS2_OK) {
        cudaStream_t s = 0;
        cudaError_t e = cudaMemcpyAsync(xy, xy_host, sizeof(double) * 2 * (size_t)n, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) {
            snprintf(g_cuda_err, sizeof(g_cuda_err), "cluster_points upload: %s", cudaGetErrorString(e));
            rc = FS2_ERR_CUDA;
        } else {
            KlSrcFlat src{xy, n};
            const int blocks = (int)((n + 255) / 256 < prop.multiProcessorCount * 8 ? (n + 255) / 256 : prop.multiProcessorCount * 8);
            KlHostOut out{centroids_host, members_host, max_clusters, n_clusters, info};
            const int cblocks = (int)((n + KL_COUNT_THREADS - 1) / KL_COUNT_THREADS < prop.multiProcessorCount
                                          ? (n + KL_COUNT_THREADS - 1) / KL_COUNT_THREADS : prop.multiProcessorCount);
            cudaFuncSetAttribute(kl_count_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KL_CACHE_BYTES);
            rc = kl_run(w, [&](const KlGrid &g) { kl_count_flat_kernel<<<cblocks, KL_COUNT_THREADS, KL_CACHE_BYTES, s>>>(src, g); },
                        [&](auto op) { kl_pass_flat<<<blocks, 256, 0, s>>>(src, op); }, n, eps, min_samples,
                        prop.multiProcessorCount, out, nullptr, s);
        }
    }
    cudaDeviceSynchronize();
    kl_work_destroy(w);
    return rc;
}


// ---- the same over a sharded filter: one call per stage, the exchanges between them belong to the caller --------
static int kl_ensure(fs2_handle h)
{
    if (h->kl) return FS2_OK;
    int rc = kl_work_create(&h->kl, kl_env_u32("FS2_KL_TILES", 16384), kl_env_u32("FS2_KL_POINTS", 1u << 23),
                            kl_env_u32("FS2_KL_CLUSTERS", 1u << 16), h->P);
    if (rc != FS2_OK) h->kl = nullptr;
    return rc;
}

extern "C" int fs2_kl_record_bytes(void) { return (int)(KL_REC_WORDS * sizeof(kl_u64)); }

extern "C" int fs2_kl_shard_begin(fs2_handle h, double eps, int64_t *n_local_points, void *stream)
{
    if (!h || !(eps > 0.0) || !n_local_points) return FS2_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    KL_TRY(cudaSetDevice(h->cfg.device));
    { int r_ = sync_maps(h, (cudaStream_t)stream); if (r_ != FS2_OK) return r_; }
    int rc = kl_ensure(h);
    if (rc != FS2_OK) return rc;
    KlWork *w = h->kl;
    const int nbp = (int)((h->P + 1023) / 1024);
    KlInCount cin{h->count};
    kl_scan_sums<<<nbp, 256, 0, s>>>(cin, (long long)h->P, w->bsum);
    kl_scan_prefix<<<1, 1024, 0, s>>>(w->bsum, nbp, w->scan_total + 1);
    kl_scan_apply<<<nbp, 256, 0, s>>>(cin, (long long)h->P, w->bsum, w->pbase);
    h->launches += 3;
    kl_u64 N = 0;
    KL_TRY(cudaMemcpyAsync(&N, w->scan_total + 1, sizeof(kl_u64), cudaMemcpyDeviceToHost, s));
    KL_TRY(cudaStreamSynchronize(s));
    rc = kl_prepare(w, eps, s);
    if (rc != FS2_OK) return rc;
    w->shard_n_local = N;
    w->shard_stage = 1;
    *n_local_points = (int64_t)N;
    return FS2_OK;
}

// placed shards (fs2_place.cuh): the global index of a point follows the LOGICAL particle order, which the caller
// knows (exclusive prefix of the map lengths in logical order); replaces the local prefix fs2_kl_shard_begin computed.
// Call between fs2_kl_shard_begin and fs2_kl_shard_count(h, 0, ...).
extern "C" int fs2_kl_shard_set_bases(fs2_handle h, const int64_t *bases_dev, void *stream)
{
    if (!h || !h->kl || h->kl->shard_stage != 1 || !bases_dev) return FS2_ERR_INVALID;
    KL_TRY(cudaSetDevice(h->cfg.device));
    KL_TRY(cudaMemcpyAsync(h->kl->pbase, bases_dev, sizeof(kl_u64) * (size_t)h->P, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return FS2_OK;
}

__global__ void kl_count_tiles_kernel(KlGrid g, unsigned *n)
{
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t <= g.hmask && g.hkeys[t] != KL_KEY_EMPTY) atomicAdd(n, 1u);
}

extern "C" int fs2_kl_shard_count(fs2_handle h, int64_t index_offset, int32_t *n_tiles, void *stream)
{
    if (!h || !h->kl || h->kl->shard_stage != 1 || index_offset < 0 || !n_tiles) return FS2_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    KL_TRY(cudaSetDevice(h->cfg.device));
    { int r_ = sync_maps(h, (cudaStream_t)stream); if (r_ != FS2_OK) return r_; }
    KlWork *w = h->kl;
    w->shard_base0 = (kl_u64)index_offset;
    KlSrcState src{h->lm, h->slot, h->count, w->pbase, w->shard_base0, h->P, h->lcap};
    KL_TRY(cudaFuncSetAttribute(kl_count_state_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KL_CACHE_BYTES));
    KL_TRY(cudaMemsetAsync(w->counter, 0, sizeof(unsigned) * 4, s));
    kl_count_state_kernel<<<h->sm_count, KL_COUNT_THREADS, KL_CACHE_BYTES, s>>>(src, w->g);
    kl_count_tiles_kernel<<<(w->tcap + 255) / 256, 256, 0, s>>>(w->g, w->counter);
    h->launches += 2;
    KL_TRY(cudaGetLastError());
    unsigned nt = 0;
    int err = 0;
    KL_TRY(cudaMemcpyAsync(&nt, w->counter, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    KL_TRY(cudaMemcpyAsync(&err, w->g.err, sizeof(int), cudaMemcpyDeviceToHost, s));
    KL_TRY(cudaStreamSynchronize(s));
    if (err & (KL_ERR_NONFINITE | KL_ERR_RANGE)) return FS2_ERR_INVALID;
    if (err) return FS2_ERR_NOMEM;
    *n_tiles = (int32_t)nt;
    w->shard_stage = 2;
    return FS2_OK;
}

extern "C" int fs2_kl_shard_export(fs2_handle h, void *records_dev, int32_t cap_records, void *stream)
{
    if (!h || !h->kl || h->kl->shard_stage != 2 || (!records_dev && cap_records > 0) || cap_records < 0) return FS2_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    KL_TRY(cudaSetDevice(h->cfg.device));
    KlWork *w = h->kl;
    KL_TRY(cudaMemsetAsync(w->counter, 0, sizeof(unsigned) * 4, s));
    kl_export_tiles_kernel<<<w->tcap, KL_TC, 0, s>>>(w->g, (kl_u64 *)records_dev, (unsigned)cap_records, w->counter);
    h->launches += 1;
    KL_TRY(cudaGetLastError());
    w->shard_stage = 3;
    return FS2_OK;
}

extern "C" int fs2_kl_shard_merge(fs2_handle h, const void *records_dev, int32_t n_records, int64_t min_samples,
                                  int64_t *involved_points, void *stream)
{
    if (!h || !h->kl || h->kl->shard_stage != 3 || n_records < 0 || (n_records > 0 && !records_dev) || min_samples < 1 || !involved_points)
        return FS2_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    KL_TRY(cudaSetDevice(h->cfg.device));
    KlWork *w = h->kl;
    const double eps = sqrt(w->g.eps2);
    (void)eps;
    // the local counts are in the records: start from an empty grid and merge every shard's tiles, own included
    const unsigned T = w->tcap;
    const size_t nc = (size_t)T * KL_TC;
    KL_TRY(cudaMemsetAsync(w->g.hkeys, 0xff, sizeof(kl_u64) * T, s));
    KL_TRY(cudaMemsetAsync(w->g.cnt, 0, sizeof(unsigned) * nc, s));
    KL_TRY(cudaMemsetAsync(w->g.minidx, 0xff, sizeof(kl_u64) * nc, s));
    KL_TRY(cudaMemsetAsync(w->g.sx, 0, sizeof(kl_u64) * nc, s));
    KL_TRY(cudaMemsetAsync(w->g.sy, 0, sizeof(kl_u64) * nc, s));
    if (n_records) kl_merge_tiles_kernel<<<n_records, KL_TC, 0, s>>>(w->g, (const kl_u64 *)records_dev, (unsigned)n_records);
    int nl = 1;
    unsigned total = 0;
    int rc = kl_cell_level(w, min_samples, nullptr, &total, &nl, s);
    h->launches += nl;
    if (rc != FS2_OK) return rc;
    w->shard_total = total;
    *involved_points = (int64_t)total;
    w->shard_stage = 4;
    return FS2_OK;
}

extern "C" int fs2_kl_shard_extract(fs2_handle h, double *points_dev, int64_t cap, int64_t *n_points, void *stream)
{
    if (!h || !h->kl || h->kl->shard_stage != 4 || cap < 0 || (cap > 0 && !points_dev) || !n_points || cap >= (1ll << 32)) return FS2_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    KL_TRY(cudaSetDevice(h->cfg.device));
    { int r_ = sync_maps(h, (cudaStream_t)stream); if (r_ != FS2_OK) return r_; }
    KlWork *w = h->kl;
    KlSrcState src{h->lm, h->slot, h->count, w->pbase, w->shard_base0, h->P, h->lcap};
    KL_TRY(cudaMemsetAsync(w->counter, 0, sizeof(unsigned) * 4, s));
    kl_pass_state<<<h->sm_count * 8, 256, 0, s>>>(src, KlExtractOp{w->g, points_dev, (unsigned)cap, w->counter});
    h->launches += 1;
    KL_TRY(cudaGetLastError());
    unsigned n = 0;
    KL_TRY(cudaMemcpyAsync(&n, w->counter, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    KL_TRY(cudaStreamSynchronize(s));
    *n_points = (int64_t)n;                  // may exceed cap: the caller grows the buffer and calls again
    return FS2_OK;
}

extern "C" int fs2_kl_shard_finish(fs2_handle h, const double *points_dev, int64_t n_points, int64_t n_total_points,
                                   int32_t max_clusters, double *centroids_host, int64_t *members_host, int32_t *n_clusters,
                                   fs2_kl_info *info, void *stream)
{
    if (!h || !h->kl || h->kl->shard_stage != 4 || max_clusters < 0 || !n_clusters || (max_clusters > 0 && !centroids_host) ||
        n_points != (int64_t)h->kl->shard_total || (n_points > 0 && !points_dev))
        return FS2_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    KL_TRY(cudaSetDevice(h->cfg.device));
    KlWork *w = h->kl;
    if (info) { memset(info, 0, sizeof(*info)); info->involved_points = n_points; }
    *n_clusters = 0;
    KlHostOut out{centroids_host, members_host, max_clusters, n_clusters, info};
    int nl = 0;
    const int blocks = h->sm_count * 8;
    int rc = kl_finish(w, [&](auto op) { kl_pass_flat3<<<blocks, 256, 0, s>>>(points_dev, (long long)n_points, op); },
                       w->shard_total, n_total_points, h->sm_count, out, &nl, s);
    h->launches += nl;
    w->shard_stage = 0;
    return rc;
}


// ======================================================================================================
// ICP.get_transformation (row N4), batched
// ======================================================================================================
extern "C" int fs2_icp(const double *source_host, const double *target_host, int32_t B, int32_t n_source, int32_t n_target,
                       int32_t max_iterations, double threshold, int32_t device, double *rotation_host,
                       double *translation_host, int32_t *iterations_host, void *stream)
{
    if (!source_host || !target_host || !rotation_host || !translation_host || B <= 0 || n_source <= 0 || n_target <= 0 ||
        n_source > ICP_MAX_POINTS || n_target > ICP_MAX_POINTS || max_iterations < 0)
        return FS2_ERR_INVALID;
    if (device < 0 || device >= 64) return FS2_ERR_INVALID;
    std::lock_guard<std::mutex> guard(g_fe_mutex);        // scratch shared with the front-end (slots 16-20), grown on demand
    FS2_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    double *src = nullptr, *tgt = nullptr, *rot = nullptr, *tr = nullptr;
    int *it = nullptr;
    int rc = FS2_OK;
    const size_t sb = sizeof(double) * 2 * (size_t)B * n_source, tb = sizeof(double) * 2 * (size_t)B * n_target;
    const int smem = (int)(16 * ((size_t)n_source + n_target) + 4 * (size_t)n_source);
#define ICP_TRY(call) do { if ((call) != cudaSuccess) { snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", #call, cudaGetErrorString(cudaGetLastError())); rc = FS2_ERR_CUDA; goto done; } } while (0)
    ICP_TRY(fe_buf(device, 16, (void **)&src, sb));
    ICP_TRY(fe_buf(device, 17, (void **)&tgt, tb));
    ICP_TRY(fe_buf(device, 18, (void **)&rot, sizeof(double) * 4 * (size_t)B));
    ICP_TRY(fe_buf(device, 19, (void **)&tr, sizeof(double) * 2 * (size_t)B));
    ICP_TRY(fe_buf(device, 20, (void **)&it, sizeof(int) * (size_t)B));
    ICP_TRY(cudaMemcpyAsync(src, source_host, sb, cudaMemcpyHostToDevice, s));
    ICP_TRY(cudaMemcpyAsync(tgt, target_host, tb, cudaMemcpyHostToDevice, s));
    ICP_TRY(cudaFuncSetAttribute(icp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    icp_kernel<<<B, ICP_THREADS, smem, s>>>(src, tgt, n_source, n_target, max_iterations, threshold, rot, tr, it);
    ICP_TRY(cudaGetLastError());
    ICP_TRY(cudaMemcpyAsync(rotation_host, rot, sizeof(double) * 4 * (size_t)B, cudaMemcpyDeviceToHost, s));
    ICP_TRY(cudaMemcpyAsync(translation_host, tr, sizeof(double) * 2 * (size_t)B, cudaMemcpyDeviceToHost, s));
    if (iterations_host) ICP_TRY(cudaMemcpyAsync(iterations_host, it, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost, s));
    ICP_TRY(cudaStreamSynchronize(s));
done:
#undef ICP_TRY
    return rc;      // the scratch stays with the device (fs2_frontend_release frees it)
}
