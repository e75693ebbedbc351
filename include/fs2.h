/*
 * fs2.h -- C ABI of libfs2.so, the B200 (sm_100a) FastSLAM filter-step library.
 *
 * The reference (cy-rae/fast-slam) is pure Python and has no FFI of its own; the interface it exposes
 * for this path is the class `FastSLAM2` (fast_slam_2/algorithms/fast_slam_2.py:15-223) used by
 * jde_robots_main.py:13,38.  Each entry point below replaces one piece of that class; the citation
 * names the reference lines it stands in for.  INTEGRATION.md shows the ctypes stub a maintainer of
 * the reference would add to call them.
 *
 * Conventions
 *   - every function returns an int status: 0 = FS2_OK, negative = error (fs2_strerror), and never
 *     throws.  Numerical trouble inside one (particle, observation) update is NOT an error: like the
 *     reference's thread pool, which swallows the exception (fast_slam_2.py:45,53), the update is
 *     skipped and a bit is set in the particle's status word (fs2_ptrs.status, FS2_ST_*).
 *   - pointers named *_dev are device pointers on the handle's device, borrowed for the call; pointers
 *     named *_host are host pointers.  `stream` is a cudaStream_t passed as void* (0 = default stream).
 *     Calls are asynchronous on that stream unless stated otherwise.
 *   - all arithmetic state is IEEE fp64, like the reference.
 *
 * HBM store owned by a handle (P = num_particles, Lcap = landmark_capacity):
 *     x, y, yaw, w : double[P]          pose and importance weight   (particle.py:11-20)
 *     count        : int32[P]           landmarks in the particle's map
 *     lm           : double[P][Lcap][6] one contiguous map per SLOT, 48 B per landmark
 *                                       x, y, c00, c01, c10, c11  (landmark.py:13-21).  Particle p's map is
 *                                       slot p until the first resample; after that a private slot table
 *                                       maps particles to slots (copy-on-resample), so read maps through
 *                                       fs2_download_state / fs2_download_particles.
 *     status       : int32[P]           FS2_ST_* bits, sticky until fs2_reset
 */
#ifndef FS2_H
#define FS2_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FS2_ABI_VERSION 1

#define FS2_OK 0
#define FS2_ERR_INVALID (-1)     /* bad argument                                   */
#define FS2_ERR_CUDA (-2)        /* a CUDA runtime call failed (fs2_last_cuda_error) */
#define FS2_ERR_NOMEM (-3)       /* device allocation failed                       */
#define FS2_ERR_UNSUPPORTED (-4)

/* per-particle status bits */
#define FS2_ST_SINGULAR_LM 1 /* a singular landmark covariance met in association: update skipped (geometry_utils.py:22 raises) */
#define FS2_ST_SINGULAR_Q 2  /* singular observation covariance: update skipped (fast_slam_2.py:142 raises)      */
#define FS2_ST_PDF_FAILED 4  /* likelihood rejected Q: landmark replaced, weight untouched (fast_slam_2.py:156 raises) */
#define FS2_ST_MAP_FULL 8    /* Lcap reached, new landmark dropped (the reference's lists are unbounded)         */

/* fs2_config.flags */
#define FS2_FLAG_FORCE_SEQUENTIAL 1 /* always take the observation-by-observation path in fs2_update (debug / cross-check) */

typedef struct fs2_filter_s *fs2_handle;

typedef struct fs2_config {
    int64_t num_particles;        /* particles owned by this handle (this GPU's shard); config.py:7 NUM_PARTICLES on 1 GPU */
    int64_t global_particles;     /* particles over all shards (== num_particles on 1 GPU); 0 means num_particles         */
    int64_t global_offset;        /* global index of this shard's first particle                                          */
    int32_t landmark_capacity;    /* Lcap >= 1: map slots per particle                                                    */
    int32_t device;               /* CUDA device ordinal                                                                  */
    int32_t flags;                /* FS2_FLAG_*                                                                           */
    int32_t spare_slots;          /* extra map slots beyond num_particles (peer-memory gather, fs2_gather_p2p); 0 = none     */
    double translation_noise;     /* config.py:11 TRANSLATION_NOISE                                                       */
    double rotation_noise;        /* config.py:12 ROTATION_NOISE                                                          */
    double measurement_noise[4];  /* config.py:15 MEASUREMENT_NOISE, row-major 2x2                                        */
    double max_landmark_distance; /* config.py:18 MAXIMUM_LANDMARK_DISTANCE (Mahalanobis gate)                            */
    uint64_t seed;                /* stream of the device random generator (fs2_draw_noise)                               */
} fs2_config;

typedef struct fs2_ptrs {
    double *x, *y, *yaw, *w;
    int32_t *count;
    double *lm;
    int32_t *status;
    double *noise;      /* double[P] scratch the step entry points draw the motion noise into */
    double *cumsum;     /* double[global P] exact sequential running sum of the last resample  */
    int32_t *ancestor;  /* int32[P]  ancestor (global index) chosen for each local slot by the last resample */
    double *stats;      /* double[FS2_STATS_LEN] see fs2_weight_stats                           */
    int64_t num_particles;
    int32_t landmark_capacity;
    int32_t reserved;
} fs2_ptrs;

/* layout of the stats block written by fs2_normalize / read back by fs2_step_host */
#define FS2_STAT_TOTAL 0    /* sum of weights before normalisation (this shard)        */
#define FS2_STAT_SUMSQ 1    /* sum of squared weights after normalisation (this shard) */
#define FS2_STAT_NEFF 2     /* fast_slam_2.py:212-223 on this shard's numbers          */
#define FS2_STAT_WMAX 3     /* largest weight                                          */
#define FS2_STAT_ARGMAX 4   /* LOCAL index of its first occurrence                     */
#define FS2_STAT_EST_X 5    /* pose of that particle                                   */
#define FS2_STAT_EST_Y 6
#define FS2_STAT_EST_YAW 7
#define FS2_STAT_RESAMPLED 8 /* fs2_finish_step: 1 if the step resampled (decided on the device, fast_slam_2.py:62) */
#define FS2_STAT_COPIES 9    /* maps copied by that resample (offspring beyond the first of each survivor)        */
#define FS2_STAT_ANOMALY 10  /* weights the exact scan does not accept (negative / NaN): literal serial loop taken */
#define FS2_STAT_STUCK 11    /* a slot beyond the running total (quirk Q10)                                        */
#define FS2_STAT_ARGMAX_ID 12 /* fs2_normalize / fs2_estimate: LOGICAL (global) index of the arg-max particle       */
#define FS2_STAT_DEFERRED 13  /* fs2_finish_step: of those copies, the ones the next update kernel writes (fs2_sync_maps)   */
#define FS2_STATS_LEN 16

/* result of a whole step, host side */
typedef struct fs2_step_result {
    double x, y, yaw;  /* fast_slam_2.py:67 return value                         */
    double neff;       /* fast_slam_2.py:59                                      */
    double total;      /* weight total before normalisation                      */
    int32_t resampled; /* 1 if the step took fast_slam_2.py:62-64                */
    int32_t status_or; /* reserved                                               */
} fs2_step_result;

int fs2_abi_version(void);
const char *fs2_strerror(int status);
const char *fs2_last_cuda_error(void);

/* FastSLAM2.__init__ (fast_slam_2.py:20-31): allocates the store on cfg->device and resets it. Synchronous. */
int fs2_create(const fs2_config *cfg, fs2_handle *out);
int fs2_destroy(fs2_handle h);
/* particles at (0,0,0), weight 1/global_particles, empty maps (fast_slam_2.py:25-31, particle.py:19-20) */
int fs2_reset(fs2_handle h, void *stream);
int fs2_get_ptrs(fs2_handle h, fs2_ptrs *out);

/*
 * One Gaussian draw per particle, N(0, sigma^2), counter-based (Philox4x32-10 keyed by cfg.seed, counter =
 * (global particle index, step)), so every GPU count sees the same draws.  Stands in for the
 * np.random.normal call at fast_slam_2.py:79/81; tests read the buffer back and hand the same numbers to
 * the oracle (quirk Q15).  noise_dev: double[P].
 */
int fs2_draw_noise(fs2_handle h, double sigma, uint64_t step, double *noise_dev, void *stream);

/* __move_particle for every particle (fast_slam_2.py:69-87).  noise_dev: double[P], already scaled. */
int fs2_motion(fs2_handle h, double rotation, double translation, const double *noise_dev, void *stream);

/*
 * __update_particle for every particle and every measurement, in measurement order per particle
 * (fast_slam_2.py:48-53, 89-159; LandmarkUtils.associate_landmarks landmark_utils.py:92-117;
 * GeometryUtils.mahalanobis_distance geometry_utils.py:14-23).
 * obs_host: double[M][2] = (distance, yaw) per measurement (measurement.py:9-16), HOST memory, M >= 0.
 * assoc_dev: optional int32[M][P]: index matched by each (measurement, particle), -1 = new landmark,
 * -2 = update skipped.
 */
int fs2_update(fs2_handle h, const double *obs_host, int32_t M, int32_t *assoc_dev, void *stream);

/* fs2_motion followed by fs2_update in one launch (pose never leaves registers in between). */
int fs2_motion_update(fs2_handle h, double rotation, double translation, const double *noise_dev,
                      const double *obs_host, int32_t M, int32_t *assoc_dev, void *stream);

/*
 * __normalize_weights + __calculate_effective_particles + __estimate_robot_position
 * (fast_slam_2.py:161-175, 212-223, 201-210) in two passes over w.
 *   fs2_weight_total : stats[FS2_STAT_TOTAL] = sum of this shard's weights.
 *   fs2_normalize    : applies the reference's rule with *total_dev (on 1 GPU: &stats[FS2_STAT_TOTAL];
 *                      with several shards: the all-gathered total), then fills the rest of stats.
 */
int fs2_weight_total(fs2_handle h, void *stream);
int fs2_normalize(fs2_handle h, const double *total_dev, void *stream);
/* stats only (sum w^2, Neff, first arg-max and its pose) without touching the weights: the estimate
 * after a resample (fast_slam_2.py:67, quirk Q11) */
int fs2_estimate(fs2_handle h, void *stream);

/*
 * __low_variance_resample, index part (fast_slam_2.py:183-196): for every destination slot m in
 * [m_begin, m_begin + m_count) the ancestor index k(m) = min{k : c_k >= u0 + m/n} clamped to n-1,
 * where c_k is the SEQUENTIAL left-to-right fp64 running sum of w_all_dev[0..n) -- reproduced bit for
 * bit by a parallel scan that emulates the sequential rounding.  ancestor_dev: int32[m_count].
 * On 1 GPU: w_all_dev = ptrs.w, n = P, m_begin = 0, m_count = P.
 */
int fs2_resample_indices(fs2_handle h, const double *w_all_dev, int64_t n, double u0, int64_t m_begin,
                         int64_t m_count, int32_t *ancestor_dev, void *stream);

/*
 * __low_variance_resample, copy part (deepcopy of the survivors incl. weight and map, fast_slam_2.py:196-199):
 * new slot m := old particle ancestor_dev[m] (LOCAL indices), through a second buffer, then the buffers swap.
 */
int fs2_gather(fs2_handle h, const int32_t *ancestor_dev, void *stream);

/*
 * The same when particles are sharded over several GPUs: ancestor_dev[m] < P names a local
 * particle, ancestor_dev[m] >= P names staged particle (ancestor - P) received from another GPU.
 * Particles travel as records of (8 + 6 * Lcap) doubles built by fs2_pack_records on the sending GPU:
 *   { x, y, yaw, w, count, 0, 0, 0, map rows (count * 6 doubles) ... }
 * records_dev: the n_staged received records, in staging order.
 */
int fs2_gather_ext(fs2_handle h, const int32_t *ancestor_dev, const double *records_dev, int64_t n_staged, void *stream);
/*
 * Peer-memory form of the same migration for the GPUs of one node: every shard's store is mapped into every
 * process with CUDA IPC, and the DESTINATION pulls the particles it needs straight out of the owning GPU's store
 * over NVLink (no packing on the source, no all_to_all).
 *   fs2_ipc_export      writes 7 cudaIpcMemHandle_t (448 bytes) for this shard's x, y, yaw, w, lm, count, slot
 *   fs2_ipc_open_peers  all_handles = the exports of all `world` ranks in rank order (world <= 16)
 *   fs2_gather_p2p      the copy-on-resample gather with GLOBAL ancestors (ancestors_all_dev: int32[global_particles],
 *                       the ancestor of every global slot): local copies and pulls of remote ancestors' poses and maps
 *                       (peer loads over NVLink) happen in the same kernels.  A particle that survives only on another
 *                       GPU keeps its slot for this round, so copies go to the pool's free slots (num_particles +
 *                       spare_slots of them); returns FS2_ERR_NOMEM, with the store untouched, when those do not
 *                       suffice -- fall back to the record path.  New poses land in a second buffer: every rank must
 *                       finish this call's kernels (stream-ordered barrier, e.g. a tiny NCCL all_reduce) before any runs
 *   fs2_gather_commit   which publishes the new pose / weight / count / slot arrays.
 *   fs2_pull_records    records_dev[r] = record (layout above) of GLOBAL particle global_ids_dev[r] (int64), all owned
 *                       by src_rank, read straight from that GPU.  Every rank must have finished pulling (barrier)
 *                       before any rank runs fs2_gather_ext, which rewrites its store.
 */
int fs2_ipc_export(fs2_handle h, void *handles_out);
int fs2_ipc_open_peers(fs2_handle h, const void *all_handles, int32_t world, int32_t rank);
int fs2_gather_p2p(fs2_handle h, const int32_t *ancestors_all_dev, void *stream);
int fs2_gather_commit(fs2_handle h, void *stream);

/*
 * Decoupled placement for sharded filters (csrc/fs2_place.cuh): the reference's particle ORDER (running sum of
 * __low_variance_resample, first arg-max, Serializer; fast_slam_2.py:183-210) is kept as a logical numbering, while an
 * offspring stays on the GPU its ancestor lives on; only what exceeds a GPU's share migrates, fat lineages first.
 *   fs2_place_enable       logical id of local particle i = global_offset + i to begin with; motion noise and the
 *                          arg-max tie rule follow the logical ids from then on.  Needs fs2_ipc_open_peers.
 *   fs2_place_logical_ids  device pointer to int32[P]: logical id of every local particle (increasing).
 *   fs2_place_resample     anc_all_dev  int32[N]: LOGICAL ancestor of every new logical particle (fs2_resample_indices on
 *                                       the weights in logical order);
 *                          place_dev    int32[N]: rank * P + local index of every current logical particle;
 *                          place_new_dev int32[N]: receives the same for the new particles.
 *                          Plans (identically on every rank) and runs this rank's gather -- local copies and NVLink pulls
 *                          out of the other ranks' stores.  info_host[0] = offspring that changed GPU (whole job),
 *                          info_host[1] = maps this rank pulled.  FS2_ERR_NOMEM: too few spare map slots.
 *   fs2_place_commit       after a barrier over all ranks: publishes the new poses / weights / ids.
 */
int fs2_place_enable(fs2_handle h, void *stream);
void *fs2_place_logical_ids(fs2_handle h);
int fs2_place_resample(fs2_handle h, const int32_t *anc_all_dev, const int32_t *place_dev, int32_t *place_new_dev,
                       int64_t *info_host, void *stream);
int fs2_place_commit(fs2_handle h, void *stream);
int fs2_pull_records(fs2_handle h, int32_t src_rank, const int64_t *global_ids_dev, int64_t n, double *records_dev, void *stream);

/* records_dev[r] = record of LOCAL particle sel_dev[r] (int64), r < nsel */
int fs2_pack_records(fs2_handle h, const int64_t *sel_dev, int64_t nsel, double *records_dev, void *stream);

/*
 * FastSLAM2.iterate (fast_slam_2.py:33-67) on one GPU, everything from HOST arguments: draws the motion
 * noise on the device (step index -> counter) unless noise_host is given, runs motion + update,
 * normalises, and resamples when Neff < P/2 with the start point u0 (fast_slam_2.py:183; drawn by the
 * caller so that the random stream stays the caller's).  Synchronous: returns after the estimate is on
 * the host.  assoc_dev / ancestor_dev optional device outputs as above.
 */
/*
 * The weight half of FastSLAM2.iterate (fast_slam_2.py:56-67) as one asynchronous chain: __normalize_weights,
 * __calculate_effective_particles, the decision "Neff < N/2" TAKEN ON THE DEVICE, __low_variance_resample (exact scan,
 * search, copy-on-resample gather) when it says so, and __estimate_robot_position.  Nothing is read back: the caller
 * copies the stats block (FS2_STAT_*: total, Neff before the resample, resampled flag, copies, estimate) when it
 * wants it.  ancestor_dev (optional, int32[P]) receives the resampling indices of a resampling step.
 * Single-shard filters only (sharded filters are driven stage-wise).
 */
int fs2_finish_step(fs2_handle h, double u0, int32_t *ancestor_dev, void *stream);

/*
 * Deferred map copies.  The deep copies of fast_slam_2.py:192-196 (copy.deepcopy of every resampled particle) that
 * fs2_finish_step / fs2_step_host owe for the extra offspring of a resample are written by the NEXT update kernel while
 * it streams the ancestor's map anyway (csrc/fs2_update_ws.cuh, DEFER).  Every entry point of this library that reads or
 * moves maps makes outstanding copies first; a caller that reads the map storage through fs2_get_ptrs on its own calls
 * this before.  No-op when nothing is outstanding.  FS2_DEFER=0 in the environment disables the deferral.
 */
int fs2_sync_maps(fs2_handle h, void *stream);

int fs2_step_host(fs2_handle h, double rotation, double translation, const double *obs_host, int32_t M,
                  const double *noise_host, uint64_t step, double u0, int32_t *assoc_dev,
                  int32_t *ancestor_dev, fs2_step_result *out, void *stream);

/*
 * LandmarkUtils.get_measurements_to_landmarks (fast_slam_2/utils/landmark_utils.py:20-89) for a batch of B scans
 * of N points each: LineFilter.filter (algorithms/line_filter.py:12-21, gaussian sigma), Hough image + cv2.HoughLines
 * (algorithms/hough_transformation.py:14-145), GeometryUtils.cluster_points (utils/geometry_utils.py:26-62, DBSCAN
 * eps 0.5 / min_samples 1), the corner test (landmark_utils.py:66-89) and calculate_distance_and_angle
 * (geometry_utils.py:65-74).  scans_host: double[B][N][2] (x, y) in the robot frame, HOST memory.
 * meas_host: double[B][fs2_frontend_max_measurements()][2] = (distance, yaw); k_host[b] = measurements of scan b;
 * status_host (optional): per-scan bits -- 1: more than 128 Hough peaks, 2: more than 2048 intersections, 4: more than
 * 64 clusters (the surplus is dropped), 8: empty scan, 16: a beam or point that is not finite, 32: Hough image of more
 * than 2^28 pixels (the call then returns FS2_ERR_INVALID).  Synchronous; pinned host memory makes the copies faster.
 * The Hough accumulator lives in shared memory (csrc/fs2_frontend.cuh: fe_raster_list, fe_vote_peaks); FS2_FE_LEGACY=1
 * in the environment selects the global-accumulator kernels, which scans of more than 2016 points use anyway.
 */
int fs2_frontend_max_measurements(void);
int fs2_frontend(const double *scans_host, int32_t B, int32_t N, double sigma, int32_t device, double *meas_host,
                 int32_t *k_host, int32_t *status_host, void *stream);

/*
 * The same from raw laser data: Robot.scan_environment (fast_slam_2/models/robot.py:32-58) fused in front.
 * ranges_host: double[B][N] beam ranges; angles_host: double[N] beam angles in radians (the reference's 180-beam
 * laser: radians(i - 90), robot.py:52).  A beam with range < min_range or > max_range is dropped (robot.py:48),
 * the others become (range cos angle, range sin angle) in beam order, so scans of a batch may keep different
 * numbers of points.  status bit 8: no beam of the scan was in range (k = 0; the reference raises there).
 */
/* HoughTransformation.detect_line_intersections (fast_slam_2/algorithms/hough_transformation.py:14-41) for B sets of
 * N (already filtered) points: the intersections of the detected lines that are at least 45 degrees apart, in metres,
 * in the reference's order.  inter_host: float[B][fs2_hough_max_intersections()][2]; n_inter_host[b] = how many. */
int fs2_hough_max_intersections(void);
/* The front-end decides "distance <= eps" (DBSCAN eps 0.5 of landmark_utils.py:57, corner threshold 0.1 of :63, both
 * np.sqrt(dx**2 + dy**2) <= eps in the reference) as "squared distance <= T(eps)"; this returns T(eps), the largest double
 * whose correctly rounded square root is <= eps, so that the CPU suite can pin the equivalence (no device needed). */
double fs2_frontend_sq_threshold(double eps);
int fs2_hough_intersections(const double *points_host, int32_t B, int32_t N, int32_t device, float *inter_host,
                            int32_t *n_inter_host, int32_t *status_host, void *stream);

/* LineFilter.filter (fast_slam_2/algorithms/line_filter.py:12-21) alone: scipy gaussian_filter1d (mode "reflect",
 * radius int(4 sigma + 0.5)) along the point index of each scan, x and y separately; filtered_host: double[B][N][2] */
int fs2_line_filter(const double *scans_host, int32_t B, int32_t N, double sigma, int32_t device,
                    double *filtered_host, void *stream);

/* the front-end keeps its device scratch (pixel lists, ~110 KB per scan; Hough accumulators on the FS2_FE_LEGACY path,
 * ~2.6 MB per scan) between calls; this frees it */
int fs2_frontend_release(int32_t device);

int fs2_frontend_polar(const double *ranges_host, const double *angles_host, int32_t B, int32_t N, double min_range,
                       double max_range, double sigma, int32_t device, double *meas_host, int32_t *k_host,
                       int32_t *status_host, void *stream);

/*
 * Map clustering: LandmarkUtils.update_known_landmarks (fast_slam_2/utils/landmark_utils.py:120-144) on the
 * filter's maps as they sit in device memory.  Every landmark mean of every particle is a point, in particle
 * order; min_samples = int(min_samples_frac * points / particles) (landmark_utils.py:129-130, 0.7) unless a
 * positive `min_samples` is given (a shard of a larger filter passes the global value); the points go through
 * GeometryUtils.cluster_points (utils/geometry_utils.py:26-62): sklearn DBSCAN(eps, min_samples), one centroid
 * per cluster in label order.  centroids_host: double[max_clusters][2]; members_host (optional): points per
 * cluster.  *n_clusters = clusters found, or -1 when the reference returns early (min_samples < 1,
 * landmark_utils.py:133-134).  FS2_ERR_NOMEM: more clusters than max_clusters, or the point-level part needs more
 * room than the workspace has (fs2_last_cuda_error says what; environment FS2_KL_POINTS / FS2_KL_CLUSTERS
 * size it at first use; the tile grid starts at 2048 tiles and grows by itself (kept under a quarter full) up to 65536 unless FS2_KL_TILES pins
 * it).  FS2_ERR_UNSUPPORTED: the exact point-level part would need more distance tests than its budget (FS2_KL_WORK,
 * default 3e10) -- dense cells about eps apart; the call stops instead of running for minutes.  Synchronous.
 */
typedef struct fs2_kl_info {
    int64_t n_points;        /* points clustered                                              */
    int64_t min_samples;     /* the value used                                                */
    int64_t involved_points; /* points that went through the exact point-level path           */
    int64_t noise_points;    /* label -1                                                      */
    int32_t tiles;           /* tile capacity of the grid                                     */
    int32_t clusters;
    int32_t err_bits;        /* 1 tiles full, 2 non-finite input, 4 out of range, 8 cell count, 16 clusters, 32 work budget */
    int32_t skipped;         /* 1 = min_samples < 1, nothing done                             */
    int32_t tiles_used;      /* tiles (eps x eps) that hold points                            */
    int32_t reserved;
} fs2_kl_info;
int fs2_known_landmarks(fs2_handle h, double eps, double min_samples_frac, int64_t min_samples, int32_t max_clusters,
                        double *centroids_host, int64_t *members_host, int32_t *n_clusters, fs2_kl_info *info,
                        void *stream);

/*
 * The same clustering over a filter that is sharded across GPUs (one handle per rank): one call per stage, the
 * caller does the exchanges in between (fast_slam_b200/dist.py: ShardedFilter.known_landmarks).  The result is the
 * one fs2_known_landmarks gives for the unsharded filter, identical on every rank.
 *   begin   : point indices of the local maps; *n_local_points = landmarks on this shard.
 *             [all-gather n_local -> index_offset of each rank, global N, min_samples]
 *   count   : pass 1 over the local maps (points numbered from index_offset); *n_tiles = occupied tiles.
 *   export  : the occupied tiles as records of fs2_kl_record_bytes() bytes each into records_dev.
 *             [all-gather the records]
 *   merge   : every rank merges ALL records (its own included) and runs the cell-level part;
 *             *involved_points = points (of all ranks) the exact point-level part needs.
 *   extract : the local ones among them as double[n][3] = (x, y, point index); *n_points may exceed cap, then
 *             nothing past cap was written: call again with a larger buffer.   [all-gather the points]
 *   finish  : point-level part on the gathered points (n_points must equal involved_points) and the centroids.
 */
int fs2_kl_record_bytes(void);
int fs2_kl_shard_begin(fs2_handle h, double eps, int64_t *n_local_points, void *stream);
/* placed shards (fs2_place_enable): per-particle global point offsets in LOGICAL particle order, int64[P] on the device;
 * call between fs2_kl_shard_begin and fs2_kl_shard_count(h, 0, ...) */
int fs2_kl_shard_set_bases(fs2_handle h, const int64_t *bases_dev, void *stream);
int fs2_kl_shard_count(fs2_handle h, int64_t index_offset, int32_t *n_tiles, void *stream);
int fs2_kl_shard_export(fs2_handle h, void *records_dev, int32_t cap_records, void *stream);
int fs2_kl_shard_merge(fs2_handle h, const void *records_dev, int32_t n_records, int64_t min_samples,
                       int64_t *involved_points, void *stream);
int fs2_kl_shard_extract(fs2_handle h, double *points_dev, int64_t cap, int64_t *n_points, void *stream);
int fs2_kl_shard_finish(fs2_handle h, const double *points_dev, int64_t n_points, int64_t n_total_points,
                        int32_t max_clusters, double *centroids_host, int64_t *members_host, int32_t *n_clusters,
                        fs2_kl_info *info, void *stream);

/* GeometryUtils.cluster_points (utils/geometry_utils.py:26-62) for a host array double[n][2] */
int fs2_cluster_points(const double *xy_host, int64_t n, double eps, int64_t min_samples, int32_t device,
                       int32_t max_clusters, double *centroids_host, int64_t *members_host, int32_t *n_clusters,
                       fs2_kl_info *info);

/*
 * ICP.get_transformation (fast_slam_2/algorithms/icp.py:13-58) for B pairs of point sets: source_host
 * double[B][n_source][2], target_host double[B][n_target][2] (at most 4096 points each).  Per pair the rotation
 * matrix (row-major double[4]) and translation (double[2]) that align source to target, and the iterations run
 * (optional).  The reference's defaults: max_iterations 100, threshold 1e-5.  Synchronous.
 */
int fs2_icp(const double *source_host, const double *target_host, int32_t B, int32_t n_source, int32_t n_target,
            int32_t max_iterations, double threshold, int32_t device, double *rotation_host, double *translation_host,
            int32_t *iterations_host, void *stream);

/* host-only debugging aid: the per-step observation block (robot-frame Cartesian + screen cell tables) as the
 * update kernel receives it; layout = struct Fs2ObsBatch of fast_slam_b200/csrc/fs2_update.cuh */
int fs2_debug_obs_batch_size(void);
int fs2_debug_obs_batch(const double *obs_host, int32_t M, void *out);

/* launches issued by this handle's entry points since creation (bench.py's gpu_launches) */
int64_t fs2_launch_count(fs2_handle h);

/* host <-> device state exchange for tests and for the `.particles` view of the Python layer */
int fs2_upload_state(fs2_handle h, const double *x, const double *y, const double *yaw, const double *w,
                     const int32_t *count, const double *lm /* [P][Lcap][6] */, void *stream);
int fs2_download_state(fs2_handle h, double *x, double *y, double *yaw, double *w, int32_t *count,
                       double *lm, int32_t *status, void *stream);

/* the same for nsel selected particles (LOCAL indices sel_host[nsel]); lm: [nsel][Lcap][6] */
int fs2_download_particles(fs2_handle h, const int64_t *sel_host, int64_t nsel, double *x, double *y, double *yaw,
                           double *w, int32_t *count, double *lm, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FS2_H */
