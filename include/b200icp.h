/* b200icp.h -- C ABI of the B200-native ICP scan-registration path.
 *
 * This is the drop-in boundary for the ICP seam that mola::LidarOdometry
 * drives (reference: src/LidarOdometry.cpp).  Each entry point names the
 * reference interface it replaces.  Plain pointers and sizes only; all device
 * memory, streams and kernels (sm_100a) live behind the opaque handles.
 * There is no CPU fallback: every call fails with B200ICP_ERR_CUDA when no
 * CUDA device / kernel image is usable.
 *
 * Threading (reference: LidarOdometry.h:167-172, cpp:94-96, 711-729 -- one
 * shared ICP object, align() called concurrently from pool threads):
 *   b200icp_align / _align_batch / _knn / _match / _voxel_decimate are
 *   re-entrant on one b200icp_t (per-call workspace + stream from a pool);
 *   handles are immutable after creation.
 */
#ifndef B200ICP_H
#define B200ICP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200ICP_INVALID_IDX 0xFFFFFFFFu
#define B200ICP_MAX_KNN 8

enum
{
    B200ICP_OK = 0,
    B200ICP_ERR_BAD_ARG = -1,
    B200ICP_ERR_CUDA = -2,
    B200ICP_ERR_YAML = -3,
    B200ICP_ERR_UNSUPPORTED = -4,
    B200ICP_ERR_NOMEM = -5
};

/* mp2p_icp::IterTermReason, logged as an integer at LidarOdometry.cpp:888.
 * Numeric values frozen here. */
enum
{
    B200ICP_TERM_UNDEFINED = 0,
    B200ICP_TERM_NO_PAIRINGS = 1,
    B200ICP_TERM_SOLVER_ERROR = 2,
    B200ICP_TERM_MAX_ITERATIONS = 3,
    B200ICP_TERM_STALLED = 4
};

enum { B200ICP_SOLVER_GAUSS_NEWTON = 0, B200ICP_SOLVER_HORN = 1 };
enum { B200ICP_MATCHER_POINT2PLANE = 0, B200ICP_MATCHER_POINTS_DISTANCE = 1 };

/* What load_icp_set_of_params() builds from one ICP YAML block
 * (LidarOdometry.cpp:57-88): mp2p_icp::Parameters::load_from (cpp:78),
 * initialize_solvers (cpp:81), initialize_matchers (cpp:84),
 * initialize_quality_evaluators (cpp:87).  Keys: params/icp-settings-regular.yaml:7-46. */
typedef struct b200icp_params
{
    uint32_t max_iterations;     /* params.maxIterations */
    double   min_abs_step_trans; /* params.minAbsStep_trans */
    double   min_abs_step_rot;   /* params.minAbsStep_rot */
    int32_t  use_scale_outlier_detector;
    double   scale_outlier_threshold;
    int32_t  use_robust_kernel;
    double   robust_kernel_param; /* radians here, degrees in the YAML */
    double   robust_kernel_scale;
    int32_t  solver_kind;           /* solvers[0].class */
    uint32_t solver_max_iterations; /* solvers[0].params.maxIterations */
    double   gn_min_delta;          /* optional solvers[0].params.minDelta */
    int32_t  matcher_kind;          /* matchers[0].class */
    double   distance_threshold;
    double   plane_eigen_threshold;
    uint32_t knn;
    uint32_t min_plane_points; /* optional matchers[0].params.minimumPlanePoints */
    uint32_t run_from_iteration;
    uint32_t run_up_to_iteration;
    double   quality_threshold_distance; /* quality[0].params.thresholdDistance */
    double   cov_fd_step;
} b200icp_params_t;

/* mp2p_icp::Parameters, the part of the block above that the reference hands to EVERY align() call next to the
 * shared ICP object (in.icp_params, LidarOdometry.cpp:869-871; chosen per scan at cpp:287-290): the iteration
 * budget, the step tolerances and pairingsWeightParameters.  Solvers, matchers and quality evaluators belong to
 * the object (initialize_solvers / _matchers / _quality_evaluators, cpp:80-87) and are not part of a call. */
typedef struct b200icp_call_params
{
    uint32_t max_iterations;
    double   min_abs_step_trans;
    double   min_abs_step_rot;
    int32_t  use_scale_outlier_detector;
    double   scale_outlier_threshold;
    int32_t  use_robust_kernel;
    double   robust_kernel_param;
    double   robust_kernel_scale;
} b200icp_call_params_t;

/* mp2p_icp::Results as consumed at LidarOdometry.cpp:873-888:
 * optimal_tf {mean, cov}, quality, nIterations, terminationReason. */
typedef struct b200icp_result
{
    double   pose[6];  /* x y z yaw pitch roll (mrpt::math::TPose3D order) */
    double   R[9];     /* row-major rotation of the same pose */
    double   t[3];
    double   cov[36];  /* row-major, order x y z yaw pitch roll */
    double   quality;  /* [0,1] */
    uint32_t n_iterations;
    uint32_t termination_reason;
    uint32_t n_pairings; /* of the last matcher run */
    uint32_t cov_singular;
} b200icp_result_t;

/* per-kernel device timings, CUDA events on the launching stream */
typedef struct b200icp_profile
{
    uint64_t match_launches;  /* the matcher's search: transform + exact kNN (per-lane shell walk + cooperative pass) */
    double   match_ms;
    uint64_t match_queries;   /* queries processed by those launches */
    uint64_t solve_launches;
    double   solve_ms;
    uint64_t index_builds;
    double   index_ms;
    uint64_t index_points;
    uint64_t knn_launches;
    double   knn_ms;
    uint64_t knn_queries;
    uint64_t voxel_launches;
    double   voxel_ms;
    uint64_t voxel_points;
    uint64_t total_kernel_launches; /* every kernel this library launched */
    uint64_t fit_launches;    /* plane fit + gates + moments kernel */
    double   fit_ms;
    uint64_t graph_replays;   /* single registrations whose first batch ran as one CUDA-graph launch */
} b200icp_profile_t;

typedef struct b200icp       b200icp_t;       /* the mp2p_icp::ICP object + its Parameters */
typedef struct b200icp_cloud b200icp_cloud_t; /* one point layer of a metric_map_t, resident in HBM */

/* thread-local message of the last failing call on this thread */
const char* b200icp_last_error(void);
/* number of usable CUDA devices (0 => every other call fails) */
int b200icp_device_count(void);

/* --- ICP object (LidarOdometry.cpp:57-88) -------------------------------- */
void b200icp_default_params(b200icp_params_t* p);
/* Parses the text of one ICP settings block (icp_class / params / solvers /
 * matchers / quality). Unknown icp_class / class names fail like cpp:70-75. */
int b200icp_params_from_yaml(const char* yaml_text, b200icp_params_t* out);
int b200icp_create(const b200icp_params_t* params, int device, b200icp_t** out);
int b200icp_create_from_yaml(const char* yaml_text, int device, b200icp_t** out);
void b200icp_destroy(b200icp_t* icp);
int b200icp_get_params(const b200icp_t* icp, b200icp_params_t* out);
int b200icp_device(const b200icp_t* icp);

/* --- clouds (mp2p_icp::metric_map_t point layer; MRPT CPointsMap SoA) ---- */
/* Copies n points (host SoA float, any memory; pinned gives async DMA) to
 * the device and builds the search index there (replaces the lazy nanoflann
 * kd-tree build).  search_radius <= 0 => the ICP object's matcher threshold. */
int b200icp_cloud_upload(b200icp_t* icp, const float* x, const float* y, const float* z,
                         size_t n, float search_radius, b200icp_cloud_t** out);
/* Coordinates only, no search index: the input of a filter stage (b200icp_voxel_decimate) whose OUTPUT is what
 * gets registered -- apply_generators followed by apply_filter_pipeline, LidarOdometry.cpp:215-224.  Such a cloud
 * is refused (B200ICP_ERR_BAD_ARG) as a side of b200icp_knn / _match / _align. */
int b200icp_cloud_upload_raw(b200icp_t* icp, const float* x, const float* y, const float* z, size_t n,
                             b200icp_cloud_t** out);
/* Same from DEVICE pointers (inputs already resident in HBM). */
int b200icp_cloud_from_device(b200icp_t* icp, const float* dx, const float* dy, const float* dz,
                              size_t n, float search_radius, b200icp_cloud_t** out);
void   b200icp_cloud_free(b200icp_cloud_t* c);
size_t b200icp_cloud_size(const b200icp_cloud_t* c);
/* bytes of HBM the cloud and its index occupy (key-frame store accounting) */
size_t b200icp_cloud_device_bytes(const b200icp_cloud_t* c);
/* original-order coordinates back to the host */
int b200icp_cloud_download(const b200icp_cloud_t* c, float* x, float* y, float* z);

/* --- voxel decimation (apply_filter_pipeline, LidarOdometry.cpp:223-224) - */
/* One point per occupied voxel: the lowest original index (or the voxel mean
 * with use_average), output ordered by ascending original index.  The result
 * is a new indexed cloud; keep_idx (optional, capacity n) receives the kept
 * original indices. */
int b200icp_voxel_decimate(b200icp_t* icp, const b200icp_cloud_t* in, float resolution,
                           int use_average, float search_radius, b200icp_cloud_t** out,
                           uint32_t* keep_idx);

/* FilterEdgesPlanes on the device: the filter class the reference's parameter
 * files name (params/kitti-default.yaml:21-32 `pointcloud_filter_class:
 * mola::lidar_segmentation::FilterEdgesPlanes` + `pointcloud_filter_params`;
 * defaults include/mola-fe-lidar/LidarOdometry.h:76-80), applied where the
 * reference calls apply_filter_pipeline (LidarOdometry.cpp:223-224).  Per voxel
 * of voxel_filter_resolution with at least min_points_per_voxel points: eigenvalues
 * e0 <= e1 <= e2 of the covariance; e2 < max_e2_e0*e0 && e1 < max_e1_e0*e0 ->
 * "edges"; else e2 > min_e2_e0*e0 && e1 > min_e1_e0*e0 and a normal that is not
 * vertical (|n.z| < 0.9) -> "planes"; every voxel_filter_decimation-th point of
 * a classified voxel joins its layer, every full_pointcloud_decimation-th point
 * of every voxel joins "full_decim" (points in ascending original index). */
typedef struct b200icp_edges_planes_params
{
    float    voxel_filter_resolution;    /* [m] (yaml:25) */
    uint32_t full_pointcloud_decimation; /* yaml:27 */
    uint32_t voxel_filter_decimation;    /* yaml:28 */
    float    voxel_filter_max_e2_e0, voxel_filter_max_e1_e0; /* yaml:29-30 */
    float    voxel_filter_min_e2_e0, voxel_filter_min_e1_e0; /* yaml:31-32 */
    uint32_t min_points_per_voxel;       /* additive; 5 */
} b200icp_edges_planes_params_t;
/* The shipped values (kitti-default.yaml:23-32). */
void b200icp_edges_planes_defaults(b200icp_edges_planes_params_t* p);
/* layers_out[0..2] = new indexed clouds "edges", "planes", "full_decim" (each
 * in ascending original index; free with b200icp_cloud_free).  layer_flags_out
 * (optional, host, capacity = size of `in`): bit 0 edges, bit 1 planes, bit 2
 * full_decim per input point.  n_classified_voxels_out: optional. */
int b200icp_filter_edges_planes(b200icp_t* icp, const b200icp_cloud_t* in,
                                const b200icp_edges_planes_params_t* params, float search_radius,
                                b200icp_cloud_t* layers_out[3], uint8_t* layer_flags_out,
                                uint32_t* n_classified_voxels_out);

/* --- nearest neighbours (kdTreeNClosestPoint3DIdx behind the matchers) --- */
/* For each point of `queries` moved by pose (x,y,z,yaw,pitch,roll; NULL =
 * identity): the k nearest points of `ref` with d2 <= max_dist^2, ascending by
 * (d2 as float32, index).  Outputs are host arrays [nq*k] in the queries'
 * ORIGINAL order, padded with B200ICP_INVALID_IDX / +inf.  max_dist must be
 * positive; +infinity = uncapped search: the capped pass at the radius `ref` was
 * indexed for, then the rows still short of k neighbours are completed exactly
 * over the whole cloud (one warp per query). */
int b200icp_knn(b200icp_t* icp, const b200icp_cloud_t* ref, const b200icp_cloud_t* queries,
                const double* pose6, uint32_t k, float max_dist, uint32_t* idx_out,
                float* d2_out);

/* --- sharded maps: per-GPU partial arg-min, merged over NVLink ------------ */
/* A packed key is (d2 as float32 bits) << 32 | index: its unsigned integer
 * order IS the (d2, index) order of the tie rule, so partial results of
 * disjoint map shards merge with a plain minimum (k = 1: an all-reduce MIN on
 * 64-bit integers; k > 1: b200icp_merge_keys_device after an exchange). */
#define B200ICP_NO_KEY 0x7F800000FFFFFFFFull /* +inf distance, invalid index */
/* Same search as b200icp_knn, results left ON THE DEVICE as packed keys
 * d_keys_out[nq*k] (queries' original order, padded with B200ICP_NO_KEY).
 * d_index_map (device, optional, [size of ref]): shard-local index -> global
 * index of the caller's unsharded map; must be increasing so that ties order
 * alike in both numberings.  Returns after the kernels have completed. */
int b200icp_knn_keys_device(b200icp_t* icp, const b200icp_cloud_t* ref, const b200icp_cloud_t* queries,
                            const double* pose6, uint32_t k, float max_dist,
                            const uint32_t* d_index_map, uint64_t* d_keys_out);
/* Merges `parts` ascending key lists per query: d_parts[p*part_stride + q*k + i]
 * -> d_out[q*k + i], the k smallest of the union, ascending (device pointers). */
int b200icp_merge_keys_device(b200icp_t* icp, const uint64_t* d_parts, uint32_t parts,
                              size_t part_stride, size_t nq, uint32_t k, uint64_t* d_out);

/* Fused search + exchange over peer memory (NVLink): the search kernel itself
 * stores every result row into the gather buffer of EVERY rank at
 * [rank][query][k] (P2P stores, overlapping the search), or -- k = 1 with
 * `atomic_min` -- folds its key into every rank's [query] slot with a
 * system-scope atomicMin, so that after a barrier each rank holds the merged
 * arg-min without any collective data movement.  d_gather[r] is rank r's
 * buffer as mapped into THIS process (own buffer: the local pointer; peers:
 * b200icp_peer_open).  Buffers must hold world*nq*k keys (nq keys with
 * atomic_min, pre-filled with B200ICP_NO_KEY by their owner).  The caller
 * orders the phases across ranks (barrier before: buffers ready; barrier
 * after: all stores have landed). */
int b200icp_knn_keys_scatter(b200icp_t* icp, const b200icp_cloud_t* ref, const b200icp_cloud_t* queries,
                             const double* pose6, uint32_t k, float max_dist,
                             const uint32_t* d_index_map, uint64_t* const* d_gather, uint32_t world,
                             uint32_t rank, int atomic_min);
/* Exchange buffers: zero-initialised device memory that other processes on the
 * same node can map (CUDA IPC).  handle_out: 64 opaque bytes to hand to the peers. */
int b200icp_peer_alloc(b200icp_t* icp, size_t bytes, void** d_ptr, unsigned char handle_out[64]);
int b200icp_peer_free(b200icp_t* icp, void* d_ptr);
int b200icp_peer_open(b200icp_t* icp, const unsigned char handle[64], void** d_peer_ptr);
int b200icp_peer_close(b200icp_t* icp, void* d_peer_ptr);
/* Barrier across the ranks of one node through peer memory, no collective
 * library: d_flags[r] = rank r's flag array (>= 8 uint64, zero-initialised, in
 * an exchange buffer), `epoch` a number that grows by one per barrier.  A tiny
 * kernel stores `epoch` into slot [rank] of every rank's array (system-scope
 * release, after everything this handle's stream has written) and waits until
 * its own array shows `epoch` in all slots.  Returns B200ICP_ERR_CUDA if a peer
 * did not arrive within ~2 s (never spins forever). */
int b200icp_peer_barrier(b200icp_t* icp, uint64_t* const* d_flags, uint32_t world, uint32_t rank,
                         uint64_t epoch);
/* The whole sharded query in ONE call on one stream with one host
 * synchronisation, result in d_out[nq*k] (device).
 *   k = 1: reset, barrier, search folding its key into the slot of query q in
 *          the buffer of the rank that OWNS q (atomicMin_system over NVLink),
 *          barrier, the owner stores its slice into every rank's result region,
 *          barrier.
 *   k > 1: barrier, search storing the row of query q into the buffer of the
 *          rank that OWNS q (q / ceil(nq/world)), barrier, the owner merges its
 *          slice and stores the merged rows into every rank's result region,
 *          barrier (reduce-scatter + all-gather, both as P2P stores from the
 *          kernels: 2*nq*k keys of traffic per rank).
 * d_bases[r]: rank r's exchange buffer as mapped here (b200icp_peer_alloc /
 * _open), at least 256 + 2*world*ceil(nq/world)*k*8 bytes: barrier flags in
 * the first 256 bytes, keys after them.  *epoch_io: the barrier epoch, same
 * start value (0) on every rank, advanced by the call. */
int b200icp_knn_keys_exchange(b200icp_t* icp, const b200icp_cloud_t* ref, const b200icp_cloud_t* queries,
                              const double* pose6, uint32_t k, float max_dist,
                              const uint32_t* d_index_map, uint64_t* const* d_bases, uint32_t world,
                              uint32_t rank, uint64_t* epoch_io, uint64_t* d_out);
/* fills n keys at a device pointer with B200ICP_NO_KEY (exchange-buffer reset) */
int b200icp_fill_no_key(b200icp_t* icp, uint64_t* d_keys, size_t n);

/* --- a map sharded by spatial cell over the GPUs of one box, natively (NCCL owned by the library) ------------
 * One process per GPU.  Rank 0 makes a communicator id and hands the 128 bytes to the others by any means (a
 * file, MPI, a socket, torch.distributed); every rank then creates its communicator on its own b200icp_t.
 * The reference holds one kd-tree per cloud in host memory and has no counterpart (BASELINE
 * config 5): this is the `from` side of icp->align (LidarOdometry.cpp:869-871) when it is too large for one
 * GPU's search. */
#define B200ICP_COMM_ID_BYTES 128
typedef struct b200icp_comm        b200icp_comm_t;
typedef struct b200icp_sharded_map b200icp_sharded_map_t;
int  b200icp_comm_unique_id(unsigned char id_out[B200ICP_COMM_ID_BYTES]);
int  b200icp_comm_create(b200icp_t* icp, const unsigned char id[B200ICP_COMM_ID_BYTES], int world, int rank,
                         b200icp_comm_t** out);
void b200icp_comm_destroy(b200icp_comm_t* comm);
/* Every rank passes the WHOLE map (host SoA).  The points are dealt to the ranks by coarse (x, y) cell of edge
 * `cell` metres along a Morton curve -- round-robin (`interleaved`: every rank holds 1/world of every
 * neighbourhood) or in `world` runs of equal point count -- the same partition on every rank.  A rank indexes
 * its own cells and keeps a plain copy of all coordinates (16 B per point) for the plane fits. */
int  b200icp_sharded_map_create(b200icp_comm_t* comm, const float* x, const float* y, const float* z, size_t n,
                                float cell, int interleaved, float search_radius, b200icp_sharded_map_t** out);
void b200icp_sharded_map_destroy(b200icp_sharded_map_t* map);
size_t b200icp_sharded_map_local_size(const b200icp_sharded_map_t* map);
/* b200icp_knn_keys_device against the whole map: every rank searches its shard, the partial lists are merged
 * over NVLink (k = 1: ncclAllReduce(MIN) on the packed keys; k > 1: ncclAllGather + k-way merge kernel).  The same
 * [nq*k] keys, GLOBAL indices, on every rank; collective: every rank calls it with the same queries. */
int  b200icp_sharded_knn_keys(b200icp_sharded_map_t* map, const b200icp_cloud_t* queries, const double* pose6,
                              uint32_t k, float max_dist, uint64_t* d_keys_out);
/* b200icp_align_with against the whole map (collective: same local cloud, guess and call parameters on every
 * rank).  Per outer iteration: search on every shard, the lists of a slice of the local points sent to the rank
 * that owns the slice (reduce-scatter), k-way merge, plane fit and moments on the owner, one all-gather of the
 * per-group moment partials, the same fixed-order sum and the same solve on every rank.  The result is
 * BIT-IDENTICAL to b200icp_align against the unsharded map, on every rank.  Gauss-Newton solver only. */
int  b200icp_sharded_align(b200icp_sharded_map_t* map, const b200icp_cloud_t* to_local, const double guess6[6],
                           const b200icp_call_params_t* call, b200icp_result_t* out);

/* --- matcher at a fixed pose (Matcher_Point2Plane; parity hook) ---------- */
/* Host outputs in the local cloud's ORIGINAL order: paired[n] (0/1),
 * nn_idx[n*knn] (after the distance cut, padded INVALID), nn_cnt[n],
 * centroid[n*3], normal[n*3] (f64).  Any output may be NULL.  With
 * Matcher_Points_DistanceThreshold (d2 < threshold^2, strict) nn_idx is [n]
 * (the single neighbour), centroid receives the paired global point and
 * normal stays zero. */
int b200icp_match(b200icp_t* icp, const b200icp_cloud_t* from_global,
                  const b200icp_cloud_t* to_local, const double* pose6, uint8_t* paired,
                  uint32_t* nn_idx, uint32_t* nn_cnt, double* centroid, double* normal,
                  uint32_t* n_pairings);

/* --- registration (mp2p_icp::ICP::align, LidarOdometry.cpp:869-871) ------ */
/* from = global / reference cloud, to = local cloud moved by the pose;
 * guess = init_guess_to_wrt_from.  The whole iteration loop runs on the
 * device; one result struct comes back. */
int b200icp_align(b200icp_t* icp, const b200icp_cloud_t* from_global,
                  const b200icp_cloud_t* to_local, const double guess6[6],
                  b200icp_result_t* out);
/* The same with the call's own mp2p_icp::Parameters (NULL = the object's): what icp->align(from, to, guess,
 * in.icp_params, result) means at cpp:869-871 -- the object's matchers / solvers / quality evaluators, the
 * caller's iteration budget, tolerances and pairing weights.  No device object is created or destroyed. */
void b200icp_call_params_of(const b200icp_params_t* p, b200icp_call_params_t* out);
int  b200icp_align_with(b200icp_t* icp, const b200icp_cloud_t* from_global, const b200icp_cloud_t* to_local,
                        const double guess6[6], const b200icp_call_params_t* call, b200icp_result_t* out);
/* n independent registrations in lock-step launches (worker_pool_past_KFs_
 * jobs cpp:711-729 and the Monte-Carlo loop cpp:775-787). Cloud handles may
 * repeat (their indices are shared). guesses = [n*6]. */
int b200icp_align_batch(b200icp_t* icp, size_t n, const b200icp_cloud_t* const* from_global,
                        const b200icp_cloud_t* const* to_local, const double* guesses6,
                        b200icp_result_t* out);

/* --- measurement --------------------------------------------------------- */
void b200icp_profile_enable(b200icp_t* icp, int enable);
void b200icp_profile_reset(b200icp_t* icp);
void b200icp_profile_get(b200icp_t* icp, b200icp_profile_t* out);
/* raw CUstream (as void*) used by the calling thread's next call, so that a
 * harness can bracket calls with its own events */
int b200icp_synchronize(b200icp_t* icp);

#ifdef __cplusplus
}
#endif
#endif
