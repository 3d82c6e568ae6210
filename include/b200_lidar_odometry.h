/* b200_lidar_odometry.h -- C facade of the host-side mola::LidarOdometry
 * mirror (mola-fe-lidar_b200/host/LidarOdometry.{h,cpp}).
 *
 * The reference module is driven by mola-launcher: initialize(Yaml) once, then
 * onNewObservation(CObservation::Ptr&) from a RawDataSource thread
 * (include/mola-fe-lidar/LidarOdometry.h:38-43; src/LidarOdometry.cpp:90-187).
 * This facade is that same lifecycle for harnesses without the MOLA stack
 * (Python tests, bench.py): plain pointers and sizes only.
 */
#ifndef B200_LIDAR_ODOMETRY_H
#define B200_LIDAR_ODOMETRY_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200lo b200lo_t;

typedef struct b200lo_state
{
    double   last_obs_tim;
    double   accum_since_last_kf[6]; /* x y z yaw pitch roll */
    double   last_twist[6];          /* vx vy vz wx wy wz */
    int32_t  last_iter_twist_is_good;
    uint64_t last_kf;
    uint64_t n_keyframes, n_factors, n_localizations;
    uint64_t n_processed, n_dropped, n_icp;
    double   last_icp_goodness;
    double   last_icp_pose[6];
    uint32_t last_icp_iterations, last_icp_termination;
    size_t   last_points_size;
    uint64_t n_graph_edges, n_checked_pairs;
    uint64_t n_kf_spills, n_kf_reloads; /* key-frame store: clouds spilled to host memory / brought back (process-wide) */
} b200lo_state_t;

typedef struct b200lo_factor
{
    uint64_t from_kf, to_kf;
    double   rel_pose[6];
} b200lo_factor_t;

const char* b200lo_last_error(void);

/* `yaml_path`: a SLAM-system style file whose module block holds `params:`
 * (e.g. params/kitti-default.yaml wrapped by the harness), or NULL with
 * `yaml_text` given.  `mola_dir`: what `$(mola-dir mola-fe-lidar)` resolves to
 * (NULL: the package directory of this library). */
int  b200lo_create(const char* yaml_path, const char* yaml_text, const char* mola_dir, b200lo_t** out);
void b200lo_destroy(b200lo_t* lo);
void b200lo_reset(b200lo_t* lo);

/* LidarOdometry::onNewObservation: enqueues the scan on the 1-thread pool
 * (asynchronous, with the reference's >10-queued drop rule). */
int b200lo_on_new_observation(b200lo_t* lo, const char* sensor_label, double timestamp,
                              const float* x, const float* y, const float* z, size_t n);
/* Same without the copy: the caller keeps x / y / z valid (pinned host memory gives asynchronous DMA) until
 * the scan has been processed (b200lo_wait_idle, or b200lo_queue_length() == 0 and the next call returned).
 * With `b200_prefetch_uploads` (default) the scan is uploaded and indexed on its own stream while the previous
 * one is still being registered. */
int b200lo_enqueue_observation(b200lo_t* lo, const char* sensor_label, double timestamp,
                               const float* x, const float* y, const float* z, size_t n);
/* scans waiting in the module's 1-thread pool: a harness feeding faster than real time throttles on it (the
 * reference drops scans once more than 10 are queued, cpp:171-179) */
size_t b200lo_queue_length(b200lo_t* lo);
/* same, but processes the scan on the calling thread before returning; the
 * coordinates may live in pinned memory and are not copied */
int b200lo_process_observation(b200lo_t* lo, const char* sensor_label, double timestamp,
                               const float* x, const float* y, const float* z, size_t n);
void b200lo_spin_once(b200lo_t* lo);
void b200lo_wait_idle(b200lo_t* lo);
int  b200lo_get_state(b200lo_t* lo, b200lo_state_t* out);
/* factors the back-end received; returns the count (fills up to cap) */
size_t b200lo_get_factors(b200lo_t* lo, b200lo_factor_t* out, size_t cap);
/* The last loop-closure attempt (LidarOdometry.cpp:768-787): the Monte-Carlo guesses drawn (x y z yaw pitch
 * roll each), the goodness each registration reached, the winner.  Returns the number of samples (fills up to
 * cap); 0 when no loop closure has been attempted.  The reference draws from an unseeded generator (cpp:773);
 * here the draws are reproducible (`b200_montecarlo_seed`) and reported, so that a harness can follow the branch. */
size_t b200lo_last_montecarlo(b200lo_t* lo, uint64_t* from_kf, uint64_t* to_kf, double* guesses6, double* goodness,
                              size_t cap, double* best_goodness, double* best_pose6);
/* the scalar front-end parameters after YAML loading, as "key=value\n" text */
size_t b200lo_dump_params(b200lo_t* lo, char* buf, size_t cap);
/* profiler sections (name, count, total seconds, longest call) as "name,count,total,max\n" */
size_t b200lo_dump_profile(b200lo_t* lo, char* buf, size_t cap);
/* raw handle of the ICP object of one AlignKind (0,1,2) for profiling hooks */
void* b200lo_icp_handle(b200lo_t* lo, int align_kind);

#ifdef __cplusplus
}
#endif
#endif
