/* icp_oracle.h -- CPU restatement (plain C99) of the ICP scan-registration hot
 * path that mola::LidarOdometry drives through mp2p_icp::ICP::align
 * (reference call site: src/LidarOdometry.cpp:851-895, 869-871).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or
 * executed by the product library.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may call it, and only as
 * the checker or the timed CPU baseline.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in mp2p_icp /
 * mp2p_icp_filters / MRPT(+nanoflann), none of which is vendored in the
 * reference tree, pinned to a version, or installable here, and the reference
 * holds no tests or golden vectors.  The algorithm
 * bodies below therefore follow the frozen spec in SURVEY.md Appendix A
 * (A.1-A.11), cross-checked in tests/ against independent implementations
 * (scipy cKDTree, numpy eigh / lstsq, analytic known-transform recovery).
 */
#ifndef ICP_ORACLE_H
#define ICP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_INVALID_IDX 0xFFFFFFFFu

/* termination reasons, numeric values frozen in include/b200icp.h too */
enum
{
    ORC_TERM_UNDEFINED      = 0,
    ORC_TERM_NO_PAIRINGS    = 1,
    ORC_TERM_SOLVER_ERROR   = 2,
    ORC_TERM_MAX_ITERATIONS = 3,
    ORC_TERM_STALLED        = 4
};

enum { ORC_SOLVER_GAUSS_NEWTON = 0, ORC_SOLVER_HORN = 1 };
enum { ORC_MATCHER_POINT2PLANE = 0, ORC_MATCHER_POINTS_DISTANCE = 1 };

/* Mirrors the keys of params/icp-settings-regular.yaml:7-46 */
typedef struct
{
    /* params: (yaml:10-13) */
    uint32_t max_iterations;
    double   min_abs_step_trans;
    double   min_abs_step_rot;
    /* params.pairingsWeightParameters (yaml:14-21) */
    int32_t use_scale_outlier_detector;
    double  scale_outlier_threshold;
    int32_t use_robust_kernel;
    double  robust_kernel_param; /* radians (YAML value is in degrees) */
    double  robust_kernel_scale;
    /* solvers[0] (yaml:23-26) */
    int32_t  solver_kind;
    uint32_t solver_max_iterations;
    double   gn_min_delta;
    /* matchers[0] (yaml:32-39) */
    int32_t  matcher_kind;
    double   distance_threshold;
    double   plane_eigen_threshold;
    uint32_t knn;
    uint32_t min_plane_points;
    uint32_t run_from_iteration;
    uint32_t run_up_to_iteration;
    /* quality[0] (yaml:43-46) */
    double quality_threshold_distance;
    /* covariance (Appendix A.9) */
    double cov_fd_step;
} orc_params;

typedef struct
{
    double   pose[6];  /* x y z yaw pitch roll */
    double   R[9];     /* row-major */
    double   t[3];
    double   cov[36];  /* row-major, order x y z yaw pitch roll */
    double   quality;
    uint32_t n_iterations;
    uint32_t termination_reason;
    uint32_t n_pairings; /* pairings of the last matcher run */
    uint32_t cov_singular;
} orc_result;

typedef struct orc_cloud orc_cloud;

void orc_default_params(orc_params* p);

/* clouds (A.1): SoA float, index = insertion order; kd-tree built lazily */
orc_cloud* orc_cloud_create(const float* x, const float* y, const float* z, size_t n);
void       orc_cloud_free(orc_cloud* c);
size_t     orc_cloud_size(const orc_cloud* c);

/* A.3/A.4: exact kNN under the (d2, index) tie rule. max_d2 = +inf for an
 * uncapped search; candidates with d2 > max_d2 are never returned. Outputs
 * are [nq*k], ascending, padded with ORC_INVALID_IDX / +inf. */
void orc_knn_brute(const orc_cloud* ref, const float* qx, const float* qy, const float* qz,
                   size_t nq, uint32_t k, float max_d2, uint32_t* idx_out, float* d2_out);
void orc_knn_kdtree(orc_cloud* ref, const float* qx, const float* qy, const float* qz,
                    size_t nq, uint32_t k, float max_d2, uint32_t* idx_out, float* d2_out);

/* A.2 */
void orc_pose_to_Rt(const double pose[6], double R[9], double t[3]);
void orc_Rt_to_pose(const double R[9], const double t[3], double pose[6]);
void orc_transform_points(const double R[9], const double t[3], const float* x, const float* y,
                          const float* z, size_t n, float* ox, float* oy, float* oz);

/* SE(3) helpers (MRPT order: eps = (v, omega)) */
void orc_se3_exp(const double eps[6], double R[9], double t[3]);
void orc_se3_log(const double R[9], const double t[3], double eps[6]);
void orc_se3_compose(const double Ra[9], const double ta[3], const double Rb[9],
                     const double tb[3], double R[9], double t[3]);
void orc_se3_inverse_compose(const double Ra[9], const double ta[3], const double Rb[9],
                             const double tb[3], double R[9], double t[3]); /* a^-1 * b */

/* row J of SURVEY 8a: symmetric 3x3 eigen (cyclic Jacobi), ascending */
void orc_eig3_sym(const double C[9], double evals[3], double evecs[9] /* columns */);
/* column-pivoting Householder QR solve of A x = b (6x6); returns rank */
int orc_qr_solve6(const double A[36], const double b[6], double x[6]);
/* generic small symmetric inverse through the same QR (n<=6) */
int orc_inverse6(const double A[36], double Ainv[36]);

/* A.5 Point2Plane matcher at a given pose. Per local point i:
 *   paired[i]   0/1
 *   nn_idx[i*k..]  kNN indices after the distance cut (padded INVALID)
 *   nn_cnt[i]   neighbours kept
 *   centroid[i*3..], normal[i*3..]  (valid when paired)
 * Returns the number of pairings. use_kdtree=0 -> brute force search. */
size_t orc_match_point2plane(orc_cloud* global, const orc_cloud* local, const double R[9],
                             const double t[3], const orc_params* p, int use_kdtree,
                             uint8_t* paired, uint32_t* nn_idx, uint32_t* nn_cnt,
                             double* centroid, double* normal);

/* Matcher_Points_DistanceThreshold: 1-NN with d2 < thr2. nn[i] = index or INVALID */
size_t orc_match_points(orc_cloud* global, const orc_cloud* local, const double R[9],
                        const double t[3], double threshold, int use_kdtree, uint32_t* nn,
                        float* nn_d2);

/* A.6 Gauss-Newton on point-to-plane pairings {p_local, c, n}; pose in/out.
 * Returns inner iterations run. */
int orc_gn_point2plane(const double* p_local, const double* c, const double* n, size_t np,
                       uint32_t max_iters, double min_delta, double R[9], double t[3]);
/* GN on point-to-point pairings {p_local, q_global} */
int orc_gn_point2point(const double* p_local, const double* q_global, size_t np,
                       uint32_t max_iters, double min_delta, double R[9], double t[3]);
/* A.10 Horn closed form on point-to-point pairings; returns pairs used */
size_t orc_horn(const double* p_local, const double* q_global, size_t np, const orc_params* p,
                const double Rprior[9], double R[9], double t[3]);

/* A.8 */
double orc_quality_paired_ratio(orc_cloud* global, const orc_cloud* local, const double R[9],
                                const double t[3], double threshold, int use_kdtree);

/* rows G..P: whole registration. from = global / reference cloud (kd-tree
 * side), to = local cloud moved by the pose (LidarOdometry.cpp:869-871). */
int orc_icp_align(orc_cloud* from_global, const orc_cloud* to_local, const double guess[6],
                  const orc_params* p, int use_kdtree, orc_result* out);

/* A.11 voxel decimation: writes the kept ORIGINAL indices ascending to
 * keep_idx (capacity n); with use_average also the per-voxel means (else the
 * kept points' own coordinates) to ox/oy/oz. Returns the output count. */
size_t orc_voxel_decimate(const float* x, const float* y, const float* z, size_t n,
                          float resolution, int use_average, uint32_t* keep_idx, float* ox,
                          float* oy, float* oz);

/* A.13 FilterEdgesPlanes (normative restatement, see icp_oracle.c): layer[i] bit 0 "edges", bit 1 "planes",
 * bit 2 "full_decim".  Returns the number of classified voxels. */
size_t orc_filter_edges_planes(const float* x, const float* y, const float* z, size_t n, float resolution,
                               uint32_t full_decimation, uint32_t voxel_decimation, float max_e2_e0,
                               float max_e1_e0, float min_e2_e0, float min_e1_e0, uint32_t min_points,
                               uint8_t* layer);

#ifdef __cplusplus
}
#endif
#endif
