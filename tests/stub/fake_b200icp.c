/* fake_b200icp.c -- TEST DOUBLE of the device side of include/b200icp.h, used
 * only by tests/test_host_logic_cpu.py through LD_PRELOAD so that the C++ host
 * module (LidarOdometry mirror) can be exercised on a box without a GPU.
 * It is not a CPU fallback of the product: it computes nothing.  A "cloud"
 * remembers its size and its first point; "align" reports the translation
 * between the first points of the two clouds (so a test can script the
 * odometry it wants the front-end logic to see) with quality taken from the
 * first point's z coordinate of the `to` cloud (z0 >= 0 ? 1.0 : 0.1). */
#include <stdlib.h>
#include <string.h>
#include "../../include/b200icp.h"

struct b200icp { b200icp_params_t P; };
struct b200icp_cloud { size_t n; float p0[3]; };
static unsigned long g_align_calls = 0, g_batch_calls = 0, g_voxel_calls = 0, g_upload_calls = 0;
const char* b200icp_last_error(void) { return "test double: no error"; }
void b200icp_default_params(b200icp_params_t* p)
{
    memset(p, 0, sizeof(*p));
    p->max_iterations = 100, p->min_abs_step_trans = 5e-5, p->min_abs_step_rot = 1e-5;
    p->solver_max_iterations = 20, p->gn_min_delta = 1e-10, p->distance_threshold = 0.7;
    p->plane_eigen_threshold = 0.07, p->knn = 6, p->min_plane_points = 3;
    p->quality_threshold_distance = 0.1, p->cov_fd_step = 1e-7;
    p->use_scale_outlier_detector = 1, p->scale_outlier_threshold = 1.1, p->robust_kernel_scale = 400.0;
}

int b200icp_device_count(void) { return 1; }
int b200icp_create(const b200icp_params_t* p, int device, b200icp_t** out)
{
    (void)device;
    b200icp_t* o = (b200icp_t*)calloc(1, sizeof(*o));
    if (p) o->P = *p;
    *out = o;
    return B200ICP_OK;
}
void b200icp_destroy(b200icp_t* icp) { free(icp); }
int b200icp_cloud_upload(b200icp_t* icp, const float* x, const float* y, const float* z, size_t n,
                         float search_radius, b200icp_cloud_t** out)
{
    (void)icp, (void)search_radius;
    b200icp_cloud_t* c = (b200icp_cloud_t*)calloc(1, sizeof(*c));
    c->n = n;
    if (n) c->p0[0] = x[0], c->p0[1] = y[0], c->p0[2] = z[0];
    *out = c;
    __atomic_add_fetch(&g_upload_calls, 1, __ATOMIC_RELAXED);
    return B200ICP_OK;
}
unsigned long fake_upload_calls(void) { return g_upload_calls; }
/* "match": every second local point is paired; its plane is (first local point + i * 0.01 along x, moved by the
 * translation of the pose; normal +z) */
int b200icp_match(b200icp_t* icp, const b200icp_cloud_t* from_global, const b200icp_cloud_t* to_local,
                  const double* pose6, uint8_t* paired, uint32_t* nn_idx, uint32_t* nn_cnt, double* centroid,
                  double* normal, uint32_t* n_pairings)
{
    (void)icp, (void)from_global, (void)nn_idx, (void)nn_cnt;
    uint32_t np = 0;
    for (size_t i = 0; i < to_local->n; i++)
    {
        const int p = (i % 2) == 0;
        if (paired) paired[i] = (uint8_t)p;
        np += (uint32_t)p;
        if (centroid)
        {
            centroid[3 * i] = (double)to_local->p0[0] + 0.01 * (double)i + (pose6 ? pose6[0] : 0.0);
            centroid[3 * i + 1] = (double)to_local->p0[1] + (pose6 ? pose6[1] : 0.0);
            centroid[3 * i + 2] = (double)to_local->p0[2] + (pose6 ? pose6[2] : 0.0);
        }
        if (normal) normal[3 * i] = 0.0, normal[3 * i + 1] = 0.0, normal[3 * i + 2] = 1.0;
    }
    if (n_pairings) *n_pairings = np;
    return B200ICP_OK;
}
int b200icp_cloud_upload_raw(b200icp_t* icp, const float* x, const float* y, const float* z, size_t n,
                             b200icp_cloud_t** out)
{
    return b200icp_cloud_upload(icp, x, y, z, n, 0.f, out);
}
void   b200icp_cloud_free(b200icp_cloud_t* c) { free(c); }
size_t b200icp_cloud_size(const b200icp_cloud_t* c) { return c ? c->n : 0; }
size_t b200icp_cloud_device_bytes(const b200icp_cloud_t* c) { return c ? 100 * c->n : 0; }
static unsigned long g_downloads = 0;
unsigned long        fake_download_calls(void) { return g_downloads; }
int b200icp_cloud_download(const b200icp_cloud_t* c, float* x, float* y, float* z)
{
    if (!c || !x || !y || !z) return -1;
    g_downloads++;
    for (size_t i = 0; i < c->n; i++) x[i] = c->p0[0], y[i] = c->p0[1], z[i] = c->p0[2];
    return 0;
}
int b200icp_voxel_decimate(b200icp_t* icp, const b200icp_cloud_t* in, float resolution, int use_average,
                           float search_radius, b200icp_cloud_t** out, uint32_t* keep_idx)
{
    (void)icp, (void)resolution, (void)use_average, (void)search_radius, (void)keep_idx;
    b200icp_cloud_t* c = (b200icp_cloud_t*)calloc(1, sizeof(*c));
    *c = *in;
    c->n = in->n / 2; /* visible effect of the filter stage */
    *out = c;
    g_voxel_calls++;
    return B200ICP_OK;
}
/* "FilterEdgesPlanes": layers of n/4, n/3 and n/2 points, visible at the seam; the parameters are recorded */
static b200icp_edges_planes_params_t g_last_ep;
const b200icp_edges_planes_params_t* fake_last_edges_planes_params(void) { return &g_last_ep; }
void b200icp_edges_planes_defaults(b200icp_edges_planes_params_t* p)
{
    p->voxel_filter_resolution = 1.0f, p->full_pointcloud_decimation = 10, p->voxel_filter_decimation = 10;
    p->voxel_filter_max_e2_e0 = 30.f, p->voxel_filter_max_e1_e0 = 30.f;
    p->voxel_filter_min_e2_e0 = 80.f, p->voxel_filter_min_e1_e0 = 80.f, p->min_points_per_voxel = 5;
}
int b200icp_filter_edges_planes(b200icp_t* icp, const b200icp_cloud_t* in, const b200icp_edges_planes_params_t* params,
                                float search_radius, b200icp_cloud_t* layers_out[3], uint8_t* layer_flags_out,
                                uint32_t* n_classified_voxels_out)
{
    (void)icp, (void)search_radius, (void)layer_flags_out;
    g_last_ep = *params;
    const size_t div[3] = {4, 3, 2};
    for (int l = 0; l < 3; l++)
    {
        b200icp_cloud_t* c = (b200icp_cloud_t*)calloc(1, sizeof(*c));
        *c = *in;
        c->n = in->n / div[l];
        layers_out[l] = c;
    }
    if (n_classified_voxels_out) *n_classified_voxels_out = 1;
    return B200ICP_OK;
}
static void fake_result(const b200icp_cloud_t* from, const b200icp_cloud_t* to, const double* guess,
                        b200icp_result_t* r)
{
    memset(r, 0, sizeof(*r));
    for (int i = 0; i < 3; i++) r->pose[i] = (double)to->p0[i] - (double)from->p0[i];
    r->pose[2] = 0.0;
    r->pose[3] = guess ? guess[3] : 0.0; /* yaw: echo the guess */
    r->R[0] = r->R[4] = r->R[8] = 1.0;
    for (int i = 0; i < 3; i++) r->t[i] = r->pose[i];
    for (int i = 0; i < 6; i++) r->cov[i * 7] = 1e-4;
    r->quality = to->p0[2] >= 0.f ? 1.0 : 0.1;
    r->n_iterations = 3;
    r->termination_reason = B200ICP_TERM_STALLED;
    r->n_pairings = (uint32_t)to->n;
}
int b200icp_align(b200icp_t* icp, const b200icp_cloud_t* from, const b200icp_cloud_t* to,
                  const double guess6[6], b200icp_result_t* out)
{
    (void)icp;
    __atomic_add_fetch(&g_align_calls, 1, __ATOMIC_RELAXED);
    fake_result(from, to, guess6, out);
    return B200ICP_OK;
}
void b200icp_call_params_of(const b200icp_params_t* p, b200icp_call_params_t* out)
{
    memset(out, 0, sizeof(*out));
    out->max_iterations = p->max_iterations;
    out->min_abs_step_trans = p->min_abs_step_trans, out->min_abs_step_rot = p->min_abs_step_rot;
}
static unsigned long g_last_call_max_iterations = 0;
unsigned long        fake_last_call_max_iterations(void) { return g_last_call_max_iterations; }
int b200icp_align_with(b200icp_t* icp, const b200icp_cloud_t* from, const b200icp_cloud_t* to,
                       const double guess6[6], const b200icp_call_params_t* call, b200icp_result_t* out)
{
    (void)icp;
    __atomic_add_fetch(&g_align_calls, 1, __ATOMIC_RELAXED);
    if (call) g_last_call_max_iterations = call->max_iterations;
    fake_result(from, to, guess6, out);
    return B200ICP_OK;
}
int b200icp_align_batch(b200icp_t* icp, size_t n, const b200icp_cloud_t* const* from,
                        const b200icp_cloud_t* const* to, const double* guesses6, b200icp_result_t* out)
{
    (void)icp;
    g_batch_calls++;
    for (size_t i = 0; i < n; i++) fake_result(from[i], to[i], guesses6 + 6 * i, out + i);
    return B200ICP_OK;
}
unsigned long fake_align_calls(void) { return g_align_calls; }
unsigned long fake_batch_calls(void) { return g_batch_calls; }
unsigned long fake_voxel_calls(void) { return g_voxel_calls; }
