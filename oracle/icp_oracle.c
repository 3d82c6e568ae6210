/* icp_oracle.c -- CPU restatement of the reference ICP hot path (plain C99).
 *
 * TEST INFRASTRUCTURE ONLY (see icp_oracle.h).  PARITY UNPINNED: follows the
 * frozen spec of SURVEY.md Appendix A because the upstream arithmetic
 * (mp2p_icp, MRPT, nanoflann; CMakeLists.txt:17-24 of the reference) is not
 * vendored nor installable here and the reference ships no golden vectors.
 *
 * Build: see oracle/Makefile  (-O3 -march=native -ffp-contract=off: the float
 * distance path and the double plane-fit path must not be FMA-contracted, so
 * that the CUDA path, built with -fmad=false, can be bit-identical).
 */
#include "icp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ utils */
static inline uint32_t f2u(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
static inline float u2f(uint32_t u)
{
    float f;
    memcpy(&f, &u, 4);
    return f;
}
/* (d2, index) lexicographic key; valid for non-negative floats (A.4) */
static inline uint64_t make_key(float d2, uint32_t idx) { return ((uint64_t)f2u(d2) << 32) | idx; }

/* A.3: fixed float op order d=0,1,2, no contraction (-ffp-contract=off) */
static inline float dist2f(float qx, float qy, float qz, float px, float py, float pz)
{
    const float dx = qx - px, dy = qy - py, dz = qz - pz;
    const float a = dx * dx;
    const float b = dy * dy;
    const float c = dz * dz;
    const float ab = a + b;
    return ab + c;
}

void orc_default_params(orc_params* p)
{
    /* values of params/icp-settings-regular.yaml:10-46 */
    memset(p, 0, sizeof(*p));
    p->max_iterations = 100;
    p->min_abs_step_trans = 5e-5;
    p->min_abs_step_rot = 1e-5;
    p->use_scale_outlier_detector = 1;
    p->scale_outlier_threshold = 1.1;
    p->use_robust_kernel = 0;
    p->robust_kernel_param = 0.1 * M_PI / 180.0;
    p->robust_kernel_scale = 400.0;
    p->solver_kind = ORC_SOLVER_GAUSS_NEWTON;
    p->solver_max_iterations = 20;
    p->gn_min_delta = 1e-10; /* A.6 normative */
    p->matcher_kind = ORC_MATCHER_POINT2PLANE;
    p->distance_threshold = 0.70;
    p->plane_eigen_threshold = 0.07;
    p->knn = 6;
    p->min_plane_points = 3; /* A.5 */
    p->run_from_iteration = 0;
    p->run_up_to_iteration = 0;
    p->quality_threshold_distance = 0.10;
    p->cov_fd_step = 1e-7; /* A.9 */
}

/* ------------------------------------------------------------------ cloud */
typedef struct
{
    int32_t  left, right; /* children, -1 for leaf */
    uint32_t lo, hi;      /* leaf: range in the permuted arrays */
    int32_t  dim;
    float    divlow, divhigh;
} kd_node;

struct orc_cloud
{
    size_t n;
    float *x, *y, *z;
    /* kd-tree (row I): leaf size 10, float, built lazily */
    int       kd_built;
    kd_node*  nodes;
    size_t    n_nodes, cap_nodes;
    uint32_t* perm;           /* permuted position -> original index */
    float *   px, *py, *pz;   /* coordinates in permuted order */
};

orc_cloud* orc_cloud_create(const float* x, const float* y, const float* z, size_t n)
{
    orc_cloud* c = (orc_cloud*)calloc(1, sizeof(orc_cloud));
    c->n = n;
    c->x = (float*)malloc(sizeof(float) * (n ? n : 1));
    c->y = (float*)malloc(sizeof(float) * (n ? n : 1));
    c->z = (float*)malloc(sizeof(float) * (n ? n : 1));
    if (n)
    {
        memcpy(c->x, x, sizeof(float) * n);
        memcpy(c->y, y, sizeof(float) * n);
        memcpy(c->z, z, sizeof(float) * n);
    }
    return c;
}
void orc_cloud_free(orc_cloud* c)
{
    if (!c) return;
    free(c->x), free(c->y), free(c->z);
    free(c->nodes), free(c->perm), free(c->px), free(c->py), free(c->pz);
    free(c);
}
size_t orc_cloud_size(const orc_cloud* c) { return c->n; }

/* ------------------------------------------------------------ top-k state */
#define ORC_MAX_K 32
typedef struct
{
    uint32_t k;
    uint64_t key[ORC_MAX_K]; /* ascending */
} topk;

static inline void topk_init(topk* t, uint32_t k, float max_d2)
{
    t->k = k;
    const uint64_t sentinel = ((uint64_t)f2u(max_d2) << 32) | 0xFFFFFFFFu;
    for (uint32_t i = 0; i < k; i++) t->key[i] = sentinel;
}
static inline void topk_push(topk* t, uint64_t key)
{
    const uint32_t k = t->k;
    if (!(key < t->key[k - 1])) return;
    uint32_t i = k - 1;
    while (i > 0 && t->key[i - 1] > key)
    {
        t->key[i] = t->key[i - 1];
        i--;
    }
    t->key[i] = key;
}
static inline float topk_worst_d2(const topk* t) { return u2f((uint32_t)(t->key[t->k - 1] >> 32)); }
static void topk_write(const topk* t, float max_d2, uint32_t* idx_out, float* d2_out)
{
    const uint64_t sentinel = ((uint64_t)f2u(max_d2) << 32) | 0xFFFFFFFFu;
    for (uint32_t i = 0; i < t->k; i++)
    {
        if (t->key[i] == sentinel)
        {
            idx_out[i] = ORC_INVALID_IDX;
            if (d2_out) d2_out[i] = INFINITY;
        }
        else
        {
            idx_out[i] = (uint32_t)(t->key[i] & 0xFFFFFFFFu);
            if (d2_out) d2_out[i] = u2f((uint32_t)(t->key[i] >> 32));
        }
    }
}

/* ------------------------------------------------------- brute-force kNN */
static void knn_brute_one(const orc_cloud* ref, float qx, float qy, float qz, topk* t)
{
    for (size_t j = 0; j < ref->n; j++)
    {
        const float d2 = dist2f(qx, qy, qz, ref->x[j], ref->y[j], ref->z[j]);
        /* NaN never passes: its bit pattern is above +inf */
        topk_push(t, make_key(d2, (uint32_t)j));
    }
}

void orc_knn_brute(const orc_cloud* ref, const float* qx, const float* qy, const float* qz,
                   size_t nq, uint32_t k, float max_d2, uint32_t* idx_out, float* d2_out)
{
    if (k > ORC_MAX_K) k = ORC_MAX_K;
    for (size_t i = 0; i < nq; i++)
    {
        topk t;
        topk_init(&t, k, max_d2);
        knn_brute_one(ref, qx[i], qy[i], qz[i], &t);
        topk_write(&t, max_d2, idx_out + i * k, d2_out ? d2_out + i * k : NULL);
    }
}

/* ---------------------------------------------------------------- kd-tree */
#define KD_LEAF 10

static float kd_coord(const orc_cloud* c, uint32_t orig, int dim)
{
    return dim == 0 ? c->x[orig] : (dim == 1 ? c->y[orig] : c->z[orig]);
}

/* quickselect on perm[lo,hi) by coordinate dim so that perm[mid] is the
 * (mid-lo)-th smallest; ties broken by original index for determinism */
static int kd_less(const orc_cloud* c, uint32_t a, uint32_t b, int dim)
{
    const float fa = kd_coord(c, a, dim), fb = kd_coord(c, b, dim);
    if (fa < fb) return 1;
    if (fa > fb) return 0;
    return a < b;
}
static void kd_select(orc_cloud* c, size_t lo_, size_t hi_, size_t mid_, int dim)
{
    uint32_t* p = c->perm;
    ptrdiff_t l = (ptrdiff_t)lo_, r = (ptrdiff_t)hi_ - 1;
    const ptrdiff_t mid = (ptrdiff_t)mid_;
    while (l < r)
    {
        /* median-of-three pivot (keys are unique: ties broken by index) */
        const uint32_t pa = p[l], pb = p[l + (r - l) / 2], pd = p[r];
        uint32_t piv;
        if (kd_less(c, pa, pb, dim))
            piv = kd_less(c, pb, pd, dim) ? pb : (kd_less(c, pa, pd, dim) ? pd : pa);
        else
            piv = kd_less(c, pa, pd, dim) ? pa : (kd_less(c, pb, pd, dim) ? pd : pb);
        ptrdiff_t i = l, j = r;
        while (i <= j)
        {
            while (kd_less(c, p[i], piv, dim)) i++;
            while (kd_less(c, piv, p[j], dim)) j--;
            if (i <= j)
            {
                const uint32_t tmp = p[i];
                p[i] = p[j];
                p[j] = tmp;
                i++;
                j--;
            }
        }
        if (mid <= j)
            r = j;
        else if (mid >= i)
            l = i;
        else
            return;
    }
}

static int32_t kd_new_node(orc_cloud* c)
{
    if (c->n_nodes == c->cap_nodes)
    {
        c->cap_nodes = c->cap_nodes ? c->cap_nodes * 2 : 1024;
        c->nodes = (kd_node*)realloc(c->nodes, sizeof(kd_node) * c->cap_nodes);
    }
    return (int32_t)c->n_nodes++;
}

static int32_t kd_build_rec(orc_cloud* c, size_t lo, size_t hi)
{
    const int32_t id = kd_new_node(c);
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t i = lo; i < hi; i++)
        for (int d = 0; d < 3; d++)
        {
            const float v = kd_coord(c, c->perm[i], d);
            if (v < mn[d]) mn[d] = v;
            if (v > mx[d]) mx[d] = v;
        }
    int dim = 0;
    float span = mx[0] - mn[0];
    for (int d = 1; d < 3; d++)
        if (mx[d] - mn[d] > span) span = mx[d] - mn[d], dim = d;
    if (hi - lo <= KD_LEAF || !(span > 0.0f))
    {
        kd_node* nd = &c->nodes[id];
        nd->left = nd->right = -1;
        nd->lo = (uint32_t)lo, nd->hi = (uint32_t)hi;
        nd->dim = 0, nd->divlow = nd->divhigh = 0;
        return id;
    }
    const size_t mid = lo + (hi - lo) / 2;
    kd_select(c, lo, hi, mid, dim);
    float divlow = -INFINITY, divhigh = INFINITY;
    for (size_t i = lo; i < mid; i++)
    {
        const float v = kd_coord(c, c->perm[i], dim);
        if (v > divlow) divlow = v;
    }
    for (size_t i = mid; i < hi; i++)
    {
        const float v = kd_coord(c, c->perm[i], dim);
        if (v < divhigh) divhigh = v;
    }
    const int32_t l = kd_build_rec(c, lo, mid);
    const int32_t r = kd_build_rec(c, mid, hi);
    kd_node* nd = &c->nodes[id];
    nd->left = l, nd->right = r, nd->lo = nd->hi = 0;
    nd->dim = dim, nd->divlow = divlow, nd->divhigh = divhigh;
    return id;
}

static void kd_build(orc_cloud* c)
{
    if (c->kd_built) return;
    c->perm = (uint32_t*)malloc(sizeof(uint32_t) * (c->n ? c->n : 1));
    /* NaN / inf points can never be a neighbour (A.4): leave them out */
    size_t m = 0;
    for (size_t i = 0; i < c->n; i++)
        if (isfinite(c->x[i]) && isfinite(c->y[i]) && isfinite(c->z[i])) c->perm[m++] = (uint32_t)i;
    if (m) kd_build_rec(c, 0, m);
    c->px = (float*)malloc(sizeof(float) * (m ? m : 1));
    c->py = (float*)malloc(sizeof(float) * (m ? m : 1));
    c->pz = (float*)malloc(sizeof(float) * (m ? m : 1));
    for (size_t i = 0; i < m; i++)
    {
        c->px[i] = c->x[c->perm[i]];
        c->py[i] = c->y[c->perm[i]];
        c->pz[i] = c->z[c->perm[i]];
    }
    c->kd_built = 1;
}

typedef struct
{
    const orc_cloud* c;
    float            q[3];
    topk*            t;
} kd_query;

static void kd_search_rec(const kd_query* Q, int32_t id, double off0, double off1, double off2)
{
    const orc_cloud* c = Q->c;
    const kd_node* nd = &c->nodes[id];
    if (nd->left < 0)
    {
        for (uint32_t i = nd->lo; i < nd->hi; i++)
        {
            const float d2 = dist2f(Q->q[0], Q->q[1], Q->q[2], c->px[i], c->py[i], c->pz[i]);
            topk_push(Q->t, make_key(d2, c->perm[i]));
        }
        return;
    }
    const int dim = nd->dim;
    const float v = Q->q[dim];
    int32_t nearc, farc;
    double cut;
    if (v <= 0.5f * (nd->divlow + nd->divhigh))
    {
        nearc = nd->left, farc = nd->right;
        cut = (double)nd->divhigh - (double)v;
    }
    else
    {
        nearc = nd->right, farc = nd->left;
        cut = (double)v - (double)nd->divlow;
    }
    if (cut < 0) cut = 0;
    kd_search_rec(Q, nearc, off0, off1, off2);
    double o[3] = {off0, off1, off2};
    if (cut > o[dim]) o[dim] = cut;
    const double bound = o[0] * o[0] + o[1] * o[1] + o[2] * o[2];
    /* conservative vs. float rounding of dist2f, and '<=' so that equal-d2
     * candidates with a lower index are still seen (tie rule, A.4) */
    const double worst = (double)topk_worst_d2(Q->t);
    if (bound <= worst * (1.0 + 1e-6)) kd_search_rec(Q, farc, o[0], o[1], o[2]);
}

static void knn_kd_one(const orc_cloud* ref, float qx, float qy, float qz, topk* t)
{
    if (!ref->n_nodes) return;
    if (!(isfinite(qx) && isfinite(qy) && isfinite(qz))) return;
    kd_query Q;
    Q.c = ref;
    Q.q[0] = qx, Q.q[1] = qy, Q.q[2] = qz;
    Q.t = t;
    kd_search_rec(&Q, 0, 0.0, 0.0, 0.0);
}

void orc_knn_kdtree(orc_cloud* ref, const float* qx, const float* qy, const float* qz, size_t nq,
                    uint32_t k, float max_d2, uint32_t* idx_out, float* d2_out)
{
    if (k > ORC_MAX_K) k = ORC_MAX_K;
    kd_build(ref);
    for (size_t i = 0; i < nq; i++)
    {
        topk t;
        topk_init(&t, k, max_d2);
        knn_kd_one(ref, qx[i], qy[i], qz[i], &t);
        topk_write(&t, max_d2, idx_out + i * k, d2_out ? d2_out + i * k : NULL);
    }
}

/* ------------------------------------------------------------- pose maths */
/* MRPT convention (A.2): R = Rz(yaw) Ry(pitch) Rx(roll) */
void orc_pose_to_Rt(const double pose[6], double R[9], double t[3])
{
    const double cy = cos(pose[3]), sy = sin(pose[3]);
    const double cp = cos(pose[4]), sp = sin(pose[4]);
    const double cr = cos(pose[5]), sr = sin(pose[5]);
    R[0] = cy * cp, R[1] = cy * sp * sr - sy * cr, R[2] = cy * sp * cr + sy * sr;
    R[3] = sy * cp, R[4] = sy * sp * sr + cy * cr, R[5] = sy * sp * cr - cy * sr;
    R[6] = -sp, R[7] = cp * sr, R[8] = cp * cr;
    t[0] = pose[0], t[1] = pose[1], t[2] = pose[2];
}

void orc_Rt_to_pose(const double R[9], const double t[3], double pose[6])
{
    pose[0] = t[0], pose[1] = t[1], pose[2] = t[2];
    const double cpitch = sqrt(R[0] * R[0] + R[3] * R[3]);
    double yaw, pitch, roll;
    pitch = atan2(-R[6], cpitch);
    if (cpitch < 1e-12)
    { /* gimbal lock: roll = 0, everything into yaw */
        roll = 0.0;
        yaw = atan2(-R[1], R[4]);
    }
    else
    {
        yaw = atan2(R[3], R[0]);
        roll = atan2(R[7], R[8]);
    }
    pose[3] = yaw, pose[4] = pitch, pose[5] = roll;
}

void orc_transform_points(const double R[9], const double t[3], const float* x, const float* y,
                          const float* z, size_t n, float* ox, float* oy, float* oz)
{
    for (size_t i = 0; i < n; i++)
    {
        const double px = x[i], py = y[i], pz = z[i];
        ox[i] = (float)(((R[0] * px + R[1] * py) + R[2] * pz) + t[0]);
        oy[i] = (float)(((R[3] * px + R[4] * py) + R[5] * pz) + t[1]);
        oz[i] = (float)(((R[6] * px + R[7] * py) + R[8] * pz) + t[2]);
    }
}

static void mat3_mul(const double A[9], const double B[9], double C[9])
{
    double T[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            T[i * 3 + j] = (A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j]) + A[i * 3 + 2] * B[6 + j];
    memcpy(C, T, sizeof(T));
}
static void mat3_vec(const double A[9], const double v[3], double o[3])
{
    double T[3];
    for (int i = 0; i < 3; i++) T[i] = (A[i * 3] * v[0] + A[i * 3 + 1] * v[1]) + A[i * 3 + 2] * v[2];
    o[0] = T[0], o[1] = T[1], o[2] = T[2];
}

void orc_se3_compose(const double Ra[9], const double ta[3], const double Rb[9],
                     const double tb[3], double R[9], double t[3])
{
    double tt[3];
    mat3_vec(Ra, tb, tt);
    tt[0] += ta[0], tt[1] += ta[1], tt[2] += ta[2];
    mat3_mul(Ra, Rb, R);
    t[0] = tt[0], t[1] = tt[1], t[2] = tt[2];
}

void orc_se3_inverse_compose(const double Ra[9], const double ta[3], const double Rb[9],
                             const double tb[3], double R[9], double t[3])
{
    /* a^-1 * b : R = Ra^T Rb, t = Ra^T (tb - ta) */
    double RaT[9] = {Ra[0], Ra[3], Ra[6], Ra[1], Ra[4], Ra[7], Ra[2], Ra[5], Ra[8]};
    double d[3] = {tb[0] - ta[0], tb[1] - ta[1], tb[2] - ta[2]};
    double tt[3];
    mat3_vec(RaT, d, tt);
    mat3_mul(RaT, Rb, R);
    t[0] = tt[0], t[1] = tt[1], t[2] = tt[2];
}

/* coefficients A = sin(th)/th, B = (1-cos th)/th^2, C = (th - sin th)/th^3 */
static void so3_coeffs(double th2, double* A, double* B, double* C)
{
    if (th2 < 1e-8)
    {
        *A = 1.0 - th2 / 6.0;
        *B = 0.5 - th2 / 24.0;
        *C = 1.0 / 6.0 - th2 / 120.0;
    }
    else
    {
        const double th = sqrt(th2);
        const double s = sin(th), c = cos(th);
        *A = s / th;
        *B = (1.0 - c) / th2;
        *C = (th - s) / (th2 * th);
    }
}

void orc_se3_exp(const double eps[6], double R[9], double t[3])
{
    const double vx = eps[0], vy = eps[1], vz = eps[2];
    const double wx = eps[3], wy = eps[4], wz = eps[5];
    const double th2 = (wx * wx + wy * wy) + wz * wz;
    double A, B, C;
    so3_coeffs(th2, &A, &B, &C);
    /* W = [w]x, W2 = W*W = w w^T - th2 I */
    const double W[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    const double W2[9] = {wx * wx - th2, wx * wy, wx * wz, wx * wy, wy * wy - th2,
                          wy * wz,       wx * wz, wy * wz, wz * wz - th2};
    double V[9];
    for (int i = 0; i < 9; i++)
    {
        const double I = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
        R[i] = (I + A * W[i]) + B * W2[i];
        V[i] = (I + B * W[i]) + C * W2[i];
    }
    const double v[3] = {vx, vy, vz};
    mat3_vec(V, v, t);
}

void orc_se3_log(const double R[9], const double t[3], double eps[6])
{
    /* w = vee(R - R^T)/2 = sin(th) * axis ; cos(th) = (tr-1)/2 */
    double w[3] = {0.5 * (R[7] - R[5]), 0.5 * (R[2] - R[6]), 0.5 * (R[3] - R[1])};
    const double s = sqrt((w[0] * w[0] + w[1] * w[1]) + w[2] * w[2]);
    const double c = 0.5 * (((R[0] + R[4]) + R[8]) - 1.0);
    const double th = atan2(s, c);
    double om[3];
    if (s < 1e-8 && c > 0)
    { /* th ~ 0: th/sin(th) -> 1 + th^2/6 */
        const double k = 1.0 + th * th / 6.0;
        om[0] = k * w[0], om[1] = k * w[1], om[2] = k * w[2];
    }
    else if (s < 1e-8)
    { /* th ~ pi: axis from the diagonal of (R + I)/2 = a a^T */
        double a[3] = {sqrt(fmax(0.0, 0.5 * (R[0] + 1.0))), sqrt(fmax(0.0, 0.5 * (R[4] + 1.0))),
                       sqrt(fmax(0.0, 0.5 * (R[8] + 1.0)))};
        /* fix signs relative to the largest component */
        int m = 0;
        if (a[1] > a[m]) m = 1;
        if (a[2] > a[m]) m = 2;
        for (int i = 0; i < 3; i++)
            if (i != m && (R[m * 3 + i] + R[i * 3 + m]) < 0) a[i] = -a[i];
        om[0] = th * a[0], om[1] = th * a[1], om[2] = th * a[2];
    }
    else
    {
        const double k = th / s;
        om[0] = k * w[0], om[1] = k * w[1], om[2] = k * w[2];
    }
    /* v = V^-1 t ; V^-1 = I - W/2 + D W^2 */
    const double th2 = (om[0] * om[0] + om[1] * om[1]) + om[2] * om[2];
    double D;
    if (th2 < 1e-8)
        D = 1.0 / 12.0 + th2 / 720.0;
    else
    {
        const double thn = sqrt(th2);
        D = (1.0 - (thn * sin(thn)) / (2.0 * (1.0 - cos(thn)))) / th2;
    }
    const double wx = om[0], wy = om[1], wz = om[2];
    const double W[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    const double W2[9] = {wx * wx - th2, wx * wy, wx * wz, wx * wy, wy * wy - th2,
                          wy * wz,       wx * wz, wy * wz, wz * wz - th2};
    double Vi[9];
    for (int i = 0; i < 9; i++)
    {
        const double I = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
        Vi[i] = (I - 0.5 * W[i]) + D * W2[i];
    }
    double v[3];
    mat3_vec(Vi, t, v);
    eps[0] = v[0], eps[1] = v[1], eps[2] = v[2];
    eps[3] = om[0], eps[4] = om[1], eps[5] = om[2];
}

/* --------------------------------------------------- small linear algebra */
/* cyclic Jacobi for symmetric n x n (n<=4); evals ascending, evecs columns.
 * The 3x3 instance is the plane-fit eigen solver of row J / A.5 and is
 * replicated operation-for-operation by the CUDA path. */
static void jacobi_sym(int n, double* A /* n*n, destroyed */, double* evals, double* V)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) V[i * n + j] = (i == j) ? 1.0 : 0.0;
    double frob = 0;
    for (int i = 0; i < n * n; i++) frob += A[i] * A[i];
    const double tol = 1e-30 * frob;
    for (int sweep = 0; sweep < 30; sweep++)
    {
        double off = 0;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) off += A[p * n + q] * A[p * n + q];
        if (!(off > tol)) break;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++)
            {
                const double apq = A[p * n + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
                const double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(tt * tt + 1.0);
                const double s = tt * c;
                A[p * n + p] = A[p * n + p] - tt * apq;
                A[q * n + q] = A[q * n + q] + tt * apq;
                A[p * n + q] = A[q * n + p] = 0.0;
                for (int r = 0; r < n; r++)
                {
                    if (r != p && r != q)
                    {
                        const double arp = A[r * n + p], arq = A[r * n + q];
                        const double nrp = c * arp - s * arq;
                        const double nrq = s * arp + c * arq;
                        A[r * n + p] = A[p * n + r] = nrp;
                        A[r * n + q] = A[q * n + r] = nrq;
                    }
                    const double vrp = V[r * n + p], vrq = V[r * n + q];
                    V[r * n + p] = c * vrp - s * vrq;
                    V[r * n + q] = s * vrp + c * vrq;
                }
            }
    }
    for (int i = 0; i < n; i++) evals[i] = A[i * n + i];
    /* ascending, stable exchange sort with strict '>' */
    for (int i = 0; i < n; i++)
        for (int j = 0; j + 1 < n - i; j++)
            if (evals[j] > evals[j + 1])
            {
                const double te = evals[j];
                evals[j] = evals[j + 1], evals[j + 1] = te;
                for (int r = 0; r < n; r++)
                {
                    const double tv = V[r * n + j];
                    V[r * n + j] = V[r * n + j + 1], V[r * n + j + 1] = tv;
                }
            }
}

void orc_eig3_sym(const double C[9], double evals[3], double evecs[9])
{
    double A[9];
    memcpy(A, C, sizeof(A));
    jacobi_sym(3, A, evals, evecs);
}

/* Column-pivoting Householder QR, 6x6 (A.6). Rank threshold like Eigen's
 * default: |R_kk| <= eps * n * |R_00| ends the rank. Returns the rank. */
int orc_qr_solve6(const double Ain[36], const double bin[6], double x[6])
{
    enum { N = 6 };
    double A[36], b[6];
    int perm[N];
    memcpy(A, Ain, sizeof(A));
    memcpy(b, bin, sizeof(b));
    for (int i = 0; i < N; i++) perm[i] = i;
    int rank = N;
    double r00 = 0;
    for (int k = 0; k < N; k++)
    {
        /* pivot: remaining column with the largest norm */
        int best = k;
        double bestn = -1;
        for (int j = k; j < N; j++)
        {
            double s = 0;
            for (int i = k; i < N; i++) s += A[i * N + j] * A[i * N + j];
            if (s > bestn) bestn = s, best = j;
        }
        if (best != k)
        {
            for (int i = 0; i < N; i++)
            {
                const double tmp = A[i * N + k];
                A[i * N + k] = A[i * N + best], A[i * N + best] = tmp;
            }
            const int tp = perm[k];
            perm[k] = perm[best], perm[best] = tp;
        }
        const double normx = sqrt(bestn);
        if (k == 0) r00 = normx;
        if (!(normx > 2.220446049250313e-16 * N * r00) || normx == 0.0)
        {
            rank = k;
            break;
        }
        /* Householder vector v = x + sign(x0)|x| e0 */
        double v[N];
        const double x0 = A[k * N + k];
        const double alpha = (x0 >= 0) ? -normx : normx;
        for (int i = 0; i < N; i++) v[i] = (i < k) ? 0.0 : A[i * N + k];
        v[k] = x0 - alpha;
        double vnorm2 = 0;
        for (int i = k; i < N; i++) vnorm2 += v[i] * v[i];
        if (vnorm2 > 0)
        {
            for (int j = k; j < N; j++)
            {
                double dot = 0;
                for (int i = k; i < N; i++) dot += v[i] * A[i * N + j];
                const double f = 2.0 * dot / vnorm2;
                for (int i = k; i < N; i++) A[i * N + j] -= f * v[i];
            }
            double dot = 0;
            for (int i = k; i < N; i++) dot += v[i] * b[i];
            const double f = 2.0 * dot / vnorm2;
            for (int i = k; i < N; i++) b[i] -= f * v[i];
        }
    }
    double y[N];
    for (int i = 0; i < N; i++) y[i] = 0;
    for (int i = rank - 1; i >= 0; i--)
    {
        double s = b[i];
        for (int j = i + 1; j < rank; j++) s -= A[i * N + j] * y[j];
        y[i] = s / A[i * N + i];
    }
    for (int i = 0; i < N; i++) x[perm[i]] = y[i];
    return rank;
}

int orc_inverse6(const double A[36], double Ainv[36])
{
    int rank = 6;
    for (int c = 0; c < 6; c++)
    {
        double e[6] = {0, 0, 0, 0, 0, 0}, x[6];
        e[c] = 1.0;
        const int r = orc_qr_solve6(A, e, x);
        if (r < rank) rank = r;
        for (int i = 0; i < 6; i++) Ainv[i * 6 + c] = x[i];
    }
    return rank;
}

/* ---------------------------------------------------------------- matcher */
typedef void (*knn_fn)(const orc_cloud*, float, float, float, topk*);

size_t orc_match_point2plane(orc_cloud* global, const orc_cloud* local, const double R[9],
                             const double t[3], const orc_params* p, int use_kdtree,
                             uint8_t* paired, uint32_t* nn_idx, uint32_t* nn_cnt,
                             double* centroid, double* normal)
{
    const size_t nl = local->n;
    uint32_t k = p->knn;
    if (k > ORC_MAX_K) k = ORC_MAX_K;
    const float thr = (float)p->distance_threshold;
    const float thr2 = thr * thr; /* A.5: float product */
    if (use_kdtree) kd_build(global);
    knn_fn search = use_kdtree ? knn_kd_one : knn_brute_one;
    size_t npair = 0;
    for (size_t i = 0; i < nl; i++)
    {
        if (paired) paired[i] = 0;
        if (nn_cnt) nn_cnt[i] = 0;
        if (nn_idx)
            for (uint32_t j = 0; j < k; j++) nn_idx[i * k + j] = ORC_INVALID_IDX;
        if (!global->n || !k) continue;
        /* A.2 */
        const double px = local->x[i], py = local->y[i], pz = local->z[i];
        const float qx = (float)(((R[0] * px + R[1] * py) + R[2] * pz) + t[0]);
        const float qy = (float)(((R[3] * px + R[4] * py) + R[5] * pz) + t[1]);
        const float qz = (float)(((R[6] * px + R[7] * py) + R[8] * pz) + t[2]);
        topk tk;
        topk_init(&tk, k, thr2);
        search(global, qx, qy, qz, &tk);
        uint32_t idx[ORC_MAX_K];
        topk_write(&tk, thr2, idx, NULL);
        uint32_t m = 0;
        while (m < k && idx[m] != ORC_INVALID_IDX) m++;
        if (nn_cnt) nn_cnt[i] = m;
        if (nn_idx)
            for (uint32_t j = 0; j < m; j++) nn_idx[i * k + j] = idx[j];
        if (m < p->min_plane_points) continue;
        /* row J: mean and covariance (1/m) of the neighbours, f64, in
         * neighbour order */
        double sx = 0, sy = 0, sz = 0;
        for (uint32_t j = 0; j < m; j++)
        {
            sx += (double)global->x[idx[j]];
            sy += (double)global->y[idx[j]];
            sz += (double)global->z[idx[j]];
        }
        const double inv = 1.0 / (double)m;
        const double cx = sx * inv, cy = sy * inv, cz = sz * inv;
        double c00 = 0, c01 = 0, c02 = 0, c11 = 0, c12 = 0, c22 = 0;
        for (uint32_t j = 0; j < m; j++)
        {
            const double dx = (double)global->x[idx[j]] - cx;
            const double dy = (double)global->y[idx[j]] - cy;
            const double dz = (double)global->z[idx[j]] - cz;
            c00 += dx * dx, c01 += dx * dy, c02 += dx * dz;
            c11 += dy * dy, c12 += dy * dz, c22 += dz * dz;
        }
        double C[9] = {c00 * inv, c01 * inv, c02 * inv, c01 * inv, c11 * inv,
                       c12 * inv, c02 * inv, c12 * inv, c22 * inv};
        double ev[3], V[9];
        jacobi_sym(3, C, ev, V);
        if (ev[0] > p->plane_eigen_threshold * ev[2]) continue;
        double nx = V[0], ny = V[3], nz = V[6]; /* column 0 */
        /* normative sign: first non-zero component positive */
        const double lead = (nx != 0.0) ? nx : ((ny != 0.0) ? ny : nz);
        if (lead < 0) nx = -nx, ny = -ny, nz = -nz;
        const double dist =
            fabs((nx * ((double)qx - cx) + ny * ((double)qy - cy)) + nz * ((double)qz - cz));
        if (dist > p->distance_threshold) continue;
        if (paired) paired[i] = 1;
        if (centroid) centroid[i * 3] = cx, centroid[i * 3 + 1] = cy, centroid[i * 3 + 2] = cz;
        if (normal) normal[i * 3] = nx, normal[i * 3 + 1] = ny, normal[i * 3 + 2] = nz;
        npair++;
    }
    return npair;
}

size_t orc_match_points(orc_cloud* global, const orc_cloud* local, const double R[9],
                        const double t[3], double threshold, int use_kdtree, uint32_t* nn,
                        float* nn_d2)
{
    const float thr = (float)threshold;
    const float thr2 = thr * thr;
    if (use_kdtree) kd_build(global);
    knn_fn search = use_kdtree ? knn_kd_one : knn_brute_one;
    size_t np = 0;
    for (size_t i = 0; i < local->n; i++)
    {
        const double px = local->x[i], py = local->y[i], pz = local->z[i];
        const float qx = (float)(((R[0] * px + R[1] * py) + R[2] * pz) + t[0]);
        const float qy = (float)(((R[3] * px + R[4] * py) + R[5] * pz) + t[1]);
        const float qz = (float)(((R[6] * px + R[7] * py) + R[8] * pz) + t[2]);
        topk tk;
        topk_init(&tk, 1, thr2);
        if (global->n) search(global, qx, qy, qz, &tk);
        uint32_t id;
        float d2;
        topk_write(&tk, thr2, &id, &d2);
        /* strict '<' (A.8) */
        if (id != ORC_INVALID_IDX && !(d2 < thr2)) id = ORC_INVALID_IDX;
        if (nn) nn[i] = id;
        if (nn_d2) nn_d2[i] = (id == ORC_INVALID_IDX) ? INFINITY : d2;
        if (id != ORC_INVALID_IDX) np++;
    }
    return np;
}

double orc_quality_paired_ratio(orc_cloud* global, const orc_cloud* local, const double R[9],
                                const double t[3], double threshold, int use_kdtree)
{
    if (!local->n) return 0.0;
    const size_t np = orc_match_points(global, local, R, t, threshold, use_kdtree, NULL, NULL);
    return (double)np / (double)local->n;
}

/* ---------------------------------------------------------------- solvers */
static void apply_delta(const double delta[6], double R[9], double t[3])
{
    double dR[9], dt[3];
    orc_se3_exp(delta, dR, dt);
    orc_se3_compose(R, t, dR, dt, R, t); /* T <- T * exp(delta) */
}

int orc_gn_point2plane(const double* P, const double* Cc, const double* Nn, size_t np,
                       uint32_t max_iters, double min_delta, double R[9], double t[3])
{
    int it = 0;
    for (uint32_t iter = 0; iter < max_iters; iter++)
    {
        double H[36], g[6];
        memset(H, 0, sizeof(H));
        memset(g, 0, sizeof(g));
        for (size_t i = 0; i < np; i++)
        {
            const double* p = P + 3 * i;
            const double* c = Cc + 3 * i;
            const double* n = Nn + 3 * i;
            double Rp[3];
            mat3_vec(R, p, Rp);
            const double r =
                (n[0] * ((Rp[0] + t[0]) - c[0]) + n[1] * ((Rp[1] + t[1]) - c[1])) +
                n[2] * ((Rp[2] + t[2]) - c[2]);
            /* m = R^T n ; J = [m^T, (p x m)^T]  (rows K, L; A.6) */
            const double m[3] = {(R[0] * n[0] + R[3] * n[1]) + R[6] * n[2],
                                 (R[1] * n[0] + R[4] * n[1]) + R[7] * n[2],
                                 (R[2] * n[0] + R[5] * n[1]) + R[8] * n[2]};
            const double J[6] = {m[0],
                                 m[1],
                                 m[2],
                                 p[1] * m[2] - p[2] * m[1],
                                 p[2] * m[0] - p[0] * m[2],
                                 p[0] * m[1] - p[1] * m[0]};
            for (int a = 0; a < 6; a++)
            {
                g[a] += J[a] * r;
                for (int b = 0; b < 6; b++) H[a * 6 + b] += J[a] * J[b];
            }
        }
        double mg[6], delta[6];
        for (int a = 0; a < 6; a++) mg[a] = -g[a];
        orc_qr_solve6(H, mg, delta);
        apply_delta(delta, R, t);
        it++;
        double nd = 0;
        for (int a = 0; a < 6; a++) nd += delta[a] * delta[a];
        if (sqrt(nd) < min_delta) break;
    }
    return it;
}

int orc_gn_point2point(const double* P, const double* Q, size_t np, uint32_t max_iters,
                       double min_delta, double R[9], double t[3])
{
    int it = 0;
    for (uint32_t iter = 0; iter < max_iters; iter++)
    {
        double H[36], g[6];
        memset(H, 0, sizeof(H));
        memset(g, 0, sizeof(g));
        for (size_t i = 0; i < np; i++)
        {
            const double* p = P + 3 * i;
            const double* q = Q + 3 * i;
            double Rp[3];
            mat3_vec(R, p, Rp);
            const double r[3] = {(Rp[0] + t[0]) - q[0], (Rp[1] + t[1]) - q[1],
                                 (Rp[2] + t[2]) - q[2]};
            /* J (3x6) = [R, -R [p]x] */
            const double px[9] = {0, -p[2], p[1], p[2], 0, -p[0], -p[1], p[0], 0};
            double Rpx[9];
            mat3_mul(R, px, Rpx);
            double J[18];
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++)
                {
                    J[a * 6 + b] = R[a * 3 + b];
                    J[a * 6 + 3 + b] = -Rpx[a * 3 + b];
                }
            for (int a = 0; a < 6; a++)
            {
                g[a] += (J[a] * r[0] + J[6 + a] * r[1]) + J[12 + a] * r[2];
                for (int b = 0; b < 6; b++)
                    H[a * 6 + b] += (J[a] * J[b] + J[6 + a] * J[6 + b]) + J[12 + a] * J[12 + b];
            }
        }
        double mg[6], delta[6];
        for (int a = 0; a < 6; a++) mg[a] = -g[a];
        orc_qr_solve6(H, mg, delta);
        apply_delta(delta, R, t);
        it++;
        double nd = 0;
        for (int a = 0; a < 6; a++) nd += delta[a] * delta[a];
        if (sqrt(nd) < min_delta) break;
    }
    return it;
}

size_t orc_horn(const double* P, const double* Q, size_t np, const orc_params* prm,
                const double Rprior[9], double R[9], double t[3])
{
    if (!np) return 0;
    /* centroids over all pairs (A.10) */
    double pc[3] = {0, 0, 0}, qc[3] = {0, 0, 0};
    for (size_t i = 0; i < np; i++)
        for (int d = 0; d < 3; d++) pc[d] += P[3 * i + d], qc[d] += Q[3 * i + d];
    for (int d = 0; d < 3; d++) pc[d] /= (double)np, qc[d] /= (double)np;
    double S[9];
    memset(S, 0, sizeof(S));
    size_t used = 0;
    for (size_t i = 0; i < np; i++)
    {
        double b[3], a[3]; /* b: local, a: global, centroid-relative */
        for (int d = 0; d < 3; d++) b[d] = P[3 * i + d] - pc[d], a[d] = Q[3 * i + d] - qc[d];
        const double bn = sqrt((b[0] * b[0] + b[1] * b[1]) + b[2] * b[2]);
        const double an = sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]);
        double w = 1.0;
        if (prm->use_scale_outlier_detector)
        { /* row M */
            const double mx = bn > an ? bn : an, mn = bn > an ? an : bn;
            if (!(mn > 0.0) || mx / mn > prm->scale_outlier_threshold) continue;
        }
        if (prm->use_robust_kernel && bn > 0 && an > 0)
        {
            double bu[3] = {b[0] / bn, b[1] / bn, b[2] / bn}, rb[3];
            mat3_vec(Rprior, bu, rb);
            double cs = (rb[0] * a[0] + rb[1] * a[1] + rb[2] * a[2]) / an;
            if (cs > 1) cs = 1;
            if (cs < -1) cs = -1;
            const double ang = acos(cs);
            if (ang > prm->robust_kernel_param)
            {
                const double e = ang - prm->robust_kernel_param;
                w *= 1.0 / (1.0 + prm->robust_kernel_scale * e * e);
            }
        }
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) S[r * 3 + c] += w * b[r] * a[c];
        used++;
    }
    const double Sxx = S[0], Sxy = S[1], Sxz = S[2], Syx = S[3], Syy = S[4], Syz = S[5],
                 Szx = S[6], Szy = S[7], Szz = S[8];
    double N[16] = {Sxx + Syy + Szz, Syz - Szy,        Szx - Sxz,         Sxy - Syx,
                    Syz - Szy,       Sxx - Syy - Szz,  Sxy + Syx,         Szx + Sxz,
                    Szx - Sxz,       Sxy + Syx,        -Sxx + Syy - Szz,  Syz + Szy,
                    Sxy - Syx,       Szx + Sxz,        Syz + Szy,         -Sxx - Syy + Szz};
    double ev[4], V[16];
    jacobi_sym(4, N, ev, V);
    double qw = V[0 * 4 + 3], qx = V[1 * 4 + 3], qy = V[2 * 4 + 3], qz = V[3 * 4 + 3];
    const double qn = sqrt(((qw * qw + qx * qx) + qy * qy) + qz * qz);
    qw /= qn, qx /= qn, qy /= qn, qz /= qn;
    if (qw < 0) qw = -qw, qx = -qx, qy = -qy, qz = -qz;
    R[0] = 1 - 2 * (qy * qy + qz * qz), R[1] = 2 * (qx * qy - qw * qz), R[2] = 2 * (qx * qz + qw * qy);
    R[3] = 2 * (qx * qy + qw * qz), R[4] = 1 - 2 * (qx * qx + qz * qz), R[5] = 2 * (qy * qz - qw * qx);
    R[6] = 2 * (qx * qz - qw * qy), R[7] = 2 * (qy * qz + qw * qx), R[8] = 1 - 2 * (qx * qx + qy * qy);
    double Rp[3];
    mat3_vec(R, pc, Rp);
    for (int d = 0; d < 3; d++) t[d] = qc[d] - Rp[d];
    return used;
}

/* ------------------------------------------------------------- covariance */
/* A.9: forward-difference Jacobian of the stacked residuals wrt
 * (x,y,z,yaw,pitch,roll); H = J^T J ; cov = H^-1 */
static void covariance_fd(int pt2pl, const double* P, const double* Cc, const double* Nn,
                          size_t np, const double R[9], const double t[3], double h,
                          double cov[36], uint32_t* singular)
{
    double x0[6];
    orc_Rt_to_pose(R, t, x0);
    double R0[9], t0[3];
    orc_pose_to_Rt(x0, R0, t0);
    double Rj[6][9], tj[6][3];
    for (int j = 0; j < 6; j++)
    {
        double xj[6];
        memcpy(xj, x0, sizeof(xj));
        xj[j] += h;
        orc_pose_to_Rt(xj, Rj[j], tj[j]);
    }
    double H[36];
    memset(H, 0, sizeof(H));
    for (size_t i = 0; i < np; i++)
    {
        const double* p = P + 3 * i;
        double base[3];
        mat3_vec(R0, p, base);
        for (int d = 0; d < 3; d++) base[d] += t0[d];
        if (pt2pl)
        {
            const double* c = Cc + 3 * i;
            const double* n = Nn + 3 * i;
            const double r0 =
                (n[0] * (base[0] - c[0]) + n[1] * (base[1] - c[1])) + n[2] * (base[2] - c[2]);
            double J[6];
            for (int j = 0; j < 6; j++)
            {
                double v[3];
                mat3_vec(Rj[j], p, v);
                for (int d = 0; d < 3; d++) v[d] += tj[j][d];
                const double rj =
                    (n[0] * (v[0] - c[0]) + n[1] * (v[1] - c[1])) + n[2] * (v[2] - c[2]);
                J[j] = (rj - r0) / h;
            }
            for (int a = 0; a < 6; a++)
                for (int b = 0; b < 6; b++) H[a * 6 + b] += J[a] * J[b];
        }
        else
        {
            double J[3][6];
            for (int j = 0; j < 6; j++)
            {
                double v[3];
                mat3_vec(Rj[j], p, v);
                for (int d = 0; d < 3; d++) J[d][j] = ((v[d] + tj[j][d]) - base[d]) / h;
            }
            for (int a = 0; a < 6; a++)
                for (int b = 0; b < 6; b++)
                    H[a * 6 + b] += (J[0][a] * J[0][b] + J[1][a] * J[1][b]) + J[2][a] * J[2][b];
        }
    }
    const int rank = orc_inverse6(H, cov);
    *singular = 0;
    if (rank < 6)
    {
        memset(cov, 0, sizeof(double) * 36);
        *singular = 1;
    }
}

/* ------------------------------------------------------------ ICP (row G) */
int orc_icp_align(orc_cloud* from_global, const orc_cloud* to_local, const double guess[6],
                  const orc_params* p, int use_kdtree, orc_result* out)
{
    memset(out, 0, sizeof(*out));
    const size_t nl = to_local->n;
    uint32_t k = p->knn ? p->knn : 1;
    if (k > ORC_MAX_K) k = ORC_MAX_K;
    double R[9], t[3];
    orc_pose_to_Rt(guess, R, t);
    double Rprev[9], tprev[3];
    memcpy(Rprev, R, sizeof(R));
    memcpy(tprev, t, sizeof(t));

    uint8_t*  paired = (uint8_t*)malloc(nl ? nl : 1);
    double*   cen = (double*)malloc(sizeof(double) * 3 * (nl ? nl : 1));
    double*   nor = (double*)malloc(sizeof(double) * 3 * (nl ? nl : 1));
    uint32_t* nn1 = (uint32_t*)malloc(sizeof(uint32_t) * (nl ? nl : 1));
    double*   P = (double*)malloc(sizeof(double) * 3 * (nl ? nl : 1));
    double*   A = (double*)malloc(sizeof(double) * 3 * (nl ? nl : 1));
    double*   B = (double*)malloc(sizeof(double) * 3 * (nl ? nl : 1));
    size_t    np = 0;
    const int pt2pl = (p->matcher_kind == ORC_MATCHER_POINT2PLANE);

    out->termination_reason = ORC_TERM_UNDEFINED;
    uint32_t it = 0;
    for (it = 0; it < p->max_iterations; it++)
    {
        /* matcher gating (A.5) */
        const int active = (p->run_from_iteration <= it) &&
                           (p->run_up_to_iteration == 0 || it <= p->run_up_to_iteration);
        np = 0;
        if (active && nl && from_global->n)
        {
            if (pt2pl)
            {
                orc_match_point2plane(from_global, to_local, R, t, p, use_kdtree, paired, NULL,
                                      NULL, cen, nor);
                for (size_t i = 0; i < nl; i++)
                    if (paired[i])
                    {
                        P[3 * np] = to_local->x[i], P[3 * np + 1] = to_local->y[i],
                                P[3 * np + 2] = to_local->z[i];
                        memcpy(A + 3 * np, cen + 3 * i, 3 * sizeof(double));
                        memcpy(B + 3 * np, nor + 3 * i, 3 * sizeof(double));
                        np++;
                    }
            }
            else
            {
                orc_match_points(from_global, to_local, R, t, p->distance_threshold, use_kdtree,
                                 nn1, NULL);
                for (size_t i = 0; i < nl; i++)
                    if (nn1[i] != ORC_INVALID_IDX)
                    {
                        P[3 * np] = to_local->x[i], P[3 * np + 1] = to_local->y[i],
                                P[3 * np + 2] = to_local->z[i];
                        A[3 * np] = from_global->x[nn1[i]], A[3 * np + 1] = from_global->y[nn1[i]],
                                A[3 * np + 2] = from_global->z[nn1[i]];
                        np++;
                    }
            }
        }
        out->n_pairings = (uint32_t)np;
        if (!np)
        {
            out->termination_reason = ORC_TERM_NO_PAIRINGS;
            break;
        }
        /* solver */
        if (p->solver_kind == ORC_SOLVER_GAUSS_NEWTON)
        {
            if (pt2pl)
                orc_gn_point2plane(P, A, B, np, p->solver_max_iterations, p->gn_min_delta, R, t);
            else
                orc_gn_point2point(P, A, np, p->solver_max_iterations, p->gn_min_delta, R, t);
        }
        else
        {
            if (pt2pl)
            { /* Horn consumes point-to-point pairs only: use plane centroids */
                double Rn[9], tn[3];
                if (orc_horn(P, A, np, p, R, Rn, tn) < 3)
                {
                    out->termination_reason = ORC_TERM_SOLVER_ERROR;
                    break;
                }
                memcpy(R, Rn, sizeof(Rn)), memcpy(t, tn, sizeof(tn));
            }
            else
            {
                double Rn[9], tn[3];
                if (orc_horn(P, A, np, p, R, Rn, tn) < 3)
                {
                    out->termination_reason = ORC_TERM_SOLVER_ERROR;
                    break;
                }
                memcpy(R, Rn, sizeof(Rn)), memcpy(t, tn, sizeof(tn));
            }
        }
        /* convergence: delta = log(prev^-1 * new), split (xyz, rot) */
        double dR[9], dt[3], d[6];
        orc_se3_inverse_compose(Rprev, tprev, R, t, dR, dt);
        orc_se3_log(dR, dt, d);
        const double dxyz = sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
        const double drot = sqrt((d[3] * d[3] + d[4] * d[4]) + d[5] * d[5]);
        if (dxyz < p->min_abs_step_trans && drot < p->min_abs_step_rot)
        {
            out->termination_reason = ORC_TERM_STALLED;
            break;
        }
        memcpy(Rprev, R, sizeof(R));
        memcpy(tprev, t, sizeof(t));
    }
    out->n_iterations = it;
    if (it >= p->max_iterations) out->termination_reason = ORC_TERM_MAX_ITERATIONS;

    memcpy(out->R, R, sizeof(R));
    memcpy(out->t, t, sizeof(t));
    orc_Rt_to_pose(R, t, out->pose);
    out->quality = orc_quality_paired_ratio(from_global, to_local, R, t,
                                            p->quality_threshold_distance, use_kdtree);
    if (np)
        covariance_fd(pt2pl, P, A, B, np, R, t, p->cov_fd_step, out->cov, &out->cov_singular);
    else
        out->cov_singular = 1;

    free(paired), free(cen), free(nor), free(nn1), free(P), free(A), free(B);
    return 0;
}

/* ------------------------------------------------------- voxel decimation */
typedef struct
{
    int32_t  kx, ky, kz;
    uint32_t first; /* lowest original index, INVALID = empty slot */
    uint32_t count;
    double   sx, sy, sz;
} vox_slot;

static inline uint64_t vox_hash(int32_t a, int32_t b, int32_t c)
{
    uint64_t h = (uint64_t)(uint32_t)a * 0x9E3779B97F4A7C15ull;
    h ^= (uint64_t)(uint32_t)b * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= (uint64_t)(uint32_t)c * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
    return h;
}

size_t orc_voxel_decimate(const float* x, const float* y, const float* z, size_t n,
                          float resolution, int use_average, uint32_t* keep_idx, float* ox,
                          float* oy, float* oz)
{
    if (!n) return 0;
    size_t cap = 16;
    while (cap < 2 * n) cap <<= 1;
    vox_slot* tab = (vox_slot*)malloc(sizeof(vox_slot) * cap);
    for (size_t i = 0; i < cap; i++) tab[i].first = ORC_INVALID_IDX;
    uint32_t* slot_of = (uint32_t*)malloc(sizeof(uint32_t) * n);
    size_t m = 0;
    for (size_t i = 0; i < n; i++)
    {
        if (!(isfinite(x[i]) && isfinite(y[i]) && isfinite(z[i])))
        {
            slot_of[i] = ORC_INVALID_IDX;
            continue;
        }
        /* A.11: f32 division then floor */
        const int32_t kx = (int32_t)floorf(x[i] / resolution);
        const int32_t ky = (int32_t)floorf(y[i] / resolution);
        const int32_t kz = (int32_t)floorf(z[i] / resolution);
        size_t s = (size_t)(vox_hash(kx, ky, kz) & (cap - 1));
        for (;;)
        {
            vox_slot* v = &tab[s];
            if (v->first == ORC_INVALID_IDX)
            {
                v->kx = kx, v->ky = ky, v->kz = kz;
                v->first = (uint32_t)i, v->count = 0;
                v->sx = v->sy = v->sz = 0;
                keep_idx[m++] = (uint32_t)i; /* first-seen == lowest index, ascending */
                break;
            }
            if (v->kx == kx && v->ky == ky && v->kz == kz) break;
            s = (s + 1) & (cap - 1);
        }
        tab[s].count++;
        tab[s].sx += (double)x[i], tab[s].sy += (double)y[i], tab[s].sz += (double)z[i];
        slot_of[i] = (uint32_t)s;
    }
    for (size_t j = 0; j < m; j++)
    {
        const uint32_t i = keep_idx[j];
        if (use_average)
        {
            const vox_slot* v = &tab[slot_of[i]];
            if (ox) ox[j] = (float)(v->sx / (double)v->count);
            if (oy) oy[j] = (float)(v->sy / (double)v->count);
            if (oz) oz[j] = (float)(v->sz / (double)v->count);
        }
        else
        {
            if (ox) ox[j] = x[i];
            if (oy) oy[j] = y[i];
            if (oz) oz[j] = z[i];
        }
    }
    free(tab), free(slot_of);
    return m;
}

/* ----------------------------------------------------------------- A.13 */
/* FilterEdgesPlanes (the class the reference's stale keys name,
 * params/kitti-default.yaml:21-32, LidarOdometry.h:76-80).  The filter body
 * lives in mp2p_icp_filters (absent, unpinned): this is a restatement of its
 * published behaviour, NORMATIVE for this repo where upstream is uncertain:
 *   voxel key as A.11 (f32 division, floor); per voxel with >= min_points
 *   points: mean and covariance (1/n) in f64, points taken in ascending
 *   original index, eigenvalues e0 <= e1 <= e2 and the eigenvector v0 of e0
 *   by the cyclic Jacobi of row J;
 *     e2 < max_e2_e0 * e0 && e1 < max_e1_e0 * e0            -> "edges"
 *     else e2 > min_e2_e0 * e0 && e1 > min_e1_e0 * e0
 *          && |v0.z| < 0.9 (ground-like planes are dropped)  -> "planes"
 *   every voxel_decimation-th point of a classified voxel (t = 0, d, 2d, ...
 *   in ascending original index) joins the class layer; every
 *   full_decimation-th point of EVERY voxel joins "full_decim".
 * layer[i]: bit 0 edges, bit 1 planes, bit 2 full_decim.  Returns the number of
 * classified voxels. */
typedef struct
{
    uint64_t key;
    uint32_t idx;
} ep_rec;
static int ep_cmp(const void* a, const void* b)
{
    const ep_rec *x = (const ep_rec*)a, *y = (const ep_rec*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}
size_t orc_filter_edges_planes(const float* x, const float* y, const float* z, size_t n, float resolution,
                               uint32_t full_decimation, uint32_t voxel_decimation, float max_e2_e0,
                               float max_e1_e0, float min_e2_e0, float min_e1_e0, uint32_t min_points,
                               uint8_t* layer)
{
    if (!n) return 0;
    if (full_decimation < 1) full_decimation = 1;
    if (voxel_decimation < 1) voxel_decimation = 1;
    ep_rec* rec = (ep_rec*)malloc(sizeof(ep_rec) * n);
    size_t  m = 0;
    for (size_t i = 0; i < n; i++)
    {
        layer[i] = 0;
        if (!(isfinite(x[i]) && isfinite(y[i]) && isfinite(z[i]))) continue;
        const float fx = floorf(x[i] / resolution), fy = floorf(y[i] / resolution), fz = floorf(z[i] / resolution);
        if (!(fabsf(fx) <= 1048575.0f) || !(fabsf(fy) <= 1048575.0f) || !(fabsf(fz) <= 1048575.0f)) continue;
        rec[m].key = (uint64_t)((int32_t)fx + 1048576) | ((uint64_t)((int32_t)fy + 1048576) << 21) |
                     ((uint64_t)((int32_t)fz + 1048576) << 42);
        rec[m].idx = (uint32_t)i;
        m++;
    }
    qsort(rec, m, sizeof(ep_rec), ep_cmp);
    size_t classified = 0;
    for (size_t a = 0; a < m;)
    {
        size_t b = a;
        while (b < m && rec[b].key == rec[a].key) b++;
        const size_t cnt = b - a;
        uint8_t      cls = 0;
        if (cnt >= min_points && cnt >= 1)
        {
            double sx = 0, sy = 0, sz = 0;
            for (size_t j = a; j < b; j++)
                sx += (double)x[rec[j].idx], sy += (double)y[rec[j].idx], sz += (double)z[rec[j].idx];
            const double inv = 1.0 / (double)cnt;
            const double cx = sx * inv, cy = sy * inv, cz = sz * inv;
            double c00 = 0, c01 = 0, c02 = 0, c11 = 0, c12 = 0, c22 = 0;
            for (size_t j = a; j < b; j++)
            {
                const double dx = (double)x[rec[j].idx] - cx, dy = (double)y[rec[j].idx] - cy,
                             dz = (double)z[rec[j].idx] - cz;
                c00 += dx * dx, c01 += dx * dy, c02 += dx * dz;
                c11 += dy * dy, c12 += dy * dz, c22 += dz * dz;
            }
            double C[9] = {c00 * inv, c01 * inv, c02 * inv, c01 * inv, c11 * inv, c12 * inv, c02 * inv, c12 * inv, c22 * inv};
            double ev[3], V[9];
            jacobi_sym(3, C, ev, V);
            const double e0 = ev[0], e1 = ev[1], e2 = ev[2];
            if (e2 < (double)max_e2_e0 * e0 && e1 < (double)max_e1_e0 * e0)
                cls = 1;
            else if (e2 > (double)min_e2_e0 * e0 && e1 > (double)min_e1_e0 * e0 && fabs(V[6]) < 0.9)
                cls = 2; /* V[6] = z component of the first column (eigenvector of e0) */
            if (cls) classified++;
        }
        for (size_t j = a; j < b; j++)
        {
            const size_t t = j - a;
            uint8_t      f = 0;
            if (cls && (t % voxel_decimation) == 0) f |= cls;
            if ((t % full_decimation) == 0) f |= 4;
            layer[rec[j].idx] = f;
        }
        a = b;
    }
    free(rec);
    return classified;
}
