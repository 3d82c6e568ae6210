// edges_planes.cu -- FilterEdgesPlanes on the device: the point-cloud filter the
// reference's parameter files name (params/kitti-default.yaml:21-32
// `pointcloud_filter_class: mola::lidar_segmentation::FilterEdgesPlanes`, defaults
// in include/mola-fe-lidar/LidarOdometry.h:76-80; applied through
// apply_filter_pipeline, LidarOdometry.cpp:223-224).  SURVEY.md 8f rank 3; the
// algorithm is the spec the CPU checker restates as A.13.
//
//   voxel key per point (A.11) -> radix sort of (key, original index): a voxel
//   is a run of the sorted array, its points in ascending original index ->
//   one thread per voxel: mean, covariance (f64, that order), cyclic Jacobi,
//   the eigen-ratio gates -> per-point layer flags -> ONE 64-bit exclusive scan
//   numbers the three output layers at once (21-bit fields) -> compaction in
//   ascending original index -> three indexed clouds ("edges", "planes",
//   "full_decim").
// The per-voxel sums run sequentially in one thread on purpose: the gates
// compare eigenvalue RATIOS with thresholds, and the layer flags are checked
// bit for bit against the oracle, which sums in the same order.  A voxel holds
// tens to a few hundred points at the shipped 1 m resolution; a cloud that falls
// into ONE voxel makes that thread walk all of it (correct, just slow).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "icp_math.cuh"
#include "runtime.cuh"

namespace b2
{
constexpr unsigned long long kNoVoxel = 0xFFFFFFFFFFFFFFFFull;

struct EdgesPlanesDev
{
    float    res;
    uint32_t full_decim, vox_decim;
    float    max_e2_e0, max_e1_e0, min_e2_e0, min_e1_e0;
    uint32_t min_points;
};

__global__ void ep_key_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z,
                              uint32_t n, float res, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float        px = x[i], py = y[i], pz = z[i];
    unsigned long long key = kNoVoxel;
    if (isfinite(px) && isfinite(py) && isfinite(pz))
    {
        // A.11: f32 division, then floor; 21 bits per axis
        const float fx = floorf(px / res), fy = floorf(py / res), fz = floorf(pz / res);
        const float lim = 1048575.0f;
        if ((fabsf(fx) <= lim) && (fabsf(fy) <= lim) && (fabsf(fz) <= lim))
            key = (unsigned long long)((int)fx + 1048576) | ((unsigned long long)((int)fy + 1048576) << 21) |
                  ((unsigned long long)((int)fz + 1048576) << 42);
    }
    keys[i] = key;
    vals[i] = i;
}

// thread j works when sorted position j is the first point of a voxel
__global__ void ep_classify_kernel(const unsigned long long* __restrict__ skeys, const uint32_t* __restrict__ svals,
                                   uint32_t n, const float* __restrict__ x, const float* __restrict__ y,
                                   const float* __restrict__ z, EdgesPlanesDev P, unsigned long long* __restrict__ flags,
                                   uint32_t* __restrict__ n_classified)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const unsigned long long key = skeys[j];
    if (key == kNoVoxel)
    {
        flags[svals[j]] = 0ull;  // not in any voxel: in no layer
        return;
    }
    if (j != 0 && skeys[j - 1] == key) return;
    uint32_t end = j + 1;
    while (end < n && skeys[end] == key) end++;
    const uint32_t cnt = end - j;
    uint32_t       cls = 0;
    if (cnt >= P.min_points)
    {
        double sx = 0, sy = 0, sz = 0;
        for (uint32_t t = j; t < end; t++)
        {
            const uint32_t i = svals[t];
            sx += (double)x[i], sy += (double)y[i], sz += (double)z[i];
        }
        const double inv = 1.0 / (double)cnt;
        const double cx = sx * inv, cy = sy * inv, cz = sz * inv;
        double       c00 = 0, c01 = 0, c02 = 0, c11 = 0, c12 = 0, c22 = 0;
        for (uint32_t t = j; t < end; t++)
        {
            const uint32_t i = svals[t];
            const double   dx = (double)x[i] - cx, dy = (double)y[i] - cy, dz = (double)z[i] - cz;
            c00 += dx * dx, c01 += dx * dy, c02 += dx * dz;
            c11 += dy * dy, c12 += dy * dz, c22 += dz * dz;
        }
        double C[9] = {c00 * inv, c01 * inv, c02 * inv, c01 * inv, c11 * inv, c12 * inv, c02 * inv, c12 * inv, c22 * inv};
        double ev[3], V[9];
        jacobi3(C, ev, V);
        const double e0 = ev[0], e1 = ev[1], e2 = ev[2];
        if (e2 < (double)P.max_e2_e0 * e0 && e1 < (double)P.max_e1_e0 * e0)
            cls = 1;
        else if (e2 > (double)P.min_e2_e0 * e0 && e1 > (double)P.min_e1_e0 * e0 && fabs(V[6]) < 0.9)
            cls = 2;
        if (cls) atomicAdd(n_classified, 1u);
    }
    for (uint32_t t = j; t < end; t++)
    {
        const uint32_t u = t - j;
        uint32_t       f = 0;
        if (cls && (u % P.vox_decim) == 0) f |= cls;
        if ((u % P.full_decim) == 0) f |= 4u;
        // one counter field per layer: the scan numbers all three outputs
        flags[svals[t]] = (unsigned long long)(f & 1u) | ((unsigned long long)((f >> 1) & 1u) << 21) |
                          ((unsigned long long)((f >> 2) & 1u) << 42);
    }
}

struct EpOut
{
    float*    x[3];
    float*    y[3];
    float*    z[3];
    uint32_t* keep[3];
};

__global__ void ep_compact_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z,
                                  uint32_t n, const unsigned long long* __restrict__ flags,
                                  const unsigned long long* __restrict__ pos, EpOut o, uint8_t* __restrict__ layer,
                                  uint32_t* __restrict__ totals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long f = flags[i], p = pos[i];
    if (i == n - 1)
    {
        const unsigned long long t = p + f;
#pragma unroll
        for (int l = 0; l < 3; l++) totals[l] = (uint32_t)((t >> (21 * l)) & 0x1FFFFFull);
    }
    uint32_t bits = 0;
#pragma unroll
    for (int l = 0; l < 3; l++)
        if ((f >> (21 * l)) & 1ull)
        {
            const uint32_t d = (uint32_t)((p >> (21 * l)) & 0x1FFFFFull);
            o.x[l][d] = x[i], o.y[l][d] = y[i], o.z[l][d] = z[i];
            o.keep[l][d] = i;
            bits |= 1u << l;
        }
    if (layer) layer[i] = (uint8_t)bits;
}

int run_edges_planes(::b200icp* ctx, const b200icp_cloud* in, const b200icp_edges_planes_params_t* prm,
                     float search_radius, b200icp_cloud** out3, uint8_t* layer_out, uint32_t* n_voxels_out)
{
    if (!(prm->voxel_filter_resolution > 0) || !std::isfinite(prm->voxel_filter_resolution))
    {
        set_error("voxel_filter_resolution must be positive and finite");
        return B200ICP_ERR_BAD_ARG;
    }
    if (in->n >= (1u << 21))
    {
        set_error("FilterEdgesPlanes: at most %u points per cloud", (1u << 21) - 1);
        return B200ICP_ERR_BAD_ARG;
    }
    Lease L(ctx);
    if (!L.ws) return B200ICP_ERR_CUDA;
    Workspace*     ws = L.ws;
    cudaStream_t   s = ws->stream;
    const uint32_t n = (uint32_t)in->n;
    B2_CUDA_TRY(cudaStreamWaitEvent(s, in->ready, 0));
    EdgesPlanesDev P;
    P.res = prm->voxel_filter_resolution;
    P.full_decim = prm->full_pointcloud_decimation ? prm->full_pointcloud_decimation : 1u;
    P.vox_decim = prm->voxel_filter_decimation ? prm->voxel_filter_decimation : 1u;
    P.max_e2_e0 = prm->voxel_filter_max_e2_e0, P.max_e1_e0 = prm->voxel_filter_max_e1_e0;
    P.min_e2_e0 = prm->voxel_filter_min_e2_e0, P.min_e1_e0 = prm->voxel_filter_min_e1_e0;
    P.min_points = prm->min_points_per_voxel;
    uint32_t totals[4] = {0, 0, 0, 0};
    EpOut    o = {};
    if (n)
    {
        size_t sort_bytes = 0, scan_bytes = 0;
        B2_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (unsigned long long*)nullptr,
                                                    (unsigned long long*)nullptr, (uint32_t*)nullptr,
                                                    (uint32_t*)nullptr, (int)n, 0, 64, s));
        B2_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (unsigned long long*)nullptr,
                                                  (unsigned long long*)nullptr, (int)n, s));
        unsigned long long *keys, *skeys, *flags, *pos;
        uint32_t *          vals, *svals, *misc;
        uint8_t*            d_layer;
        void*               tmp;
        auto                layout = [&](Carver& c) {
            keys = c.take<unsigned long long>(n), skeys = c.take<unsigned long long>(n);
            flags = c.take<unsigned long long>(n), pos = c.take<unsigned long long>(n);
            vals = c.take<uint32_t>(n), svals = c.take<uint32_t>(n);
            for (int l = 0; l < 3; l++)
                o.x[l] = c.take<float>(n), o.y[l] = c.take<float>(n), o.z[l] = c.take<float>(n),
                o.keep[l] = c.take<uint32_t>(n);
            d_layer = c.take<uint8_t>(n);
            misc = c.take<uint32_t>(4);
            tmp = c.take<char>(std::max(sort_bytes, scan_bytes));
        };
        Carver sz(nullptr);
        layout(sz);
        if (int r = ws->reserve_device(sz.off)) return r;
        Carver real(ws->d_scratch);
        layout(real);
        const int blocks = (int)((n + 255) / 256);
        B2_CUDA_TRY(cudaMemsetAsync(misc, 0, 4 * sizeof(uint32_t), s));
        ep_key_kernel<<<blocks, 256, 0, s>>>(in->dx, in->dy, in->dz, n, P.res, keys, vals);
        B2_CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, sort_bytes, keys, skeys, vals, svals, (int)n, 0, 64, s));
        ep_classify_kernel<<<blocks, 256, 0, s>>>(skeys, svals, n, in->dx, in->dy, in->dz, P, flags, misc + 3);
        B2_CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, scan_bytes, flags, pos, (int)n, s));
        ep_compact_kernel<<<blocks, 256, 0, s>>>(in->dx, in->dy, in->dz, n, flags, pos, o, layer_out ? d_layer : nullptr,
                                                 misc);
        ws->launches += 5;
        B2_CUDA_TRY(cudaGetLastError());
        if (int r = ws->reserve_pinned(64)) return r;
        uint32_t* h = (uint32_t*)ws->h_pinned;
        B2_CUDA_TRY(cudaMemcpyAsync(h, misc, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        if (layer_out) B2_CUDA_TRY(cudaMemcpyAsync(layer_out, d_layer, n, cudaMemcpyDeviceToHost, s));
        B2_CUDA_TRY(cudaStreamSynchronize(s));
        for (int l = 0; l < 4; l++) totals[l] = h[l];
    }
    if (n_voxels_out) *n_voxels_out = totals[3];
    // the three layers as clouds; the scratch holding the compacted coordinates is reused by the index builds, so
    // all copies are enqueued first (same stream: ordered before the builds)
    b200icp_cloud* c[3] = {nullptr, nullptr, nullptr};
    auto           drop = [&]() {
        for (auto* p : c)
            if (p) b200icp_cloud_free(p);
    };
    for (int l = 0; l < 3; l++)
    {
        // classified layers keep every voxel_decimation-th point of a voxel: spacing like the voxel size
        if (int r = cloud_alloc(ctx, ws, totals[l], search_radius, &c[l], 0.0f))
        {
            drop();
            return r;
        }
        const size_t m = totals[l];
        if (!m) continue;
        cudaError_t e = cudaMemcpyAsync(c[l]->dx, o.x[l], m * sizeof(float), cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c[l]->dy, o.y[l], m * sizeof(float), cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c[l]->dz, o.z[l], m * sizeof(float), cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess)
        {
            set_error("FilterEdgesPlanes output copy failed: %s", cudaGetErrorString(e));
            drop();
            return B200ICP_ERR_CUDA;
        }
    }
    for (int l = 0; l < 3; l++)
        if (int r = cloud_build_index(ctx, ws, c[l]))
        {
            drop();
            return r;
        }
    for (int l = 0; l < 3; l++) out3[l] = c[l];
    return B200ICP_OK;
}

}  // namespace b2
