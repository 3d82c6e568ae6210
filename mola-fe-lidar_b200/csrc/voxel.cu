// voxel.cu -- voxel-grid decimation of an incoming scan on the device
// (the reference's apply_filter_pipeline -> FilterDecimateVoxels step,
// LidarOdometry.cpp:223-224).
//
// One point per occupied voxel = the LOWEST original index in it (or the voxel
// mean), output ordered by ascending original index.  A GPU hash over voxel
// keys instead of a sort: insert (atomicCAS claim + atomicMin of the index),
// flag the winners, exclusive scan, compact.  12 B read per input point,
// 12 B written per output point.
#include <cub/device/device_scan.cuh>

#include "runtime.cuh"

namespace b2
{
constexpr unsigned long long kEmptyVox = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ bool vox_key(float x, float y, float z, float res, unsigned long long& key)
{
    // A.11: f32 division, then floor
    const float fx = floorf(x / res), fy = floorf(y / res), fz = floorf(z / res);
    const float lim = 1048575.0f;  // 2^20 - 1
    if (!(fabsf(fx) <= lim) || !(fabsf(fy) <= lim) || !(fabsf(fz) <= lim)) return false;
    const unsigned long long kx = (unsigned long long)((int)fx + 1048576);
    const unsigned long long ky = (unsigned long long)((int)fy + 1048576);
    const unsigned long long kz = (unsigned long long)((int)fz + 1048576);
    key = kx | (ky << 21) | (kz << 42);
    return true;
}

__device__ __forceinline__ uint32_t vox_hash(unsigned long long k, uint32_t shift)
{
    return (uint32_t)((k * 0x9E3779B97F4A7C15ull) >> shift);
}

__global__ void voxel_insert_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                    const float* __restrict__ z, uint32_t n, float res,
                                    unsigned long long* __restrict__ keys,
                                    uint32_t* __restrict__ minidx, double* __restrict__ sums,
                                    uint32_t* __restrict__ counts, uint32_t* __restrict__ slot_of,
                                    uint32_t shift, uint32_t mask, int use_average,
                                    uint32_t* __restrict__ overflow)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float px = x[i], py = y[i], pz = z[i];
    uint32_t    my_slot = kInvalid;
    if (isfinite(px) && isfinite(py) && isfinite(pz))
    {
        unsigned long long key;
        if (!vox_key(px, py, pz, res, key))
            atomicExch(overflow, 1u);
        else
        {
            uint32_t slot = vox_hash(key, shift);
            for (;;)
            {
                const unsigned long long prev = atomicCAS(keys + slot, kEmptyVox, key);
                if (prev == kEmptyVox || prev == key) break;
                slot = (slot + 1) & mask;
            }
            atomicMin(minidx + slot, i);
            if (use_average)
            {
                atomicAdd(sums + 3 * (size_t)slot + 0, (double)px);
                atomicAdd(sums + 3 * (size_t)slot + 1, (double)py);
                atomicAdd(sums + 3 * (size_t)slot + 2, (double)pz);
                atomicAdd(counts + slot, 1u);
            }
            my_slot = slot;
        }
    }
    slot_of[i] = my_slot;
}

__global__ void voxel_flag_kernel(const uint32_t* __restrict__ slot_of,
                                  const uint32_t* __restrict__ minidx, uint32_t n,
                                  uint32_t* __restrict__ flag)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = slot_of[i];
    flag[i] = (s != kInvalid && minidx[s] == i) ? 1u : 0u;
}

__global__ void voxel_compact_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                     const float* __restrict__ z, uint32_t n,
                                     const uint32_t* __restrict__ flag,
                                     const uint32_t* __restrict__ pos,
                                     const uint32_t* __restrict__ slot_of,
                                     const double* __restrict__ sums,
                                     const uint32_t* __restrict__ counts, int use_average,
                                     float* __restrict__ ox, float* __restrict__ oy,
                                     float* __restrict__ oz, uint32_t* __restrict__ keep,
                                     uint32_t* __restrict__ total)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == n - 1) *total = pos[i] + flag[i];
    if (!flag[i]) return;
    const uint32_t o = pos[i];
    if (use_average)
    {
        const uint32_t s = slot_of[i];
        const double   c = (double)counts[s];
        ox[o] = (float)(sums[3 * (size_t)s + 0] / c);
        oy[o] = (float)(sums[3 * (size_t)s + 1] / c);
        oz[o] = (float)(sums[3 * (size_t)s + 2] / c);
    }
    else
        ox[o] = x[i], oy[o] = y[i], oz[o] = z[i];
    keep[o] = i;
}

int run_voxel(::b200icp* ctx, const b200icp_cloud* in, float resolution, int use_average,
              float search_radius, b200icp_cloud** out, uint32_t* keep_idx)
{
    if (!(resolution > 0) || !std::isfinite(resolution))
    {
        set_error("voxel resolution must be positive and finite");
        return B200ICP_ERR_BAD_ARG;
    }
    Lease L(ctx);
    if (!L.ws) return B200ICP_ERR_CUDA;
    Workspace*     ws = L.ws;
    cudaStream_t   s = ws->stream;
    const uint32_t n = (uint32_t)in->n;
    B2_CUDA_TRY(cudaStreamWaitEvent(s, in->ready, 0));
    uint32_t m = 0;
    float *  ox = nullptr, *oy = nullptr, *oz = nullptr;
    uint32_t* keep = nullptr;
    const bool prof = ctx->profile_on;
    if (n)
    {
        uint32_t cap = 1024;
        while (cap < 2 * (size_t)n) cap <<= 1;
        uint32_t lg = 0;
        while ((1u << lg) < cap) lg++;
        size_t scan_bytes = 0;
        B2_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (uint32_t*)nullptr,
                                                  (uint32_t*)nullptr, (int)n, s));
        unsigned long long* keys;
        uint32_t *minidx, *counts, *slot_of, *flag, *pos, *misc;
        double*   sums;
        void*     scan_tmp;
        auto layout = [&](Carver& c) {
            keys = c.take<unsigned long long>(cap);
            minidx = c.take<uint32_t>(cap);
            counts = c.take<uint32_t>(use_average ? cap : 1);
            sums = c.take<double>(use_average ? 3 * (size_t)cap : 1);
            slot_of = c.take<uint32_t>(n);
            flag = c.take<uint32_t>(n);
            pos = c.take<uint32_t>(n);
            ox = c.take<float>(n), oy = c.take<float>(n), oz = c.take<float>(n);
            keep = c.take<uint32_t>(n);
            misc = c.take<uint32_t>(4);
            scan_tmp = c.take<char>(scan_bytes);
        };
        Carver sz(nullptr);
        layout(sz);
        if (int r = ws->reserve_device(sz.off)) return r;
        Carver real(ws->d_scratch);
        layout(real);
        if (prof)
        {
            if (int r = ws->reserve_prof_events(1)) return r;
            B2_CUDA_TRY(cudaEventRecord(ws->prof_ev[0], s));
        }
        B2_CUDA_TRY(cudaMemsetAsync(keys, 0xFF, (size_t)cap * sizeof(unsigned long long), s));
        B2_CUDA_TRY(cudaMemsetAsync(minidx, 0xFF, (size_t)cap * sizeof(uint32_t), s));
        B2_CUDA_TRY(cudaMemsetAsync(misc, 0, 4 * sizeof(uint32_t), s));
        if (use_average)
        {
            B2_CUDA_TRY(cudaMemsetAsync(counts, 0, (size_t)cap * sizeof(uint32_t), s));
            B2_CUDA_TRY(cudaMemsetAsync(sums, 0, 3 * (size_t)cap * sizeof(double), s));
        }
        const int blocks = (int)((n + 255) / 256);
        voxel_insert_kernel<<<blocks, 256, 0, s>>>(in->dx, in->dy, in->dz, n, resolution, keys, minidx,
                                                   sums, counts, slot_of, 64 - lg, cap - 1,
                                                   use_average, misc + 1);
        voxel_flag_kernel<<<blocks, 256, 0, s>>>(slot_of, minidx, n, flag);
        B2_CUDA_TRY(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, flag, pos, (int)n, s));
        voxel_compact_kernel<<<blocks, 256, 0, s>>>(in->dx, in->dy, in->dz, n, flag, pos, slot_of, sums,
                                                    counts, use_average, ox, oy, oz, keep, misc);
        ws->launches += 5;
        if (prof) B2_CUDA_TRY(cudaEventRecord(ws->prof_ev[1], s));
        B2_CUDA_TRY(cudaGetLastError());
        if (int r = ws->reserve_pinned(64)) return r;
        uint32_t* h = (uint32_t*)ws->h_pinned;
        B2_CUDA_TRY(cudaMemcpyAsync(h, misc, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        B2_CUDA_TRY(cudaStreamSynchronize(s));
        if (h[1])
        {
            set_error("voxel grid range exceeded: |coordinate / resolution| must stay below 2^20");
            return B200ICP_ERR_BAD_ARG;
        }
        m = h[0];
        if (prof)
        {
            float ms = 0;
            B2_CUDA_TRY(cudaEventElapsedTime(&ms, ws->prof_ev[0], ws->prof_ev[1]));
            std::lock_guard<std::mutex> lk(ctx->mtx);
            ctx->prof.voxel_launches++;
            ctx->prof.voxel_ms += ms;
            ctx->prof.voxel_points += n;
        }
    }
    b200icp_cloud* c = nullptr;
    // one point per voxel: neighbours are about a voxel apart, the index cells follow (never below the default)
    if (int r = cloud_alloc(ctx, ws, m, search_radius, &c, 0.7f * resolution)) return r;
    if (m)
    {
        cudaError_t e = cudaMemcpyAsync(c->dx, ox, m * sizeof(float), cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->dy, oy, m * sizeof(float), cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->dz, oz, m * sizeof(float), cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess && keep_idx)
        {
            e = cudaMemcpyAsync(keep_idx, keep, m * sizeof(uint32_t), cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        }
        if (e != cudaSuccess)
        {   // the new cloud must not outlive a failed copy
            set_error("voxel output copy failed: %s", cudaGetErrorString(e));
            b200icp_cloud_free(c);
            return B200ICP_ERR_CUDA;
        }
    }
    // the scratch holding ox/oy/oz is reused by the index build: the D2D copies
    // above are ordered before it on the same stream
    if (int r = cloud_build_index(ctx, ws, c))
    {
        b200icp_cloud_free(c);
        return r;
    }
    *out = c;
    return B200ICP_OK;
}

}  // namespace b2
