// cloud.cu -- device-resident clouds and the GPU-built search index.
//
// Replaces the lazily built nanoflann kd-tree on the `from` cloud
// (mrpt::math::KDTreeCapable) with a uniform grid:
//   bbox -> per-point 30-bit Morton cell key -> radix sort (key, index) ->
//   points gathered as float4 in cell order + inverse permutation ->
//   open-addressing hash of occupied cells -> [start,end).
// Everything after the H2D copy happens on the device with no host sync.
// Algorithmic traffic: 12 B read + 16 B float4 write + 4 B rank write per
// point, plus the sort passes (36 B/point).
#include <cub/device/device_radix_sort.cuh>

#include "runtime.cuh"

namespace b2
{
// ---- order-preserving float <-> uint encoding for atomic min/max ---------
__device__ __forceinline__ uint32_t enc_f(float f)
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_f(uint32_t e)
{
    const uint32_t u = (e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e;
    return __uint_as_float(u);
}

__device__ __forceinline__ uint32_t spread10(uint32_t v)
{
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ __forceinline__ uint32_t compact10(uint32_t v)
{
    v &= 0x09249249u;
    v = (v | (v >> 2)) & 0x030C30C3u;
    v = (v | (v >> 4)) & 0x0300F00Fu;
    v = (v | (v >> 8)) & 0x030000FFu;
    v = (v | (v >> 16)) & 0x3FFu;
    return v;
}

// bbox of the finite points: block reduce + 6 atomics per block
__global__ void bbox_kernel(const float* __restrict__ x, const float* __restrict__ y,
                            const float* __restrict__ z, uint32_t n, uint32_t* __restrict__ bb)
{
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float px = x[i], py = y[i], pz = z[i];
        if (isfinite(px) && isfinite(py) && isfinite(pz))
        {
            mn[0] = fminf(mn[0], px), mx[0] = fmaxf(mx[0], px);
            mn[1] = fminf(mn[1], py), mx[1] = fmaxf(mx[1], py);
            mn[2] = fminf(mn[2], pz), mx[2] = fmaxf(mx[2], pz);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; d++)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xFFFFFFFFu, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xFFFFFFFFu, mx[d], o));
        }
    }
    if ((threadIdx.x & 31) == 0)
    {
#pragma unroll
        for (int d = 0; d < 3; d++)
        {
            atomicMin(bb + d, enc_f(mn[d]));
            atomicMax(bb + 3 + d, enc_f(mx[d]));
        }
    }
}

__global__ void grid_setup_kernel(const uint32_t* __restrict__ bb, float cell_req, GridDev* g)
{
    float mn[3], mx[3];
    for (int d = 0; d < 3; d++) mn[d] = dec_f(bb[d]), mx[d] = dec_f(bb[3 + d]);
    const bool any = mn[0] <= mx[0];
    if (!any)
        for (int d = 0; d < 3; d++) mn[d] = mx[d] = 0.f;
    float ext = fmaxf(fmaxf(mx[0] - mn[0], mx[1] - mn[1]), mx[2] - mn[2]);
    // at most 1000 of the 1024 cells per axis may be spanned
    const float cell = fmaxf(cell_req, ext / 1000.0f);
    g->ox = mn[0], g->oy = mn[1], g->oz = mn[2];
    g->cell = cell;
    g->inv_cell = 1.0f / cell;
    g->slack = 6e-4f;  // > ulp(1024) = 1.2e-4 cells of rounding in (p - o) * inv_cell
    g->n_valid = 0;
    g->n_cells = 0;
    for (int d = 0; d < 3; d++) g->bmin[d] = mn[d], g->bmax[d] = mx[d];
}

__global__ void cell_key_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                const float* __restrict__ z, uint32_t n,
                                const GridDev* __restrict__ g, uint32_t* __restrict__ keys,
                                uint32_t* __restrict__ vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float px = x[i], py = y[i], pz = z[i];
    uint32_t key = 0xFFFFFFFFu;
    if (isfinite(px) && isfinite(py) && isfinite(pz))
    {
        const float inv = g->inv_cell;
        const int cx = min(max((int)floorf((px - g->ox) * inv), 0), kGridMax);
        const int cy = min(max((int)floorf((py - g->oy) * inv), 0), kGridMax);
        const int cz = min(max((int)floorf((pz - g->oz) * inv), 0), kGridMax);
        key = spread10(cx) | (spread10(cy) << 1) | (spread10(cz) << 2);
    }
    keys[i] = key;
    vals[i] = i;
}

// After the sort: gather float4 points in cell order, write the inverse
// permutation, and let the first point of every cell publish [start,end).
__global__ void reorder_kernel(const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ svals,
                               uint32_t n, const float* __restrict__ x, const float* __restrict__ y,
                               const float* __restrict__ z, float4* __restrict__ pts,
                               uint32_t* __restrict__ rank, uint32_t* __restrict__ hkeys,
                               uint2* __restrict__ hvals, uint32_t hshift, uint32_t hmask,
                               GridDev* __restrict__ g)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t key = skeys[j];
    const uint32_t i = svals[j];
    if (j == 0)
    {  // n_valid = first position holding the invalid key
        uint32_t lo = 0, hi = n;
        while (lo < hi)
        {
            const uint32_t mid = (lo + hi) >> 1;
            if (skeys[mid] < 0xFFFFFFFFu)
                lo = mid + 1;
            else
                hi = mid;
        }
        g->n_valid = lo;
    }
    if (key == 0xFFFFFFFFu)
    {
        rank[i] = kInvalid;
        return;
    }
    pts[j] = make_float4(x[i], y[i], z[i], __uint_as_float(i));
    rank[i] = j;
    if (j == 0 || skeys[j - 1] != key)
    {
        uint32_t lo = j + 1, hi = n;  // upper bound of this key
        while (lo < hi)
        {
            const uint32_t mid = (lo + hi) >> 1;
            if (skeys[mid] <= key)
                lo = mid + 1;
            else
                hi = mid;
        }
        const uint32_t ckey = compact10(key) | (compact10(key >> 1) << kGridBits) |
                              (compact10(key >> 2) << (2 * kGridBits));
        uint32_t slot = hash_slot(ckey, hshift);
        for (;;)
        {
            const uint32_t prev = atomicCAS(hkeys + slot, kEmptyKey, ckey);
            if (prev == kEmptyKey) break;
            slot = (slot + 1) & hmask;
        }
        hvals[slot] = make_uint2(j, lo);
        atomicAdd(&g->n_cells, 1u);
    }
}

int cloud_alloc(::b200icp* ctx, Workspace* ws, size_t n, float search_radius, b200icp_cloud** out)
{
    if (n >= 0x7FFFFFFFull)
    {
        set_error("cloud too large: %zu points", n);
        return B200ICP_ERR_BAD_ARG;
    }
    auto* c = new b200icp_cloud();
    c->ctx = ctx;
    c->n = n;
    float radius = search_radius > 0 ? search_radius : (float)ctx->P.distance_threshold;
    if (!(radius > 0) || !std::isfinite(radius)) radius = 1.0f;
    c->cell_req = radius * 1.002f;
    uint32_t cap = 1024;
    while (cap < 2 * n) cap <<= 1;
    c->hcap = cap;
    uint32_t lg = 0;
    while ((1u << lg) < cap) lg++;
    c->hshift = 32 - lg;
    const size_t nn = n ? n : 1;
    Carver cv(nullptr);
    auto layout = [&](Carver& k) {
        c->dx = k.take<float>(nn), c->dy = k.take<float>(nn), c->dz = k.take<float>(nn);
        c->pts = k.take<float4>(nn);
        c->rank = k.take<uint32_t>(nn);
        c->hkeys = k.take<uint32_t>(cap);
        c->hvals = k.take<uint2>(cap);
        c->grid = k.take<GridDev>(1);
        c->bbox_enc = k.take<uint32_t>(8);
    };
    layout(cv);
    cudaError_t e = cudaMallocAsync(&c->slab, cv.off, ws->stream);
    if (e != cudaSuccess)
    {
        set_error("cudaMallocAsync(%zu bytes) failed: %s", cv.off, cudaGetErrorString(e));
        delete c;
        return e == cudaErrorMemoryAllocation ? B200ICP_ERR_NOMEM : B200ICP_ERR_CUDA;
    }
    Carver real(c->slab);
    layout(real);
    e = cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming);
    if (e != cudaSuccess)
    {
        set_error("cudaEventCreate failed: %s", cudaGetErrorString(e));
        cudaFreeAsync(c->slab, ws->stream);
        delete c;
        return B200ICP_ERR_CUDA;
    }
    *out = c;
    return B200ICP_OK;
}

int cloud_build_index(::b200icp* ctx, Workspace* ws, b200icp_cloud* c)
{
    cudaStream_t   s = ws->stream;
    const uint32_t n = (uint32_t)c->n;
    const bool     prof = ctx->profile_on;
    if (prof)
    {
        if (int r = ws->reserve_prof_events(1)) return r;
        B2_CUDA_TRY(cudaEventRecord(ws->prof_ev[0], s));
    }
    // identities for the atomic min / max, empty hash
    B2_CUDA_TRY(cudaMemsetAsync(c->bbox_enc, 0xFF, 3 * sizeof(uint32_t), s));
    B2_CUDA_TRY(cudaMemsetAsync(c->bbox_enc + 3, 0x00, 3 * sizeof(uint32_t), s));
    B2_CUDA_TRY(cudaMemsetAsync(c->hkeys, 0xFF, (size_t)c->hcap * sizeof(uint32_t), s));
    if (n)
    {
        const int bb_blocks = (int)std::min<size_t>((n + 1023) / 1024, (size_t)ctx->sm_count * 4);
        bbox_kernel<<<bb_blocks, 256, 0, s>>>(c->dx, c->dy, c->dz, n, c->bbox_enc);
        ws->launches++;
    }
    grid_setup_kernel<<<1, 1, 0, s>>>(c->bbox_enc, c->cell_req, c->grid);
    ws->launches++;
    if (n)
    {
        size_t temp_bytes = 0;
        cub::DoubleBuffer<uint32_t> dk(nullptr, nullptr), dv(nullptr, nullptr);
        B2_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, dk, dv, (int)n, 0, 32, s));
        Carver cv(nullptr);
        cv.take<uint32_t>(n), cv.take<uint32_t>(n), cv.take<uint32_t>(n), cv.take<uint32_t>(n);
        cv.take<char>(temp_bytes);
        if (int r = ws->reserve_device(cv.off)) return r;
        Carver   k(ws->d_scratch);
        uint32_t* k0 = k.take<uint32_t>(n);
        uint32_t* k1 = k.take<uint32_t>(n);
        uint32_t* v0 = k.take<uint32_t>(n);
        uint32_t* v1 = k.take<uint32_t>(n);
        void*     temp = k.take<char>(temp_bytes);
        const int blocks = (int)((n + 255) / 256);
        cell_key_kernel<<<blocks, 256, 0, s>>>(c->dx, c->dy, c->dz, n, c->grid, k0, v0);
        ws->launches++;
        cub::DoubleBuffer<uint32_t> keys(k0, k1), vals(v0, v1);
        B2_CUDA_TRY(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, vals, (int)n, 0, 32, s));
        ws->launches += 5;  // upsweep/scan/downsweep passes (onesweep: histogram + 4 passes)
        reorder_kernel<<<blocks, 256, 0, s>>>(keys.Current(), vals.Current(), n, c->dx, c->dy,
                                              c->dz, c->pts, c->rank, c->hkeys, c->hvals,
                                              c->hshift, c->hcap - 1, c->grid);
        ws->launches++;
    }
    B2_CUDA_TRY(cudaGetLastError());
    B2_CUDA_TRY(cudaEventRecord(c->ready, s));
    if (prof)
    {
        B2_CUDA_TRY(cudaEventRecord(ws->prof_ev[1], s));
        B2_CUDA_TRY(cudaEventSynchronize(ws->prof_ev[1]));
        float ms = 0;
        B2_CUDA_TRY(cudaEventElapsedTime(&ms, ws->prof_ev[0], ws->prof_ev[1]));
        std::lock_guard<std::mutex> lk(ctx->mtx);
        ctx->prof.index_builds++;
        ctx->prof.index_ms += ms;
        ctx->prof.index_points += n;
    }
    return B200ICP_OK;
}

}  // namespace b2
