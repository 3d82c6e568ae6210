// cloud.cu -- device-resident clouds and the GPU-built search index.
//
// Replaces the lazily built nanoflann kd-tree on the `from` cloud
// (mrpt::math::KDTreeCapable) with a uniform grid:
//   bbox -> per-point key (30-bit Morton of the block | 6-bit fine cell |
//   3-bit octant in the fine cell) -> radix sort (key, index) -> points
//   gathered as float4 in cell order + inverse permutation + the bounding box
//   of every group of 8 sorted points -> scan of fine-cell heads -> open-addressing hash of
//   occupied blocks {start, first fine cell, 64-bit occupancy mask} and the
//   start offset of every occupied fine cell.
// Everything after the H2D copy happens on the device with no host sync.
// Algorithmic traffic: 12 B read + 16 B float4 write + 4 B rank write per
// point, plus the sort passes (36 B/point).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>

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

__global__ void grid_setup_kernel(const uint32_t* __restrict__ bb, float block_req, GridDev* g)
{
    float mn[3], mx[3];
    for (int d = 0; d < 3; d++) mn[d] = dec_f(bb[d]), mx[d] = dec_f(bb[3 + d]);
    const bool any = mn[0] <= mx[0];
    if (!any)
        for (int d = 0; d < 3; d++) mn[d] = mx[d] = 0.f;
    float ext = fmaxf(fmaxf(mx[0] - mn[0], mx[1] - mn[1]), mx[2] - mn[2]);
    // at most 1000 of the 1024 blocks per axis may be spanned
    const float block = fmaxf(block_req, ext / 1000.0f);
    const float cell = block * 0.25f;
    g->ox = mn[0], g->oy = mn[1], g->oz = mn[2];
    g->cell = cell;
    g->inv_cell = 1.0f / cell;
    g->slack = 2.5e-3f;  // > ulp(4096) = 4.9e-4 fine cells of rounding in (p - o) * inv_cell
    g->n_valid = 0;
    g->n_cells = 0;
    g->n_blocks = 0;
    g->n_items = 0;
    for (int d = 0; d < 3; d++) g->bmin[d] = mn[d], g->bmax[d] = mx[d];
}


// sort key = Morton30 of the block | fine cell inside the block | octant inside the fine cell
__global__ void cell_key_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                const float* __restrict__ z, uint32_t n,
                                const GridDev* __restrict__ g, unsigned long long* __restrict__ keys,
                                uint32_t* __restrict__ vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float        px = x[i], py = y[i], pz = z[i];
    unsigned long long key = kInvalidSortKey;
    if (isfinite(px) && isfinite(py) && isfinite(pz))
    {
        // half-cell coordinates: floor(2u) >> 1 == floor(u) exactly, so the fine
        // cell is the one the search assumes and the low bit is the octant
        const float    inv = g->inv_cell;
        const int      hmax = 2 * kFineMax + 1;
        const uint32_t hx = (uint32_t)min(max((int)floorf(((px - g->ox) * inv) * 2.0f), 0), hmax);
        const uint32_t hy = (uint32_t)min(max((int)floorf(((py - g->oy) * inv) * 2.0f), 0), hmax);
        const uint32_t hz = (uint32_t)min(max((int)floorf(((pz - g->oz) * inv) * 2.0f), 0), hmax);
        key = point_sort_key(hx, hy, hz);
    }
    keys[i] = key;
    vals[i] = i;
}

// After the sort: gather float4 points in cell order, write the inverse
// permutation, flag the first point of every fine cell and of every work item.
__global__ void gather_kernel(const unsigned long long* __restrict__ skeys,
                              const uint32_t* __restrict__ svals, uint32_t n,
                              const float* __restrict__ x, const float* __restrict__ y,
                              const float* __restrict__ z, float4* __restrict__ pts,
                              float4* __restrict__ gbox, uint32_t* __restrict__ rank,
                              unsigned long long* __restrict__ flags, GridDev* __restrict__ g)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    // bounding box of every group of kGroup sorted points: 8 adjacent lanes
    // (no early return before the shuffles; n is padded by the last warp)
    const bool               in = j < n;
    const unsigned long long key = in ? skeys[j] : kInvalidSortKey;
    const uint32_t           i = in ? svals[j] : 0u;
    const bool               ok = key < kInvalidSortKey;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (ok) px = x[i], py = y[i], pz = z[i];
    {
        float lo[3] = {ok ? px : INFINITY, ok ? py : INFINITY, ok ? pz : INFINITY};
        float hi[3] = {ok ? px : -INFINITY, ok ? py : -INFINITY, ok ? pz : -INFINITY};
#pragma unroll
        for (int o = 1; o < kGroup; o <<= 1)
#pragma unroll
            for (int d = 0; d < 3; d++)
            {
                lo[d] = fminf(lo[d], __shfl_xor_sync(0xFFFFFFFFu, lo[d], o));
                hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xFFFFFFFFu, hi[d], o));
            }
        if (in && (j % kGroup) == 0)
        {
            gbox[2 * (j / kGroup)] = make_float4(lo[0], lo[1], lo[2], 0.f);
            gbox[2 * (j / kGroup) + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
        }
    }
    if (!in) return;
    if (j == 0)
    {  // n_valid = first position holding the invalid key
        uint32_t lo = 0, hi = n;
        while (lo < hi)
        {
            const uint32_t mid = (lo + hi) >> 1;
            if (skeys[mid] < kInvalidSortKey)
                lo = mid + 1;
            else
                hi = mid;
        }
        g->n_valid = lo;
    }
    if (key >= kInvalidSortKey)
    {
        rank[i] = kInvalid;
        flags[j] = 0ull;
        return;
    }
    pts[j] = make_float4(px, py, pz, __uint_as_float(i));
    rank[i] = j;
    // low word: first point of a fine cell; high word: first point of a work
    // item -- one 64-bit scan numbers both.  Items are plain runs of kItem
    // consecutive sorted points: dense regions give compact tiles, a run that
    // straddles far-apart blocks simply takes the per-lane fallback search
    // (measured: cutting at block-group heads as well was slightly slower)
    const unsigned long long prev = j ? skeys[j - 1] : ~0ull;
    const unsigned long long fine = (j == 0 || fine_of_key(prev) != fine_of_key(key)) ? 1ull : 0ull;
    const unsigned long long item = ((j % kItem) == 0) ? 1ull : 0ull;
    flags[j] = fine | (item << 32);
}

__device__ __forceinline__ uint32_t block_key_of(unsigned long long sort_key)
{
    const uint32_t mort = (uint32_t)block_of_key(sort_key);
    return compact10(mort) | (compact10(mort >> 1) << kGridBits) | (compact10(mort >> 2) << (2 * kGridBits));
}

// first point of every block claims a hash slot and writes its record
__global__ void block_insert_kernel(const unsigned long long* __restrict__ skeys, uint32_t n,
                                    const unsigned long long* __restrict__ ords,
                                    uint32_t* __restrict__ hkeys, uint4* __restrict__ hrecs,
                                    uint32_t hshift, uint32_t hmask, GridDev* __restrict__ g)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const unsigned long long key = skeys[j];
    if (key >= kInvalidSortKey) return;
    if (j != 0 && block_of_key(skeys[j - 1]) == block_of_key(key)) return;
    const uint32_t bkey = block_key_of(key);
    uint32_t       slot = hash_slot(bkey, hshift);
    for (;;)
    {
        const uint32_t prev = atomicCAS(hkeys + slot, kEmptyKey, bkey);
        if (prev == kEmptyKey) break;
        slot = (slot + 1) & hmask;
    }
    hrecs[slot] = make_uint4(j, (uint32_t)ords[j], 0u, 0u);
    atomicAdd(&g->n_blocks, 1u);
}

// first point of every fine cell publishes its start and sets its mask bit;
// first point of every work item publishes the item
__global__ void fine_publish_kernel(const unsigned long long* __restrict__ skeys, uint32_t n,
                                    const unsigned long long* __restrict__ flags,
                                    const unsigned long long* __restrict__ ords,
                                    const uint32_t* __restrict__ hkeys, uint4* __restrict__ hrecs,
                                    uint2* __restrict__ hrange, uint32_t hshift, uint32_t hmask,
                                    uint32_t* __restrict__ fine_start,
                                    uint32_t* __restrict__ item_first, GridDev* __restrict__ g)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const unsigned long long key = skeys[j];
    if (key >= kInvalidSortKey) return;
    const unsigned long long fl = flags[j], od = ords[j];
    const uint32_t fine_flag = (uint32_t)fl, item_flag = (uint32_t)(fl >> 32);
    const uint32_t fine_ord = (uint32_t)od, item_ord = (uint32_t)(od >> 32);
    const bool last = (j + 1 == n) || (skeys[j + 1] >= kInvalidSortKey);
    if (last)
    {
        const uint32_t total = fine_ord + fine_flag;
        fine_start[total] = j + 1;
        g->n_cells = total;
        const uint32_t items = item_ord + item_flag;
        item_first[items] = j + 1;
        g->n_items = items;
    }
    if (item_flag) item_first[item_ord] = j;
    // the last point of every block publishes the block's end, the first its start
    const bool blk_last = last || (block_of_key(skeys[j + 1]) != block_of_key(key));
    const bool blk_first = (j == 0) || (block_of_key(skeys[j - 1]) != block_of_key(key));
    if (!fine_flag && !blk_last) return;
    const uint32_t bkey = block_key_of(key);
    uint32_t       slot = hash_slot(bkey, hshift);
    while (hkeys[slot] != bkey) slot = (slot + 1) & hmask;
    if (blk_last) hrange[slot].y = j + 1;
    if (blk_first) hrange[slot].x = j;
    if (!fine_flag) return;
    fine_start[fine_ord] = j;
    unsigned long long* mask = reinterpret_cast<unsigned long long*>(&hrecs[slot].z);
    atomicOr(mask, 1ull << (uint32_t)(fine_of_key(key) & 63ull));
}

int cloud_alloc(::b200icp* ctx, Workspace* ws, size_t n, float search_radius, b200icp_cloud** out, float min_cell,
                bool coords_only)
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
    // block edge: the search radius (27 blocks always cover it), or more when the caller knows the points are far
    // apart (one point per voxel after decimation): with fine cells much smaller than the point spacing the shell
    // walk would cross mostly empty cells -- up to 4 x 4 x 4 fine cells of the radius' size
    c->cell_req = radius * 1.002f;
    if (min_cell > 0.25f * c->cell_req) c->cell_req = 4.0f * std::min(min_cell, c->cell_req);
    c->indexed = !coords_only;
    uint32_t cap = 1024;
    while (!coords_only && cap < 2 * n) cap <<= 1;
    c->hcap = cap;
    uint32_t lg = 0;
    while ((1u << lg) < cap) lg++;
    c->hshift = 32 - lg;
    const size_t nn = n ? n : 1;
    Carver cv(nullptr);
    auto layout = [&](Carver& k) {
        c->dx = k.take<float>(nn), c->dy = k.take<float>(nn), c->dz = k.take<float>(nn);
        if (coords_only) return;
        c->pts = k.take<float4>(nn);
        c->gbox = k.take<float4>(2 * ((nn + kGroup - 1) / kGroup));
        c->rank = k.take<uint32_t>(nn);
        c->hkeys = k.take<uint32_t>(cap);
        c->hrecs = k.take<uint4>(cap);
        c->hrange = k.take<uint2>(cap);
        c->fine_start = k.take<uint32_t>(nn + 1);
        c->item_first = k.take<uint32_t>(nn + 1);
        c->grid = k.take<GridDev>(1);
        c->bbox_enc = k.take<uint32_t>(8);
    };
    layout(cv);
    c->slab = ctx->take_slab(cv.off, &c->slab_bytes, ws->stream);  // recycled from a freed cloud, else the pool
    if (!c->slab)
    {
        set_error("device allocation of %zu bytes for a cloud failed", cv.off);
        delete c;
        return B200ICP_ERR_NOMEM;
    }
    Carver real(c->slab);
    layout(real);
    c->ready = ctx->take_ready_event();
    if (!c->ready)
    {
        set_error("cudaEventCreate failed");
        ctx->give_slab(c->slab, c->slab_bytes, ws->stream);
        delete c;
        return B200ICP_ERR_CUDA;
    }
    *out = c;
    return B200ICP_OK;
}

int cloud_build_index(::b200icp* ctx, Workspace* ws, b200icp_cloud* c)
{
    cudaStream_t   s = ws->stream;
    if (!c->indexed)
    {   // coordinates only (the input of a filter stage): ready as soon as the copy is
        B2_CUDA_TRY(cudaEventRecord(c->ready, s));
        return B200ICP_OK;
    }
    const uint32_t n = (uint32_t)c->n;
    const bool     prof = ctx->profile_on;
    cudaEvent_t    pe0 = nullptr, pe1 = nullptr;
    if (prof)
    {
        {
            std::lock_guard<std::mutex> lk(ctx->mtx);
            if (ctx->event_pool.size() >= 2)
            {
                pe0 = ctx->event_pool.back(), ctx->event_pool.pop_back();
                pe1 = ctx->event_pool.back(), ctx->event_pool.pop_back();
            }
        }
        if (!pe0)
        {
            B2_CUDA_TRY(cudaEventCreate(&pe0));
            B2_CUDA_TRY(cudaEventCreate(&pe1));
        }
        B2_CUDA_TRY(cudaEventRecord(pe0, s));
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
        size_t sort_bytes = 0, scan_bytes = 0;
        cub::DoubleBuffer<unsigned long long> dk(nullptr, nullptr);
        cub::DoubleBuffer<uint32_t>           dv(nullptr, nullptr);
        B2_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, dk, dv, (int)n, 0, kSortBits, s));
        B2_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (unsigned long long*)nullptr,
                                                  (unsigned long long*)nullptr, (int)n, s));
        const size_t temp_bytes = std::max(sort_bytes, scan_bytes);
        unsigned long long *k0, *k1, *fflag, *ford;
        uint32_t *          v0, *v1;
        void*               temp;
        auto layout = [&](Carver& k) {
            k0 = k.take<unsigned long long>(n), k1 = k.take<unsigned long long>(n);
            v0 = k.take<uint32_t>(n), v1 = k.take<uint32_t>(n);
            fflag = k.take<unsigned long long>(n), ford = k.take<unsigned long long>(n);
            temp = k.take<char>(temp_bytes);
        };
        Carver cv(nullptr);
        layout(cv);
        if (int r = ws->reserve_device(cv.off)) return r;
        Carver k(ws->d_scratch);
        layout(k);
        const int blocks = (int)((n + 255) / 256);
        cell_key_kernel<<<blocks, 256, 0, s>>>(c->dx, c->dy, c->dz, n, c->grid, k0, v0);
        cub::DoubleBuffer<unsigned long long> keys(k0, k1);
        cub::DoubleBuffer<uint32_t>           vals(v0, v1);
        B2_CUDA_TRY(cub::DeviceRadixSort::SortPairs(temp, sort_bytes, keys, vals, (int)n, 0, kSortBits, s));
        gather_kernel<<<blocks, 256, 0, s>>>(keys.Current(), vals.Current(), n, c->dx, c->dy, c->dz,
                                             c->pts, c->gbox, c->rank, fflag, c->grid);
        B2_CUDA_TRY(cub::DeviceScan::ExclusiveSum(temp, scan_bytes, fflag, ford, (int)n, s));
        block_insert_kernel<<<blocks, 256, 0, s>>>(keys.Current(), n, ford, c->hkeys, c->hrecs,
                                                   c->hshift, c->hcap - 1, c->grid);
        fine_publish_kernel<<<blocks, 256, 0, s>>>(keys.Current(), n, fflag, ford, c->hkeys, c->hrecs,
                                                   c->hrange, c->hshift, c->hcap - 1, c->fine_start,
                                                   c->item_first, c->grid);
        ws->launches += 4 + 6 + 2;  // ours + radix sort passes + scan
    }
    B2_CUDA_TRY(cudaGetLastError());
    B2_CUDA_TRY(cudaEventRecord(c->ready, s));
    if (prof)
    {
        B2_CUDA_TRY(cudaEventRecord(pe1, s));
        std::lock_guard<std::mutex> lk(ctx->mtx);
        ctx->pending_index.push_back({pe0, pe1, n});
        if (ctx->pending_index.size() > 4096) ctx->drain_pending();
    }
    return B200ICP_OK;
}

}  // namespace b2
