// knn_search.cuh -- radius-capped exact k-nearest-neighbour search on the
// uniform-grid / cell-hash index (replaces kdTreeNClosestPoint3DIdx on the
// nanoflann kd-tree behind mp2p_icp's matchers).
//
// Result = the k smallest (d2 as float32, original index) keys with
// d2 <= cap_d2, ascending -- the tie rule of Appendix A.4. d2 is evaluated
// in the fixed float order of A.3 (the library is built with -fmad=false).
//
// One thread per query. Cells are visited nearest-first (zig-zag over the
// offsets of each axis) and a cell / row / slab is skipped as soon as its
// minimum possible distance exceeds the current k-th best, so in dense regions
// only the query's own cell and the few cells it nearly touches are scanned.
#pragma once
#include "device_types.cuh"

namespace b2
{
__device__ __forceinline__ uint64_t make_key(float d2, uint32_t idx)
{
    return ((uint64_t)__float_as_uint(d2) << 32) | (uint64_t)idx;
}
__device__ __forceinline__ uint64_t sentinel_key(float cap_d2)
{
    return ((uint64_t)__float_as_uint(cap_d2) << 32) | 0xFFFFFFFFull;
}
__device__ __forceinline__ float key_d2(uint64_t k) { return __uint_as_float((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_idx(uint64_t k) { return (uint32_t)(k & 0xFFFFFFFFull); }

__device__ __forceinline__ int zig(int s) { return (s & 1) ? -((s + 1) >> 1) : (s >> 1); }

// distance (in cell units) from a point with fractional in-cell position f to
// the cell `d` cells away along one axis, made conservative by `slack`
__device__ __forceinline__ float axis_gap(float f, int d, float slack)
{
    float g;
    if (d == 0)
        return 0.0f;
    else if (d < 0)
        g = f + (float)(-d - 1);
    else
        g = (1.0f - f) + (float)(d - 1);
    return fmaxf(g - slack, 0.0f);
}

template <int K>
__device__ __forceinline__ void topk_insert(uint64_t (&key)[K], uint64_t kk)
{
    key[K - 1] = kk;
#pragma unroll
    for (int i = K - 1; i > 0; i--)
    {
        const uint64_t a = key[i - 1], b = key[i];
        const bool sw = b < a;
        key[i - 1] = sw ? b : a;
        key[i] = sw ? a : b;
    }
}

// Looks the linear cell key up; returns [start,end) or an empty range.
__device__ __forceinline__ uint2 cell_lookup(const CloudView& cv, uint32_t ckey)
{
    uint32_t slot = hash_slot(ckey, cv.hshift);
    for (;;)
    {
        const uint32_t k = __ldg(cv.hkeys + slot);
        if (k == ckey) return __ldg(cv.hvals + slot);
        if (k == kEmptyKey) return make_uint2(0u, 0u);
        slot = (slot + 1) & cv.hmask;
    }
}

// keys must be initialised by the caller with sentinel_key(cap_d2).
template <int K>
__device__ __forceinline__ void knn_search(const CloudView& cv, const GridDev& g, float qx,
                                           float qy, float qz, float cap_d2,
                                           uint64_t (&key)[K])
{
    // rings of cells that can hold a point within the cap
    const float cap = sqrtf(cap_d2);
    const int   R = max(1, (int)ceilf(cap * g.inv_cell * 1.0005f));
    const float lim_lo = -(float)(R + 2), lim_hi = (float)(kGridMax + R + 3);
    float ux = (qx - g.ox) * g.inv_cell, uy = (qy - g.oy) * g.inv_cell,
          uz = (qz - g.oz) * g.inv_cell;
    if (!(ux == ux) || !(uy == uy) || !(uz == uz)) return;  // NaN query: no neighbours
    ux = fminf(fmaxf(ux, lim_lo), lim_hi);
    uy = fminf(fmaxf(uy, lim_lo), lim_hi);
    uz = fminf(fmaxf(uz, lim_lo), lim_hi);
    const float flx = floorf(ux), fly = floorf(uy), flz = floorf(uz);
    const int   cqx = (int)flx, cqy = (int)fly, cqz = (int)flz;
    const float fx = ux - flx, fy = uy - fly, fz = uz - flz;
    const float slack = g.slack;
    // worst admissible d2 in cell units, with a relative guard for rounding
    const float to_cells2 = g.inv_cell * g.inv_cell;
    float worst = key_d2(key[K - 1]) * to_cells2 * 1.0001f;

    for (int sz = 0; sz <= 2 * R; sz++)
    {
        const int dz = zig(sz);
        const int cz = cqz + dz;
        if (cz < 0 || cz > kGridMax) continue;
        const float gz = axis_gap(fz, dz, slack);
        const float gz2 = gz * gz;
        if (gz2 > worst) continue;
        for (int sy = 0; sy <= 2 * R; sy++)
        {
            const int dy = zig(sy);
            const int cy = cqy + dy;
            if (cy < 0 || cy > kGridMax) continue;
            const float gy = axis_gap(fy, dy, slack);
            const float gzy2 = gz2 + gy * gy;
            if (gzy2 > worst) continue;
            for (int sx = 0; sx <= 2 * R; sx++)
            {
                const int dx = zig(sx);
                const int cx = cqx + dx;
                if (cx < 0 || cx > kGridMax) continue;
                const float gx = axis_gap(fx, dx, slack);
                if (gzy2 + gx * gx > worst) continue;
                const uint32_t ckey =
                    (uint32_t)cx | ((uint32_t)cy << kGridBits) | ((uint32_t)cz << (2 * kGridBits));
                const uint2 range = cell_lookup(cv, ckey);
                for (uint32_t j = range.x; j < range.y; j++)
                {
                    const float4 c = __ldg(cv.pts + j);
                    const float  ddx = qx - c.x, ddy = qy - c.y, ddz = qz - c.z;
                    const float  a = ddx * ddx;
                    const float  b = ddy * ddy;
                    const float  e = ddz * ddz;
                    const float  ab = a + b;
                    const float  d2 = ab + e;
                    const uint64_t kk = make_key(d2, __float_as_uint(c.w));
                    if (kk < key[K - 1])
                    {
                        topk_insert<K>(key, kk);
                        worst = key_d2(key[K - 1]) * to_cells2 * 1.0001f;
                    }
                }
            }
        }
    }
}

}  // namespace b2
