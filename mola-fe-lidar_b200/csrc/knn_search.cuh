// knn_search.cuh -- radius-capped exact k-nearest-neighbour search on the
// two-level grid index (replaces kdTreeNClosestPoint3DIdx on the nanoflann
// kd-tree behind mp2p_icp's matchers).
//
// Result = the k smallest (d2 as float32, original index) keys with
// d2 <= cap_d2, ascending -- the tie rule of Appendix A.4. d2 is evaluated
// in the fixed float order of A.3 (the library is built with -fmad=false).
//
// One thread per query. Fine cells are visited in SHELLS of growing Chebyshev
// distance from the query's own fine cell (nearest first); the search stops
// as soon as the next shell cannot hold anything closer than the current k-th
// best. Per shell only the (<= 27) blocks the shell touches are probed in the
// hash, each yields a 64-bit occupancy mask, and the shell's surface inside
// the block is one AND with three precomputed bit patterns -- empty space
// costs one probe per block, dense regions are pruned per fine cell.
#pragma once
#include "device_types.cuh"

namespace b2
{
#ifdef B200ICP_DBG_COUNT
// development probe: candidates scanned per thread since the last reset
static __device__ uint32_t* g_dbg_lane_cand = nullptr;
#endif
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

// 4-bit pattern of the in-block sub indices inside [lo,hi] (0 when empty)
__device__ __forceinline__ uint32_t range4(int lo, int hi)
{
    lo = max(lo, 0);
    hi = min(hi, 3);
    return (lo > hi) ? 0u : ((2u << hi) - (1u << lo));
}
// fine cell (sx,sy,sz) of a block is bit sx + 4*sy + 16*sz of its mask
__device__ __forceinline__ uint64_t expand_x(uint32_t p) { return (uint64_t)p * 0x1111111111111111ull; }
__device__ __forceinline__ uint64_t expand_y(uint32_t p)
{
    const uint32_t n = ((p & 1u) * 0xFu) | ((p & 2u) * 0x78u) | ((p & 4u) * 0x3C0u) | ((p & 8u) * 0x1E00u);
    return (uint64_t)n * 0x0001000100010001ull;
}
__device__ __forceinline__ uint64_t expand_z(uint32_t p)
{
    return ((p & 1u) ? 0xFFFFull : 0ull) | ((p & 2u) ? 0xFFFF0000ull : 0ull) |
           ((p & 4u) ? 0xFFFF00000000ull : 0ull) | ((p & 8u) ? 0xFFFF000000000000ull : 0ull);
}

__device__ __forceinline__ bool block_lookup(const CloudView& cv, uint32_t bkey, uint4& rec)
{
    uint32_t slot = hash_slot(bkey, cv.hshift);
    for (;;)
    {
        const uint32_t k = __ldg(cv.hkeys + slot);
        if (k == bkey)
        {
            rec = __ldg(cv.hrecs + slot);
            return true;
        }
        if (k == kEmptyKey) return false;
        slot = (slot + 1) & cv.hmask;
    }
}

__device__ __forceinline__ float dist2(float qx, float qy, float qz, const float4& c)
{
    const float ddx = qx - c.x, ddy = qy - c.y, ddz = qz - c.z;
    const float a = ddx * ddx;
    const float b = ddy * ddy;
    const float e = ddz * ddz;
    const float ab = a + b;
    return ab + e;  // A.3: ((dx^2 + dy^2) + dz^2), no contraction
}

// Lower bound of dist2(q, p) over every point p inside the box [lo, hi], in
// the SAME operation order as dist2.  Every step (difference, square, sums) is
// monotonic under round-to-nearest and the library is built without FMA
// contraction, so box_lower_d2 <= dist2(q, p) holds for the COMPUTED values,
// not just in exact arithmetic: a group whose bound exceeds the current k-th
// best d2 cannot change the result, ties included.
__device__ __forceinline__ float box_lower_d2(float qx, float qy, float qz, const float4& lo, const float4& hi)
{
    const float dx = fmaxf(fmaxf(lo.x - qx, qx - hi.x), 0.0f);
    const float dy = fmaxf(fmaxf(lo.y - qy, qy - hi.y), 0.0f);
    const float dz = fmaxf(fmaxf(lo.z - qz, qz - hi.z), 0.0f);
    const float a = dx * dx;
    const float b = dy * dy;
    const float e = dz * dz;
    const float ab = a + b;
    return ab + e;
}

// a short run, point by point (four independent loads in flight)
template <int K>
__device__ __forceinline__ void scan_plain(const float4* __restrict__ pts, uint32_t beg, uint32_t end,
                                           float qx, float qy, float qz, uint64_t (&key)[K])
{
    uint32_t j = beg;
    for (; j + 4 <= end; j += 4)
    {
        const float4   c0 = __ldg(pts + j), c1 = __ldg(pts + j + 1), c2 = __ldg(pts + j + 2),
                     c3 = __ldg(pts + j + 3);
        const uint64_t k0 = make_key(dist2(qx, qy, qz, c0), __float_as_uint(c0.w));
        const uint64_t k1 = make_key(dist2(qx, qy, qz, c1), __float_as_uint(c1.w));
        const uint64_t k2 = make_key(dist2(qx, qy, qz, c2), __float_as_uint(c2.w));
        const uint64_t k3 = make_key(dist2(qx, qy, qz, c3), __float_as_uint(c3.w));
        const uint64_t w = key[K - 1];
        if (k0 < w || k1 < w || k2 < w || k3 < w)
        {
            if (k0 < key[K - 1]) topk_insert<K>(key, k0);
            if (k1 < key[K - 1]) topk_insert<K>(key, k1);
            if (k2 < key[K - 1]) topk_insert<K>(key, k2);
            if (k3 < key[K - 1]) topk_insert<K>(key, k3);
        }
    }
    for (; j < end; j++)
    {
        const float4   c = __ldg(pts + j);
        const uint64_t kk = make_key(dist2(qx, qy, qz, c), __float_as_uint(c.w));
        if (kk < key[K - 1]) topk_insert<K>(key, kk);
    }
}

#ifndef B200ICP_PLAIN_RUN
#define B200ICP_PLAIN_RUN 16
#endif
constexpr uint32_t kPlainRun = B200ICP_PLAIN_RUN;  // runs up to this length are scanned without looking at group boxes

// All points of one fine cell, [beg,end) in the sorted array.  Longer runs go
// GROUP by group (kGroup = 8 consecutive sorted points, spatially compact
// because points are sorted by octant inside the cell): the group's bounding
// box is tested against the current k-th best first -- in a dense cell most
// groups are rejected with two 16-byte loads instead of eight -- and a group
// that survives is loaded whole (eight independent loads in flight).
template <int K>
__device__ __forceinline__ void scan_range(const float4* __restrict__ pts, const float4* __restrict__ gbox,
                                           uint32_t beg, uint32_t end, float qx, float qy, float qz,
                                           uint64_t (&key)[K])
{
#ifdef B200ICP_DBG_COUNT
    if (g_dbg_lane_cand) g_dbg_lane_cand[(blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x] += end - beg;
#endif
    if (end - beg <= kPlainRun)
    {
        scan_plain<K>(pts, beg, end, qx, qy, qz, key);
        return;
    }
    const uint32_t g1 = (end - 1) / kGroup;
    for (uint32_t g = beg / kGroup; g <= g1; g++)
    {
        const float4 lo = __ldg(gbox + 2 * g), hi = __ldg(gbox + 2 * g + 1);
        if (box_lower_d2(qx, qy, qz, lo, hi) > key_d2(key[K - 1])) continue;
        const uint32_t base = g * kGroup;
        float4         c[kGroup];
#pragma unroll
        for (int u = 0; u < kGroup; u++)
        {
            // points of the neighbouring cells that share the group are loaded but masked below
            const uint32_t j = min(max(base + u, beg), end - 1);
            c[u] = __ldg(pts + j);
        }
        uint64_t       kk[kGroup];
        bool           any = false;
        const uint64_t w = key[K - 1];
#pragma unroll
        for (int u = 0; u < kGroup; u++)
        {
            const bool in = (base + u >= beg) && (base + u < end);
            kk[u] = in ? make_key(dist2(qx, qy, qz, c[u]), __float_as_uint(c[u].w)) : ~0ull;
            any |= kk[u] < w;
        }
        if (any)
        {
#pragma unroll
            for (int u = 0; u < kGroup; u++)
                if (kk[u] < key[K - 1]) topk_insert<K>(key, kk[u]);
        }
    }
}

// keys must be initialised by the caller with sentinel_key(cap_d2).
template <int K>
__device__ __forceinline__ void knn_search(const CloudView& cv, const GridDev& g, float qx,
                                           float qy, float qz, float cap_d2,
                                           uint64_t (&key)[K])
{
    const float inv = g.inv_cell;
    // shells of fine cells that can hold a point within the cap
    const int   S = max(1, (int)ceilf(sqrtf(cap_d2) * inv * 1.0005f));
    const float lim_lo = -(float)(S + 2), lim_hi = (float)(kFineMax + S + 3);
    float       ux = (qx - g.ox) * inv, uy = (qy - g.oy) * inv, uz = (qz - g.oz) * inv;
    if (!(ux == ux) || !(uy == uy) || !(uz == uz)) return;  // NaN query: no neighbours
    ux = fminf(fmaxf(ux, lim_lo), lim_hi);
    uy = fminf(fmaxf(uy, lim_lo), lim_hi);
    uz = fminf(fmaxf(uz, lim_lo), lim_hi);
    const float flx = floorf(ux), fly = floorf(uy), flz = floorf(uz);
    const int   qfx = (int)flx, qfy = (int)fly, qfz = (int)flz;
    const float fx = ux - flx, fy = uy - fly, fz = uz - flz;
    const float slack = g.slack;
    // distance from the query to the nearest face of its own fine cell
    const float gmin = fmaxf(
        fminf(fminf(fminf(fx, 1.0f - fx), fminf(fy, 1.0f - fy)), fminf(fz, 1.0f - fz)) - slack, 0.0f);
    // worst admissible d2 in fine-cell units, with a relative guard for rounding
    const float to_cells2 = inv * inv * 1.0001f;
    float       worst = key_d2(key[K - 1]) * to_cells2;

    for (int s = 0; s <= S; s++)
    {
        if (s >= 1)
        {  // every cell of shell s is at least (s - 1 + gmin) cells away
            const float b = (float)(s - 1) + gmin;
            if (b * b > worst) break;
        }
        const int x0 = qfx - s, x1 = qfx + s, y0 = qfy - s, y1 = qfy + s, z0 = qfz - s, z1 = qfz + s;
        const int bx0 = max(x0 >> 2, 0), bx1 = min(x1 >> 2, kGridMax);
        const int by0 = max(y0 >> 2, 0), by1 = min(y1 >> 2, kGridMax);
        const int bz0 = max(z0 >> 2, 0), bz1 = min(z1 >> 2, kGridMax);
        for (int bz = bz0; bz <= bz1; bz++)
        {
            const uint64_t Zs = expand_z(range4(z0 - 4 * bz, z1 - 4 * bz));
            const uint64_t Zi = expand_z(s ? range4(z0 + 1 - 4 * bz, z1 - 1 - 4 * bz) : 0u);
            for (int by = by0; by <= by1; by++)
            {
                const uint64_t Ys = expand_y(range4(y0 - 4 * by, y1 - 4 * by));
                const uint64_t Yi = expand_y(s ? range4(y0 + 1 - 4 * by, y1 - 1 - 4 * by) : 0u);
                const uint64_t ZYs = Zs & Ys, ZYi = Zi & Yi;
                for (int bx = bx0; bx <= bx1; bx++)
                {
                    const uint64_t Xs = expand_x(range4(x0 - 4 * bx, x1 - 4 * bx));
                    const uint64_t Xi = expand_x(s ? range4(x0 + 1 - 4 * bx, x1 - 1 - 4 * bx) : 0u);
                    // surface of the shell's cube inside this block
                    uint64_t m = (ZYs & Xs) & ~(ZYi & Xi);
                    if (m == 0) continue;
                    uint4 rec;
                    const uint32_t bkey = (uint32_t)bx | ((uint32_t)by << kGridBits) |
                                          ((uint32_t)bz << (2 * kGridBits));
                    if (!block_lookup(cv, bkey, rec)) continue;
                    const uint64_t occ = ((uint64_t)rec.w << 32) | (uint64_t)rec.z;
                    m &= occ;
                    while (m)
                    {
                        const int bit = __ffsll((long long)m) - 1;
                        m &= m - 1;
                        const int   cx = 4 * bx + (bit & 3), cy = 4 * by + ((bit >> 2) & 3),
                                  cz = 4 * bz + (bit >> 4);
                        const float gx = axis_gap(fx, cx - qfx, slack);
                        const float gy = axis_gap(fy, cy - qfy, slack);
                        const float gz = axis_gap(fz, cz - qfz, slack);
                        if ((gx * gx + gy * gy) + gz * gz > worst) continue;
                        const uint32_t ord = rec.y + (uint32_t)__popcll(occ & ((1ull << bit) - 1ull));
                        const uint32_t beg = __ldg(cv.fine_start + ord);
                        const uint32_t end = __ldg(cv.fine_start + ord + 1);
                        scan_range<K>(cv.pts, cv.gbox, beg, end, qx, qy, qz, key);
                        worst = key_d2(key[K - 1]) * to_cells2;
                    }
                }
            }
        }
    }
}

}  // namespace b2
