// sweep_search.cuh -- tiled brute-force exact k-nearest-neighbour search
// ("tile sweep"), the search stage of the matchers and of the kNN query.
//
// Work unit ("item") = kItem (32) consecutive points of the LOCAL cloud's own
// cell-sorted array: under a rigid transform they stay spatially compact.
// One warp holds the item's queries, one per lane.  The warp
//   1. reduces the box of its queries and grows it by the search radius,
//   2. probes the GLOBAL cloud's block hash for every block the box touches
//      (lanes in parallel; one 8-byte record = the block's point range),
//   3. streams the points of the present blocks with coalesced float4 loads,
//      keeps those inside the box and compacts them into a per-warp staging
//      buffer in shared memory,
//   4. sweeps the staged points: every lane evaluates EVERY staged point
//      against its own query (shared-memory broadcast reads, no divergence,
//      four independent distance evaluations in flight) and keeps its k best
//      64-bit keys (d2 bits << 32 | index) -- the tie rule is the integer
//      order of the key, independent of the order candidates arrive in.
// Several warps of a CTA may share one item (`nparts`): each sweeps every
// nparts-th group of blocks, the per-lane lists are merged through shared
// memory at the end.  Exactness: every point with d2 <= cap is inside the box
// (the box is padded for the rounding of d2 and of its own corners), and block
// coordinates of the box corners use the very expression the index used for the
// points (monotone in the coordinate), so no candidate is ever missed; the
// result equals knn_search<K> / the oracle's kd-tree bit for bit.
//
// Regions touching too many blocks (items straddling a jump of the Morton
// curve) take the per-lane hash walk of knn_search.cuh.
#pragma once
#include "knn_search.cuh"

namespace b2
{
constexpr int kSweepCap = 512;         // staged points per warp (8 KB)
constexpr int kSweepMaxBlocks = 4096;  // blocks probed per item before falling back

__device__ __forceinline__ uint32_t enc_order_f(float f)
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_order_f(uint32_t e)
{
    const uint32_t u = (e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e;
    return __uint_as_float(u);
}

// fine-cell coordinate of a position along one axis: the expression of
// cell_key_kernel (cloud.cu), made safe for positions far outside the grid
__device__ __forceinline__ int sweep_fine_coord(float p, float o, float inv)
{
    const float u = fminf(fmaxf((p - o) * inv, -2.0f), (float)(kFineMax + 2));
    return min(max((int)floorf(u), 0), kFineMax);
}

// all staged points against this lane's query
template <int K>
__device__ __forceinline__ void sweep_staged(const float4* buf, int cnt, float qx, float qy, float qz,
                                             uint64_t (&key)[K])
{
    int j = 0;
    for (; j + 4 <= cnt; j += 4)
    {
        const float4   c0 = buf[j], c1 = buf[j + 1], c2 = buf[j + 2], c3 = buf[j + 3];
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
    for (; j < cnt; j++)
    {
        const float4   c = buf[j];
        const uint64_t kk = make_key(dist2(qx, qy, qz, c), __float_as_uint(c.w));
        if (kk < key[K - 1]) topk_insert<K>(key, kk);
    }
}

// Warp-collective.  `valid` lanes hold a finite query (qx,qy,qz); key[] must
// hold sentinel_key(cap_d2) on entry.  part / nparts: this warp's share of the
// region's blocks.  buf: kSweepCap float4 of shared memory private to the warp.
template <int K>
__device__ __forceinline__ void sweep_search(float4* buf, const CloudView& cv, const GridDev& g,
                                             bool valid, float qx, float qy, float qz, float cap_d2,
                                             int part, int nparts, uint64_t (&key)[K])
{
    const int      lane = threadIdx.x & 31;
    const unsigned full = 0xFFFFFFFFu;
    // ---- region: box of the queries grown by the radius ---------------------
    uint32_t elo[3], ehi[3];
    {
        const float q[3] = {qx, qy, qz};
#pragma unroll
        for (int d = 0; d < 3; d++)
        {
            elo[d] = __reduce_min_sync(full, valid ? enc_order_f(q[d]) : 0xFFFFFFFFu);
            ehi[d] = __reduce_max_sync(full, valid ? enc_order_f(q[d]) : 0u);
        }
    }
    if (elo[0] > ehi[0]) return;  // no lane searches
    const float r = sqrtf(cap_d2) * 1.001f;
    float       blo[3], bhi[3];
#pragma unroll
    for (int d = 0; d < 3; d++)
    {
        const float lo = dec_order_f(elo[d]), hi = dec_order_f(ehi[d]);
        const float pad = r + 1e-6f * (fabsf(lo) + fabsf(hi) + 1.0f);
        blo[d] = lo - pad, bhi[d] = hi + pad;
    }
    if (!(bhi[0] >= g.bmin[0] && blo[0] <= g.bmax[0] && bhi[1] >= g.bmin[1] && blo[1] <= g.bmax[1] &&
          bhi[2] >= g.bmin[2] && blo[2] <= g.bmax[2]))
        return;  // the region misses the global cloud altogether
    const float inv = g.inv_cell;
    const int   bx0 = sweep_fine_coord(blo[0], g.ox, inv) >> 2, bx1 = sweep_fine_coord(bhi[0], g.ox, inv) >> 2;
    const int   by0 = sweep_fine_coord(blo[1], g.oy, inv) >> 2, by1 = sweep_fine_coord(bhi[1], g.oy, inv) >> 2;
    const int   bz0 = sweep_fine_coord(blo[2], g.oz, inv) >> 2, bz1 = sweep_fine_coord(bhi[2], g.oz, inv) >> 2;
    const int   nbx = bx1 - bx0 + 1, nby = by1 - by0 + 1, nbz = bz1 - bz0 + 1;
    const int   nb = nbx * nby * nbz;
    if (nbx > 64 || nby > 64 || nbz > 64 || nb > kSweepMaxBlocks)
    {  // very spread item: per-lane walk over the hash, same results
        if (part == 0 && valid) knn_search<K>(cv, g, qx, qy, qz, cap_d2, key);
        return;
    }
    // lanes without a query evaluate against a point at infinity: never inserted
    const float  ex = valid ? qx : INFINITY, ey = valid ? qy : INFINITY, ez = valid ? qz : INFINITY;
    const unsigned lt = (1u << lane) - 1u;
    int          cnt = 0;
    for (int i0 = part * 32; i0 < nb; i0 += 32 * nparts)
    {
        const int i = i0 + lane;
        uint2     rng = make_uint2(0u, 0u);
        if (i < nb)
        {
            const int      bx = bx0 + i % nbx, by = by0 + (i / nbx) % nby, bz = bz0 + i / (nbx * nby);
            const uint32_t bkey = (uint32_t)bx | ((uint32_t)by << kGridBits) | ((uint32_t)bz << (2 * kGridBits));
            uint32_t       slot = hash_slot(bkey, cv.hshift);
            for (;;)
            {
                const uint32_t k = __ldg(cv.hkeys + slot);
                if (k == bkey)
                {
                    rng = __ldg(cv.hrange + slot);
                    break;
                }
                if (k == kEmptyKey) break;
                slot = (slot + 1) & cv.hmask;
            }
        }
        unsigned present = __ballot_sync(full, rng.y > rng.x);
        while (present)
        {
            const int src = __ffs((int)present) - 1;
            present &= present - 1;
            const uint32_t beg = __shfl_sync(full, rng.x, src), end = __shfl_sync(full, rng.y, src);
            for (uint32_t j0 = beg; j0 < end; j0 += 64)
            {
                // two coalesced loads in flight per lane
                const uint32_t ja = j0 + lane, jb = j0 + 32 + lane;
                float4         pa = make_float4(0.f, 0.f, 0.f, 0.f), pb = pa;
                if (ja < end) pa = __ldg(cv.pts + ja);
                if (jb < end) pb = __ldg(cv.pts + jb);
                const bool ka = (ja < end) && pa.x >= blo[0] && pa.x <= bhi[0] && pa.y >= blo[1] &&
                                pa.y <= bhi[1] && pa.z >= blo[2] && pa.z <= bhi[2];
                const bool kb = (jb < end) && pb.x >= blo[0] && pb.x <= bhi[0] && pb.y >= blo[1] &&
                                pb.y <= bhi[1] && pb.z >= blo[2] && pb.z <= bhi[2];
                const unsigned ma = __ballot_sync(full, ka), mb = __ballot_sync(full, kb);
                if (ka) buf[cnt + __popc(ma & lt)] = pa;
                cnt += __popc(ma);
                if (kb) buf[cnt + __popc(mb & lt)] = pb;
                cnt += __popc(mb);
                if (cnt > kSweepCap - 64)
                {
                    __syncwarp();
                    sweep_staged<K>(buf, cnt, ex, ey, ez, key);
                    cnt = 0;
                    __syncwarp();
                }
            }
        }
    }
    __syncwarp();
    sweep_staged<K>(buf, cnt, ex, ey, ez, key);
    __syncwarp();
}

}  // namespace b2
