// coop_search.cuh -- cooperative exact k-nearest-neighbour search for HEAVY
// queries: the ones the per-lane shell walk (tile_search.cuh) gives up on after
// its candidate budget.  On LiDAR scans 12-16 % of the queries carry more than
// half of the candidates (a query between two rings, or facing a dense ring
// across the sensor's blind disc, has its 6th neighbour far away and sees long
// arcs of points at almost the same distance); walked by one lane each they
// decide how long the whole kernel lasts.
//
// Here EIGHT lanes work on one query (four queries per warp):
//   1. the blocks the query's ball touches (<= 27 when the radius cap is the
//      index's own radius) are probed in the hash, 8 at a time; a present block
//      whose cube is within reach leaves its point range in shared memory;
//   2. every block's range is taken GROUP by group (8 sorted points and their
//      bounding box, cloud.cu): 8 boxes are tested per pass, one per lane;
//   3. a group whose box is within the current k-th best is loaded whole, one
//      point per lane (128 contiguous bytes), and the sub-group's candidates go
//      into the sorted list -- one entry per lane -- smallest first until the
//      smallest left cannot enter.
// All loads of a pass are independent, so a heavy query costs a few dozen
// memory round trips instead of several hundred dependent ones, and no lane
// idles while another scans.  The result is the k smallest (d2, index) keys
// with d2 <= cap, exactly as the per-lane search defines it: any exact search
// gives the same list.
#pragma once
#include "knn_search.cuh"

namespace b2
{
constexpr int kSub = 8;                 // lanes per query
constexpr int kSubPerWarp = 32 / kSub;  // queries per warp
constexpr int kCoopBlocks = 64;         // block ranges kept per query and round

struct CoopWarpSmem
{
    uint2 range[kSubPerWarp][kCoopBlocks];  // [first, end) sorted positions of the blocks in reach
};

// hash probe that also gives the slot (for hrange)
__device__ __forceinline__ bool block_find(const CloudView& cv, uint32_t bkey, uint32_t& slot_out)
{
    uint32_t slot = hash_slot(bkey, cv.hshift);
    for (;;)
    {
        const uint32_t k = __ldg(cv.hkeys + slot);
        if (k == bkey)
        {
            slot_out = slot;
            return true;
        }
        if (k == kEmptyKey) return false;
        slot = (slot + 1) & cv.hmask;
    }
}

__device__ __forceinline__ uint64_t shfl_key(uint64_t v, int src, int width)
{
    const uint32_t lo = __shfl_sync(0xFFFFFFFFu, (uint32_t)v, src, width);
    const uint32_t hi = __shfl_sync(0xFFFFFFFFu, (uint32_t)(v >> 32), src, width);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_up_key(uint64_t v, int d, int width)
{
    const uint32_t lo = __shfl_up_sync(0xFFFFFFFFu, (uint32_t)v, d, width);
    const uint32_t hi = __shfl_up_sync(0xFFFFFFFFu, (uint32_t)(v >> 32), d, width);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_xor_key(uint64_t v, int m)
{
    const uint32_t lo = __shfl_xor_sync(0xFFFFFFFFu, (uint32_t)v, m);
    const uint32_t hi = __shfl_xor_sync(0xFFFFFFFFu, (uint32_t)(v >> 32), m);
    return ((uint64_t)hi << 32) | lo;
}

// Called by all 32 lanes.  Lanes 8s..8s+7 hold the same query (qx,qy,qz,cap)
// of sub-group s, `active` false when the sub-group has none.  On return every
// lane holds its sub-group's full list in key[] (sentinel_key(cap) = not found).
template <int K>
__device__ __forceinline__ void coop_knn(CoopWarpSmem& W, const CloudView& cv, const GridDev& g, bool active,
                                         float qx, float qy, float qz, float cap, uint64_t (&key)[K])
{
    static_assert(K <= kSub, "one list entry per lane of the sub-group");
    const int      lane = threadIdx.x & 31, sg = lane / kSub, ls = lane % kSub;
    const uint64_t sent = sentinel_key(cap);
    uint64_t       mykey = sent;       // lane ls < K holds the ls-th best of its sub-group
    float          worst = cap;        // d2 of the K-th best so far (uniform in the sub-group)

    // blocks the ball can touch: the cells of every point within reach lie in
    // [floor(u - rc - slack), floor(u + rc + slack)] per axis (same reasoning as the shell walk)
    const float inv = g.inv_cell, slack = g.slack;
    const float rc = sqrtf(cap) * inv * 1.0005f;
    int         b0[3] = {0, 0, 0}, nbd[3] = {0, 0, 0};
    float       u[3] = {(qx - g.ox) * inv, (qy - g.oy) * inv, (qz - g.oz) * inv};
    bool        any_block = active && (u[0] == u[0]) && (u[1] == u[1]) && (u[2] == u[2]);
#pragma unroll
    for (int d = 0; d < 3; d++)
    {
        const float lo = fmaxf(u[d] - rc - slack, 0.0f), hi = fminf(u[d] + rc + slack, (float)kFineMax);
        if (!(lo <= hi)) any_block = false;  // also catches NaN
        const int c0 = (int)floorf(fmaxf(lo, 0.0f)), c1 = (int)floorf(fmaxf(hi, 0.0f));
        b0[d] = c0 >> 2;
        nbd[d] = (c1 >> 2) - b0[d] + 1;
    }
    const uint32_t nb = any_block ? (uint32_t)nbd[0] * (uint32_t)nbd[1] * (uint32_t)nbd[2] : 0u;
    const uint32_t nb_max = __reduce_max_sync(0xFFFFFFFFu, nb);
    const float    to_cells2 = inv * inv * 1.0001f;

    for (uint32_t i0 = 0; i0 < nb_max; i0 += kCoopBlocks)
    {
        // ---- 1. probe up to kCoopBlocks blocks per query, 8 per round --------------------
        uint32_t nkept = 0;  // uniform in the sub-group
        __syncwarp();
        for (uint32_t r = 0; r < (uint32_t)kCoopBlocks; r += kSub)
        {
            if (i0 + r >= nb_max) break;  // warp-uniform
            const uint32_t i = i0 + r + ls;
            bool           keep = false;
            uint2          rng = make_uint2(0u, 0u);
            if (i < nb)
            {
                const int bx = b0[0] + (int)(i % (uint32_t)nbd[0]);
                const int by = b0[1] + (int)((i / (uint32_t)nbd[0]) % (uint32_t)nbd[1]);
                const int bz = b0[2] + (int)(i / ((uint32_t)nbd[0] * (uint32_t)nbd[1]));
                // distance from the query to the block's cube, in cells, conservative by `slack`
                const float gx = fmaxf(fmaxf((float)(4 * bx) - u[0], u[0] - (float)(4 * bx + 4)) - slack, 0.0f);
                const float gy = fmaxf(fmaxf((float)(4 * by) - u[1], u[1] - (float)(4 * by + 4)) - slack, 0.0f);
                const float gz = fmaxf(fmaxf((float)(4 * bz) - u[2], u[2] - (float)(4 * bz + 4)) - slack, 0.0f);
                if (!((gx * gx + gy * gy) + gz * gz > worst * to_cells2))
                {
                    const uint32_t bkey =
                        (uint32_t)bx | ((uint32_t)by << kGridBits) | ((uint32_t)bz << (2 * kGridBits));
                    uint32_t slot;
                    if (block_find(cv, bkey, slot))
                    {
                        rng = __ldg(cv.hrange + slot);
                        keep = rng.y > rng.x;
                    }
                }
            }
            const uint32_t m = (__ballot_sync(0xFFFFFFFFu, keep) >> (kSub * sg)) & ((1u << kSub) - 1u);
            if (keep) W.range[sg][nkept + __popc(m & ((1u << ls) - 1u))] = rng;
            nkept += __popc(m);
        }
        __syncwarp();

        // ---- 2. the kept blocks, group by group ------------------------------------------
        uint32_t bi = 0;            // next kept block
        uint32_t beg = 0, end = 0;  // current block's range
        uint32_t gnext = 1, glast = 0;  // groups of the current block still to test: [gnext, glast]
        for (;;)
        {
            // advance to a block with groups left
            bool work = true;
            if (gnext > glast)
            {
                if (bi < nkept)
                {
                    const uint2 rng = W.range[sg][bi++];
                    beg = rng.x, end = rng.y;
                    gnext = beg / kGroup, glast = (end - 1) / kGroup;
                }
                else
                    work = false;
            }
            if (!__any_sync(0xFFFFFFFFu, work)) break;
            // one pass: 8 boxes, one per lane
            const uint32_t myg = gnext + ls;
            bool           hit = false;
            if (work && myg <= glast)
            {
                const float4 lo = __ldg(cv.gbox + 2 * myg), hi = __ldg(cv.gbox + 2 * myg + 1);
                hit = !(box_lower_d2(qx, qy, qz, lo, hi) > worst);
            }
            uint32_t km = (__ballot_sync(0xFFFFFFFFu, hit) >> (kSub * sg)) & ((1u << kSub) - 1u);
            // surviving groups, one per round: lane ls takes point ls of the group
            while (__any_sync(0xFFFFFFFFu, km != 0))
            {
                uint64_t kk = ~0ull;
                if (km)
                {
                    const uint32_t gsel = gnext + (uint32_t)(__ffs((int)km) - 1);
                    km &= km - 1;
                    const uint32_t j = gsel * kGroup + ls;
                    if (j >= beg && j < end)
                    {
                        const float4 c = __ldg(cv.pts + j);
                        kk = make_key(dist2(qx, qy, qz, c), __float_as_uint(c.w));
                    }
                }
                // the sub-group's candidates, smallest first, until the smallest left cannot enter
                for (;;)
                {
                    uint64_t best = kk;
#pragma unroll
                    for (int o = 1; o < kSub; o <<= 1)
                    {
                        const uint64_t other = shfl_xor_key(best, o);
                        best = other < best ? other : best;
                    }
                    const uint64_t kth = shfl_key(mykey, K - 1, kSub);
                    const bool     enter = best < kth;
                    if (!__any_sync(0xFFFFFFFFu, enter)) break;
                    const uint64_t prev = shfl_up_key(mykey, 1, kSub);
                    if (enter)
                    {
                        if (best < mykey) mykey = (ls == 0 || !(best < prev)) ? best : prev;
                        if (kk == best) kk = ~0ull;
                    }
                }
                worst = key_d2(shfl_key(mykey, K - 1, kSub));
            }
            if (work) gnext += kSub;
        }
    }
#pragma unroll
    for (int i = 0; i < K; i++) key[i] = shfl_key(mykey, i, kSub);
}

}  // namespace b2
