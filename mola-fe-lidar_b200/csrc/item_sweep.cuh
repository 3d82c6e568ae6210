// item_sweep.cuh -- the search stage of BATCHED launches (8 or more jobs in one
// launch: Monte-Carlo loop, loop-closure candidates; align.cu launch_search_k):
// one warp SWEEPS the points around its item instead of every lane walking cells
// on its own.  Converged and lighter in instructions than the walk
// (tile_search.cuh), which stays the stage of single jobs, where a launch is one
// wave of items and lasts as long as its slowest item.
//
// An item is 32 spatially compact queries (one per lane: consecutive points of
// the local cloud's own cell-sorted order).  The warp works in PASSES; a pass
//   1. picks a box: first the cells its queries fall in (pass 0, skipped when the
//      lanes already hold a bound from the previous outer iteration's neighbours,
//      align.cu seeded_cap), then the box of the queries grown by the largest
//      k-th-best distance any lane holds so far -- the k nearest of lane q lie
//      inside the ball (q, d_k(q)), hence inside that box;
//   2. lists the runs of sorted points of the index cells the box touches and
//      earlier passes have not covered (lane = block: one hash probe, the
//      block's 64-bit occupancy mask cut to the box, consecutive occupied cells
//      are one contiguous run of the sorted array);
//   3. streams the runs with coalesced float4 loads -- every lane a different
//      point of the flattened run list, four loads in flight -- keeps the points
//      inside the box in a shared-memory stage, and
//   4. lets ALL lanes evaluate EVERY staged point against their own query:
//      broadcast shared-memory reads, no divergence.  A lane appends the keys
//      (d2 bits << 32 | index) below its current bound to a private list in
//      shared memory (one predicated store); the warp folds the lists into the
//      lanes' k best keys whenever one of them fills up and at the end of the
//      pass, which also tightens every lane's bound.
// The result is the k smallest keys -- the tie rule of Appendix A.4 -- whatever
// order the candidates arrive in, bit for bit what knn_search<K> returns.
//
// Lanes whose bound is much larger than their neighbours' (fewer than k points
// within the radius) would blow the box up: when a box holds more than
// kSwBudget points the pass is retried for the lanes with a bound below 1/2,
// 1/4 of the largest; lanes left out search alone afterwards (knn_search.cuh),
// bounded by what the sweep found for them.  Items spread over far-apart blocks
// are split in two (align.cu).
#pragma once
#include "knn_search.cuh"
#include "sweep_search.cuh"

namespace b2
{
// development probe (-DB200ICP_DBG_PHASES): cycles per phase of the search, summed over warps
#ifdef B200ICP_DBG_PHASES
__device__ unsigned long long g_dbg_phase[16];
#define B2_PHASE_DECL long long b2_ph_t = clock64()
#define B2_PHASE(i)                                                                                      \
    do                                                                                                   \
    {                                                                                                    \
        const long long b2_now = clock64();                                                              \
        if ((threadIdx.x & 31) == 0) atomicAdd(&g_dbg_phase[i], (unsigned long long)(b2_now - b2_ph_t)); \
        b2_ph_t = b2_now;                                                                                \
    } while (0)
#define B2_COUNT(i, v)                                                 \
    do                                                                 \
    {                                                                  \
        const unsigned long long b2_v = (unsigned long long)(v);       \
        if ((threadIdx.x & 31) == 0) atomicAdd(&g_dbg_phase[i], b2_v); \
    } while (0)
#else
#define B2_PHASE_DECL
#define B2_PHASE(i)
#define B2_COUNT(i, v)
#endif

constexpr int kSwRanges = 224;   // runs of sorted points per pass
constexpr int kSwBlocks = 384;   // blocks the box of a pass may touch
#ifndef B200ICP_SW_STAGE
#define B200ICP_SW_STAGE 256
#endif
constexpr int kSwStage = B200ICP_SW_STAGE;  // staged points per evaluation chunk (>= 160; the TMA variant halves it)
constexpr int kSwList = 16;      // keys a lane may collect between two folds
constexpr int kSwBudget = 2048;  // points one pass may load

struct SweepSmem
{
    uint2    rng[kSwRanges + 1];  // (first sorted position, exclusive prefix of the run lengths); [n].y = total
    uint32_t nrng;
    uint32_t pad;
    float4   stage[kSwStage + 8];  // one buffer of kSwStage (+4 padding) points, or two halves of kSwStage/2 (+4)
    uint64_t list[kSwList * 32];   // [slot][lane]
    uint64_t mbar[2];              // one mbarrier per half of the stage (bulk-copy variant)
};

constexpr uint32_t kSwOver = 0xFFFFFFFFu;   // more points than the budget
constexpr uint32_t kSwWide = 0xFFFFFFFEu;   // more blocks than one pass probes: the lanes are too far apart

struct CellBox
{
    int lo[3], hi[3];  // absolute fine-cell coordinates, inclusive; lo > hi: empty
};

// Warp-collective.  Lists the runs of points of the cells inside `box` and
// outside `skip` into S.rng (with the exclusive prefix of their lengths) and
// returns their total length, kSwOver or kSwWide.
__device__ __noinline__ uint32_t sweep_enumerate(const CloudView& cv, SweepSmem& S, const CellBox& box,
                                                 const CellBox& skip, uint32_t budget)
{
    const int      lane = threadIdx.x & 31;
    const unsigned full = 0xFFFFFFFFu;
    if (lane == 0) S.nrng = 0u;
    __syncwarp();
    uint32_t mine = 0;
    if (box.lo[0] <= box.hi[0] && box.lo[1] <= box.hi[1] && box.lo[2] <= box.hi[2])
    {
        const int bxa = box.lo[0] >> 2, nbx = (box.hi[0] >> 2) - bxa + 1;
        const int bya = box.lo[1] >> 2, nby = (box.hi[1] >> 2) - bya + 1;
        const int bza = box.lo[2] >> 2, nbz = (box.hi[2] >> 2) - bza + 1;
        if (nbx > kSwBlocks || nby > kSwBlocks || nbz > kSwBlocks || nbx * nby * nbz > kSwBlocks) return kSwWide;
        const int  nb = nbx * nby * nbz;
        const bool skipping = skip.lo[0] <= skip.hi[0] && skip.lo[1] <= skip.hi[1] && skip.lo[2] <= skip.hi[2];
        for (int i0 = 0; i0 < nb; i0 += 32)
        {
            const int i = i0 + lane;
            uint4     rec = make_uint4(0u, 0u, 0u, 0u);
            int       bx = 0, by = 0, bz = 0;
            if (i < nb)
            {
                bx = bxa + i % nbx, by = bya + (i / nbx) % nby, bz = bza + i / (nbx * nby);
                const uint32_t bkey = (uint32_t)bx | ((uint32_t)by << kGridBits) | ((uint32_t)bz << (2 * kGridBits));
                uint4          r;
                if (block_lookup(cv, bkey, r)) rec = r;
            }
            const uint64_t occ = ((uint64_t)rec.w << 32) | (uint64_t)rec.z;
            if (occ)
            {
                uint64_t want = expand_x(range4(box.lo[0] - 4 * bx, box.hi[0] - 4 * bx)) &
                                expand_y(range4(box.lo[1] - 4 * by, box.hi[1] - 4 * by)) &
                                expand_z(range4(box.lo[2] - 4 * bz, box.hi[2] - 4 * bz));
                if (skipping)
                    want &= ~(expand_x(range4(skip.lo[0] - 4 * bx, skip.hi[0] - 4 * bx)) &
                              expand_y(range4(skip.lo[1] - 4 * by, skip.hi[1] - 4 * by)) &
                              expand_z(range4(skip.lo[2] - 4 * bz, skip.hi[2] - 4 * bz)));
                uint64_t       w = occ & want;
                const uint64_t unw = occ & ~want;
                while (w)
                {
                    // a run: wanted cells up to the next occupied cell that is not wanted
                    const int      bit0 = __ffsll((long long)w) - 1;
                    const uint64_t below0 = (1ull << bit0) - 1ull;
                    const uint64_t above = unw & ~below0;
                    const uint64_t belowu = above ? ((1ull << (__ffsll((long long)above) - 1)) - 1ull) : ~0ull;
                    const uint32_t k0 = (uint32_t)__popcll(occ & below0), k1 = (uint32_t)__popcll(occ & belowu);
                    const uint32_t beg = __ldg(cv.fine_start + rec.y + k0), end = __ldg(cv.fine_start + rec.y + k1);
                    const uint32_t slot = atomicAdd(&S.nrng, 1u);
                    if (slot < (uint32_t)kSwRanges) S.rng[slot] = make_uint2(beg, end - beg);
                    mine += end - beg;
                    w &= ~belowu;
                }
            }
            // stop as soon as the budget is exceeded (uniform)
            if (i0 + 32 < nb && __reduce_add_sync(full, mine) > budget) return kSwOver;
        }
    }
    const uint32_t total = __reduce_add_sync(full, mine);
    __syncwarp();
    const uint32_t nr = S.nrng;
    if (nr > (uint32_t)kSwRanges || total > budget) return kSwOver;
    // run lengths -> exclusive prefix
    uint32_t run = 0;
    for (uint32_t b0 = 0; b0 < nr; b0 += 32)
    {
        const uint32_t i = b0 + lane;
        const uint32_t len = (i < nr) ? S.rng[i].y : 0u;
        uint32_t       inc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t v = __shfl_up_sync(full, inc, o);
            if (lane >= o) inc += v;
        }
        if (i < nr) S.rng[i].y = run + inc - len;
        run += __shfl_sync(full, inc, 31);
    }
    if (lane == 0) S.rng[nr].y = total;
    __syncwarp();
    return total;
}

// The lane's k best keys so far plus its list of candidates below the bound.
template <int K>
struct KeyEval
{
    float     ex, ey, ez;
    uint64_t* list;  // SweepSmem::list + lane
    uint64_t  key[K];
    int       n;
    // warp-collective: every lane folds its list into key[] (the bound key[K-1] tightens)
    __device__ __forceinline__ void fold()
    {
        const int nmax = __reduce_max_sync(0xFFFFFFFFu, n);
#pragma unroll 1
        for (int i = 0; i < nmax; i++)
        {
            uint64_t kk = (i < n) ? list[i * 32] : ~0ull;
            if (!(kk < key[K - 1])) kk = key[K - 1];  // re-inserting the last key changes nothing
            topk_insert<K>(key, kk);
        }
        n = 0;
    }
    // warp-collective: cnt (padded to a multiple of 4) staged points; all lists are folded as soon as one is full
    __device__ __forceinline__ void chunk(const float4* st, int cnt)
    {
#pragma unroll 1
        for (int j = 0; j < cnt; j += 4)
        {
            const float4   c0 = st[j], c1 = st[j + 1], c2 = st[j + 2], c3 = st[j + 3];
            const uint64_t k0 = make_key(dist2(ex, ey, ez, c0), __float_as_uint(c0.w));
            const uint64_t k1 = make_key(dist2(ex, ey, ez, c1), __float_as_uint(c1.w));
            const uint64_t k2 = make_key(dist2(ex, ey, ez, c2), __float_as_uint(c2.w));
            const uint64_t k3 = make_key(dist2(ex, ey, ez, c3), __float_as_uint(c3.w));
            const uint64_t w = key[K - 1];
            if (k0 < w) list[(n++) * 32] = k0;
            if (k1 < w) list[(n++) * 32] = k1;
            if (k2 < w) list[(n++) * 32] = k2;
            if (k3 < w) list[(n++) * 32] = k3;
            if (__any_sync(0xFFFFFFFFu, n > kSwList - 4)) fold();
        }
    }
};

// Warp-collective.  Streams the `total` points of the runs listed in S.rng
// through the stage into ev; with `filter` only the points inside [blo, bhi].
template <class Ev>
__device__ __forceinline__ void sweep_stream(const CloudView& cv, SweepSmem& S, uint32_t total, bool filter,
                                             const float (&blo)[3], const float (&bhi)[3], Ev& ev)
{
    const int      lane = threadIdx.x & 31;
    const unsigned full = 0xFFFFFFFFu, lt = (1u << lane) - 1u;
    const uint32_t nr = S.nrng;
    uint32_t       top = 1;  // largest power of two below nr (first step of the search over the prefix)
    while (2 * top < nr) top *= 2;
    // padding: infinitely far, and its key equals the sentinel of an uncapped search (never below any bound)
    const float4 far4 = make_float4(INFINITY, INFINITY, INFINITY, __uint_as_float(0xFFFFFFFFu));
    int          cnt = 0;
#pragma unroll 1
    for (uint32_t t0 = 0; t0 < total; t0 += 128)
    {
        uint32_t tt[4], r[4];
#pragma unroll
        for (int u = 0; u < 4; u++) tt[u] = min(t0 + 32 * u + lane, total - 1u), r[u] = 0;
        // last run whose prefix is <= tt, four searches interleaved
        for (uint32_t step = top; step; step >>= 1)
        {
#pragma unroll
            for (int u = 0; u < 4; u++)
            {
                const uint32_t c = r[u] + step;
                if (c < nr && S.rng[c].y <= tt[u]) r[u] = c;
            }
        }
        float4 p[4];
#pragma unroll
        for (int u = 0; u < 4; u++)
        {
            const uint2 rg = S.rng[r[u]];
            p[u] = __ldg(cv.pts + rg.x + (tt[u] - rg.y));
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
        {
            bool keep = t0 + 32 * u + lane < total;
            if (filter)
                keep = keep && p[u].x >= blo[0] && p[u].x <= bhi[0] && p[u].y >= blo[1] && p[u].y <= bhi[1] &&
                       p[u].z >= blo[2] && p[u].z <= bhi[2];
            const unsigned m = __ballot_sync(full, keep);
            if (keep) S.stage[cnt + __popc(m & lt)] = p[u];
            cnt += __popc(m);
        }
        if (cnt > kSwStage - 128 || t0 + 128 >= total)
        {
            if (lane < 4) S.stage[cnt + lane] = far4;  // pad to the evaluator's stride
            __syncwarp();
            B2_COUNT(9, cnt);
            ev.chunk(S.stage, cnt);
            cnt = 0;
            __syncwarp();
        }
    }
}

// ---- the same stream with the bulk-copy engine (B200ICP_TMA=1) ----------------
// A run is a contiguous piece of the sorted point array, so it can go to shared
// memory as ONE cp.async.bulk (1-D TMA) instead of per-thread loads: every lane
// issues the copies of a few runs into one half of the stage, the half's
// mbarrier counts the bytes (expect_tx / complete_tx), and the warp evaluates
// one half while the copies into the other are in flight.  Whole runs land in
// the stage -- there is no per-point box filter on this path -- which is still
// a superset of every lane's ball, so the keys are the same.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "B2_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra B2_WAIT_%=;\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <class Ev>
__device__ __forceinline__ void sweep_stream_bulk(const CloudView& cv, SweepSmem& S, uint32_t total, uint32_t& par,
                                                  Ev& ev)
{
    constexpr uint32_t H = kSwStage / 2;
    const int          lane = threadIdx.x & 31;
    const uint32_t     nr = S.nrng;
    const uint32_t     nchunks = (total + H - 1) / H;
    const float4       far4 = make_float4(INFINITY, INFINITY, INFINITY, __uint_as_float(0xFFFFFFFFu));
    auto issue = [&](uint32_t c) {
        const uint32_t h = c & 1u, base = c * H, lim = min(base + H, total);
        float4*        dst = S.stage + h * (H + 4);
        // the half was read with ordinary loads by the previous evaluation: order them before the async writes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (lane == 0) mbar_expect_tx(&S.mbar[h], (lim - base) * 16u);
        for (uint32_t i = lane; i < nr; i += 32)
        {
            const uint2    rg = S.rng[i];
            const uint32_t end = S.rng[i + 1].y;  // prefix of the next run = end of this one
            const uint32_t lo = max(rg.y, base), hi = min(end, lim);
            if (lo < hi) bulk_g2s(dst + (lo - base), cv.pts + rg.x + (lo - rg.y), (hi - lo) * 16u, &S.mbar[h]);
        }
    };
    issue(0);
#pragma unroll 1
    for (uint32_t c = 0; c < nchunks; c++)
    {
        const uint32_t h = c & 1u;
        if (c + 1 < nchunks) issue(c + 1);
        mbar_wait(&S.mbar[h], (par >> h) & 1u);
        par ^= 1u << h;
        const int cnt = (int)(min((c + 1) * H, total) - c * H);
        float4*   st = S.stage + h * (H + 4);
        if (lane < 4) st[cnt + lane] = far4;  // pad to the evaluator's stride
        __syncwarp();
        B2_COUNT(9, cnt);
        ev.chunk(st, cnt);
        __syncwarp();
    }
}

// Warp-collective: the search of the lanes with `in` set.  On entry key[] holds
// sentinel_key(cap) on those lanes.  `bounded`: the lanes' caps already are
// bounds of their k-th neighbour (seeds) -- no own-cells pass.  Returns 0 when
// every `in` lane is final, 1 when the lanes are too far apart for one box
// (nothing changed: the caller splits the group).
template <int K>
__device__ __forceinline__ int item_sweep(const CloudView& cv, const GridDev& g, SweepSmem& SW, bool in,
                                          bool bounded, bool bulk, uint32_t& par, float qx, float qy, float qz,
                                          float& cap, uint64_t (&key)[K], uint64_t& sent)
{
    B2_PHASE_DECL;
    const int      lane = threadIdx.x & 31;
    const unsigned full = 0xFFFFFFFFu;
    KeyEval<K>     ke;
    ke.ex = in ? qx : INFINITY, ke.ey = in ? qy : INFINITY, ke.ez = in ? qz : INFINITY;
    ke.list = SW.list + lane;
    ke.n = 0;
#pragma unroll
    for (int i = 0; i < K; i++) ke.key[i] = key[i];
    const float q[3] = {qx, qy, qz};
    const float o[3] = {g.ox, g.oy, g.oz};
    // box of the group's queries (float), shared by every pass
    float qlo[3], qhi[3];
#pragma unroll
    for (int d = 0; d < 3; d++)
    {
        qlo[d] = dec_order_f(__reduce_min_sync(full, in ? enc_order_f(q[d]) : 0xFFFFFFFFu));
        qhi[d] = dec_order_f(__reduce_max_sync(full, in ? enc_order_f(q[d]) : 0u));
    }
    CellBox skip;
#pragma unroll
    for (int d = 0; d < 3; d++) skip.lo[d] = 1, skip.hi[d] = 0;
    int   phase = bounded ? 1 : 0;  // 0: the queries' own cells; 1, 2, 3: the grown box for all / narrower lane sets
    bool  done = false;
    float r = 0.f, rall = 0.f;
    if (phase == 1)
    {
        r = in ? sqrtf(cap) : 0.0f;
        rall = __uint_as_float(__reduce_max_sync(full, __float_as_uint(r)));
    }
#pragma unroll 1
    for (;;)
    {
        bool    cov = in;
        float   blo[3], bhi[3];
        CellBox box;
        if (phase == 0)
        {
#pragma unroll
            for (int d = 0; d < 3; d++) blo[d] = qlo[d], bhi[d] = qhi[d];
        }
        else
        {
            const float thr = (phase == 1) ? rall : rall * (phase == 2 ? 0.5f : 0.25f);
            cov = in && r <= thr;
            if (__ballot_sync(full, cov) == 0u)
            {
                if (++phase > 3) break;
                continue;
            }
            const float rm = __uint_as_float(__reduce_max_sync(full, cov ? __float_as_uint(r) : 0u));
#pragma unroll
            for (int d = 0; d < 3; d++)
            {
                float lo = qlo[d], hi = qhi[d];
                if (phase > 1)
                {
                    lo = dec_order_f(__reduce_min_sync(full, cov ? enc_order_f(q[d]) : 0xFFFFFFFFu));
                    hi = dec_order_f(__reduce_max_sync(full, cov ? enc_order_f(q[d]) : 0u));
                }
                // grown by the largest covered radius, padded for the rounding of d2 and of the corners
                const float pad = rm * 1.001f + 1e-6f * (fabsf(lo) + fabsf(hi) + 1.0f);
                blo[d] = lo - pad, bhi[d] = hi + pad;
            }
        }
        // cells of the box: the expression the index sorted the points by (monotone in the coordinate), so a
        // point inside [blo, bhi] is in a cell inside `box`
#pragma unroll
        for (int d = 0; d < 3; d++)
        {
            box.lo[d] = sweep_fine_coord(blo[d], o[d], g.inv_cell);
            box.hi[d] = sweep_fine_coord(bhi[d], o[d], g.inv_cell);
        }
        const uint32_t total = sweep_enumerate(cv, SW, box, skip, (uint32_t)kSwBudget);
        B2_PHASE(3);
        B2_COUNT(11, 1);
        if (total == kSwWide && phase <= 1) return 1;  // key[] is untouched: the caller splits the group
        if (total == kSwOver || total == kSwWide)
        {  // too many points: without the own-cells pass, or leaving the lanes with the widest bounds out
            if (phase == 0)
            {
                phase = 1;
                r = in ? sqrtf(cap) : 0.0f;
                rall = __uint_as_float(__reduce_max_sync(full, __float_as_uint(r)));
                continue;
            }
            if (++phase > 3) break;
            continue;
        }
        if (total)
        {
            if (bulk)
                sweep_stream_bulk(cv, SW, total, par, ke);
            else
                sweep_stream(cv, SW, total, phase != 0, blo, bhi, ke);
        }
        ke.fold();
        B2_PHASE(4);
        B2_COUNT(8, total);
        if (phase == 0)
        {  // every point of these cells has been seen: later passes skip them; bounds from what was found
            skip = box;
            phase = 1;
            r = in ? sqrtf(key_d2(ke.key[K - 1])) : 0.0f;
            rall = __uint_as_float(__reduce_max_sync(full, __float_as_uint(r)));
            continue;
        }
        done = cov;
        break;
    }
    if (in)
    {
        if (done)
        {
#pragma unroll
            for (int i = 0; i < K; i++) key[i] = ke.key[i];
        }
        else
        {  // the keys found so far are real points: k of them bound the k-th
            if (ke.key[K - 1] != sent) cap = key_d2(ke.key[K - 1]);
            sent = sentinel_key(cap);
#pragma unroll
            for (int i = 0; i < K; i++) key[i] = sent;
        }
    }
    // the lanes left out (wide bounds in a dense place) search alone, from the bound the sweep found
    B2_COUNT(10, __popc(__ballot_sync(full, in && !done)));
    if (in && !done) knn_search<K>(cv, g, qx, qy, qz, cap, key);
    __syncwarp();
    B2_PHASE(5);
    return 0;
}

}  // namespace b2
