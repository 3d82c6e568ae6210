// item_sweep.cuh -- the converged part of the search stage: one warp SWEEPS the
// points around its item instead of every lane walking shells on its own.
//
// An item is 32 spatially compact queries (one per lane, tile_search.cuh).  Each
// lane owns an upper bound `cap` of its k-th neighbour's squared distance:
//   * seeded    -- the k neighbours of the previous outer iteration, re-measured
//                  under the new pose (align.cu, seeded_cap), or
//   * round 0   -- the k-th smallest distance to the points of the cells the
//                  item's queries fall in (a first, cell-granular sweep that only
//                  keeps k float distances per lane, branch-free).
// The true k nearest of lane q then lie inside the ball (q, sqrt(cap_q)), so the
// warp takes the box of its queries grown by the largest cap, lists the runs of
// sorted points of the index cells that box touches (lane = block of the tile,
// consecutive occupied cells of a block are one contiguous run), streams them
// with coalesced float4 loads (every lane a different point of the flattened run
// list, four loads in flight), keeps the points inside the box in a shared-memory
// stage and lets ALL lanes evaluate EVERY staged point: no divergence, broadcast
// shared-memory reads.  A lane appends the keys (d2 bits << 32 | index) below its
// bound to a private list in shared memory (one predicated store); lists are
// folded into the lane's k best keys when they fill up and at the end -- the
// result is the k smallest keys, which is the tie rule of Appendix A.4, whatever
// order the candidates arrive in.
//
// Lanes whose cap is much larger than their neighbours' (fewer than k points
// within the radius last time) would blow the box up: when the box holds more
// than kSwBudget points the sweep is retried for the lanes with a cap below
// 1/2, 1/4 of the largest; lanes left out take the per-lane shell walk
// (tile_search.cuh) afterwards, bounded by whatever the sweep found for them.
// Either way the keys equal those of knn_search<K> bit for bit.
#pragma once
#include "sweep_search.cuh"
#include "tile_search.cuh"

namespace b2
{
// development probe (-DB200ICP_DBG_PHASES): cycles per phase of the search, summed over warps
#ifdef B200ICP_DBG_PHASES
__device__ unsigned long long g_dbg_phase[16];
#define B2_PHASE_DECL long long b2_ph_t = clock64()
#define B2_PHASE(i)                                                                       \
    do                                                                                    \
    {                                                                                     \
        const long long b2_now = clock64();                                               \
        if ((threadIdx.x & 31) == 0) atomicAdd(&g_dbg_phase[i], (unsigned long long)(b2_now - b2_ph_t)); \
        b2_ph_t = b2_now;                                                                 \
    } while (0)
#define B2_COUNT(i, v)                                                                    \
    do                                                                                    \
    {                                                                                     \
        const unsigned long long b2_v = (unsigned long long)(v);                          \
        if ((threadIdx.x & 31) == 0) atomicAdd(&g_dbg_phase[i], b2_v);                    \
    } while (0)
#else
#define B2_PHASE_DECL
#define B2_PHASE(i)
#define B2_COUNT(i, v)
#endif

constexpr int kSwRanges = 192;    // runs of sorted points per sweep
constexpr int kSwStage = 256;     // staged points per evaluation chunk
constexpr int kSwList = 16;       // keys a lane may collect between two folds
constexpr int kSwBudget = 2048;   // points one sweep may load
constexpr int kSwAttempts = 3;

struct SweepSmem
{
    uint2    rng[kSwRanges + 1];  // (first sorted position, exclusive prefix of the run lengths); [n].y = total
    uint32_t nrng;
    uint32_t pad;
    float4   stage[kSwStage + 4];
    uint64_t list[kSwList * 32];  // [slot][lane]
};

constexpr uint32_t kSwOver = 0xFFFFFFFFu;

// Warp-collective.  Lists the runs of points of the tile cells inside the box
// c0..c1 (tile coordinates, inclusive) into S.rng and returns their total
// length, or kSwOver when there are more than `budget` points / kSwRanges runs.
__device__ __forceinline__ uint32_t sweep_enumerate(const WarpTile& W, const TileGeom& G, const CloudView& cv,
                                                    SweepSmem& S, const int (&c0)[3], const int (&c1)[3],
                                                    uint32_t budget)
{
    const int      lane = threadIdx.x & 31;
    const unsigned full = 0xFFFFFFFFu;
    if (lane == 0) S.nrng = 0u;
    __syncwarp();
    uint32_t mine = 0;
    if (c0[0] <= c1[0] && c0[1] <= c1[1] && c0[2] <= c1[2])
    {
        // absolute cell coordinates of the box and the blocks it touches
        const int ax0 = G.t0x + c0[0], ax1 = G.t0x + c1[0];
        const int ay0 = G.t0y + c0[1], ay1 = G.t0y + c1[1];
        const int az0 = G.t0z + c0[2], az1 = G.t0z + c1[2];
        const int bxa = (ax0 >> 2) - G.bx0, nbx = ((ax1 >> 2) - G.bx0) - bxa + 1;
        const int bya = (ay0 >> 2) - G.by0, nby = ((ay1 >> 2) - G.by0) - bya + 1;
        const int bza = (az0 >> 2) - G.bz0, nbz = ((az1 >> 2) - G.bz0) - bza + 1;
        const int nb = nbx * nby * nbz;
        for (int i0 = 0; i0 < nb; i0 += 32)
        {
            const int i = i0 + lane;
            if (i >= nb) continue;
            const int      bxi = bxa + i % nbx, byi = bya + (i / nbx) % nby, bzi = bza + i / (nbx * nby);
            const int      b = bxi + G.nbx * (byi + G.nby * bzi);
            const uint4    rec = W.rec[b];
            const uint64_t occ = ((uint64_t)rec.w << 32) | (uint64_t)rec.z;
            if (occ == 0) continue;
            const int      bx = G.bx0 + bxi, by = G.by0 + byi, bz = G.bz0 + bzi;
            const uint64_t want = expand_x(range4(ax0 - 4 * bx, ax1 - 4 * bx)) &
                                  expand_y(range4(ay0 - 4 * by, ay1 - 4 * by)) &
                                  expand_z(range4(az0 - 4 * bz, az1 - 4 * bz));
            uint64_t       w = occ & want;
            const uint64_t unw = occ & ~want;
            const uint32_t so = W.fsoff[b];
            while (w)
            {
                // a run: wanted cells up to the next occupied cell that is not wanted
                const int      bit0 = __ffsll((long long)w) - 1;
                const uint64_t below0 = (1ull << bit0) - 1ull;
                const uint64_t above = unw & ~below0;
                const uint64_t belowu = above ? ((1ull << (__ffsll((long long)above) - 1)) - 1ull) : ~0ull;
                const uint32_t k0 = (uint32_t)__popcll(occ & below0), k1 = (uint32_t)__popcll(occ & belowu);
                uint32_t       beg, end;
                if (so != kFsNone)
                    beg = W.fs[so + k0], end = W.fs[so + k1];
                else
                    beg = __ldg(cv.fine_start + rec.y + k0), end = __ldg(cv.fine_start + rec.y + k1);
                const uint32_t slot = atomicAdd(&S.nrng, 1u);
                if (slot < (uint32_t)kSwRanges) S.rng[slot] = make_uint2(beg, end - beg);
                mine += end - beg;
                w &= ~belowu;
            }
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

// K smallest distances, nothing else (round 0): a min/max chain, no branch
template <int K>
struct BoundEval
{
    float ex, ey, ez;
    float best[K];
    __device__ __forceinline__ void chunk(const float4* st, int cnt)
    {
        for (int j = 0; j < cnt; j += 2)
        {
            const float4 c0 = st[j], c1 = st[j + 1];
            float        d0 = dist2(ex, ey, ez, c0), d1 = dist2(ex, ey, ez, c1);
#pragma unroll
            for (int i = 0; i < K; i++)
            {
                const float lo = fminf(best[i], d0);
                d0 = fmaxf(best[i], d0);
                best[i] = lo;
            }
#pragma unroll
            for (int i = 0; i < K; i++)
            {
                const float lo = fminf(best[i], d1);
                d1 = fmaxf(best[i], d1);
                best[i] = lo;
            }
        }
    }
};

// keys below the lane's bound go to its list; full lists are folded into key[]
template <int K>
struct KeyEval
{
    float     ex, ey, ez;
    uint64_t* list;  // SweepSmem::list + lane
    uint64_t  key[K];
    int       n;
    __device__ __forceinline__ void fold()
    {
        for (int i = 0; i < n; i++)
        {
            const uint64_t kk = list[i * 32];
            if (kk < key[K - 1]) topk_insert<K>(key, kk);
        }
        n = 0;
    }
    __device__ __forceinline__ void chunk(const float4* st, int cnt)
    {
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
            if (n > kSwList - 4) fold();
        }
    }
};

// Warp-collective.  Streams the `total` points of the runs listed in S.rng
// through the stage; with FILTER only the points inside [blo, bhi] are staged.
template <bool FILTER, class Ev>
__device__ __forceinline__ void sweep_stream(const CloudView& cv, SweepSmem& S, uint32_t total, const float (&blo)[3],
                                             const float (&bhi)[3], Ev& ev)
{
    const int      lane = threadIdx.x & 31;
    const unsigned full = 0xFFFFFFFFu, lt = (1u << lane) - 1u;
    const uint32_t nr = S.nrng;
    // padding: infinitely far, and its key equals the sentinel of an uncapped search (never below any bound)
    const float4   far4 = make_float4(INFINITY, INFINITY, INFINITY, __uint_as_float(0xFFFFFFFFu));
    int            cnt = 0;
    for (uint32_t t0 = 0; t0 < total; t0 += 128)
    {
        float4 p[4];
        bool   ok[4];
#pragma unroll
        for (int u = 0; u < 4; u++)
        {
            const uint32_t t = t0 + 32 * u + lane;
            ok[u] = t < total;
            const uint32_t tt = min(t, total - 1u);
            uint32_t       r = 0;  // last run whose prefix is <= tt
#pragma unroll
            for (uint32_t step = 128; step; step >>= 1)
            {
                const uint32_t c = r + step;
                if (c < nr && S.rng[c].y <= tt) r = c;
            }
            const uint2 rg = S.rng[r];
            p[u] = __ldg(cv.pts + rg.x + (tt - rg.y));
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
        {
            bool keep = ok[u];
            if (FILTER)
                keep = keep && p[u].x >= blo[0] && p[u].x <= bhi[0] && p[u].y >= blo[1] && p[u].y <= bhi[1] &&
                       p[u].z >= blo[2] && p[u].z <= bhi[2];
            const unsigned m = __ballot_sync(full, keep);
            if (keep) S.stage[cnt + __popc(m & lt)] = p[u];
            cnt += __popc(m);
            if (FILTER) B2_COUNT(9, __popc(m));
        }
        if (cnt > kSwStage - 128)
        {
            if (lane < 4) S.stage[cnt + lane] = far4;  // pad to the evaluators' stride
            __syncwarp();
            ev.chunk(S.stage, cnt);
            cnt = 0;
            __syncwarp();
        }
    }
    if (lane < 4) S.stage[cnt + lane] = far4;
    __syncwarp();
    ev.chunk(S.stage, cnt);
    __syncwarp();
}

// Warp-collective; the tile (W, G) is built for the lanes with `in` set and S
// shells.  On entry key[] = sentinel(cap) on every lane.  Returns true on the
// lanes whose keys are final; the others hold key[] = sentinel(cap) again, with
// cap possibly tightened, and still have to search.
template <int K>
__device__ __forceinline__ bool item_sweep(const WarpTile& W, const TileGeom& G, const CloudView& cv, const GridDev& g,
                                           SweepSmem& SW, int S, bool in, bool need_bound, float qx, float qy,
                                           float qz, float& cap, uint64_t (&key)[K], uint64_t& sent)
{
    B2_PHASE_DECL;
    const int      lane = threadIdx.x & 31;
    const unsigned full = 0xFFFFFFFFu;
    const float    ex = in ? qx : INFINITY, ey = in ? qy : INFINITY, ez = in ? qz : INFINITY;
    const float    none[3] = {0.f, 0.f, 0.f};
    // ---- round 0: a bound from the points of the queries' own cells -----------
    if (need_bound)
    {
        const int      c0[3] = {S, S, S}, c1[3] = {G.nx - 1 - S, G.ny - 1 - S, G.nz - 1 - S};
        const uint32_t total = sweep_enumerate(W, G, cv, SW, c0, c1, (uint32_t)kSwBudget);
        if (total != kSwOver && total >= (uint32_t)K)
        {
            BoundEval<K> be;
            be.ex = ex, be.ey = ey, be.ez = ez;
#pragma unroll
            for (int i = 0; i < K; i++) be.best[i] = INFINITY;
            sweep_stream<false>(cv, SW, total, none, none, be);
            if (in && be.best[K - 1] < cap)
            {
                cap = be.best[K - 1];
                sent = sentinel_key(cap);
#pragma unroll
                for (int i = 0; i < K; i++) key[i] = sent;
            }
        }
        B2_PHASE(2);
    }
    // ---- the sweep -------------------------------------------------------------
    const float    r = in ? sqrtf(cap) : 0.0f;
    const float    rall = __uint_as_float(__reduce_max_sync(full, __float_as_uint(r)));
    bool           done = false;
    for (int att = 0; att < kSwAttempts; att++)
    {
        const float thr = (att == 0) ? rall : rall * (att == 1 ? 0.5f : 0.25f);
        const bool  cov = in && r <= thr;
        if (__ballot_sync(full, cov) == 0u) break;
        const float rm = __uint_as_float(__reduce_max_sync(full, cov ? __float_as_uint(r) : 0u));
        const float q[3] = {qx, qy, qz};
        float       blo[3], bhi[3];
        int         c0[3], c1[3];
        const float o[3] = {g.ox, g.oy, g.oz};
        const int   t0[3] = {G.t0x, G.t0y, G.t0z}, nn[3] = {G.nx, G.ny, G.nz};
#pragma unroll
        for (int d = 0; d < 3; d++)
        {
            const float lo = dec_order_f(__reduce_min_sync(full, cov ? enc_order_f(q[d]) : 0xFFFFFFFFu));
            const float hi = dec_order_f(__reduce_max_sync(full, cov ? enc_order_f(q[d]) : 0u));
            // grown by the largest covered radius, padded for the rounding of d2 and of the corners
            const float pad = rm * 1.001f + 1e-6f * (fabsf(lo) + fabsf(hi) + 1.0f);
            blo[d] = lo - pad, bhi[d] = hi + pad;
            // cells of the box: the expression the index sorted the points by (monotone), cut to the tile --
            // every point within a covered lane's cap lies within S shells of its home cell, i.e. inside the tile
            c0[d] = max(sweep_fine_coord(blo[d], o[d], g.inv_cell) - t0[d], 0);
            c1[d] = min(sweep_fine_coord(bhi[d], o[d], g.inv_cell) - t0[d], nn[d] - 1);
        }
        const uint32_t total = sweep_enumerate(W, G, cv, SW, c0, c1, (uint32_t)kSwBudget);
        B2_PHASE(3);
        B2_COUNT(11, 1);
        if (total == kSwOver) continue;  // too many points: leave the lanes with the widest caps out
        KeyEval<K> ke;
        ke.ex = ex, ke.ey = ey, ke.ez = ez;
        ke.list = SW.list + lane;
        ke.n = 0;
#pragma unroll
        for (int i = 0; i < K; i++) ke.key[i] = sent;
        if (total) sweep_stream<true>(cv, SW, total, blo, bhi, ke);
        ke.fold();
        if (in)
        {
#pragma unroll
            for (int i = 0; i < K; i++) key[i] = ke.key[i];
        }
        done = cov;
        B2_PHASE(4);
        B2_COUNT(8, total);
        B2_COUNT(10, __popc(__ballot_sync(full, in && !cov)));
        break;
    }
    if (in && !done)
    {  // the walk restarts from the bound the sweep found (the keys are real points: k of them bound the k-th)
        if (key[K - 1] != sent) cap = key_d2(key[K - 1]);
        sent = sentinel_key(cap);
#pragma unroll
        for (int i = 0; i < K; i++) key[i] = sent;
    }
    return done;
}

}  // namespace b2
