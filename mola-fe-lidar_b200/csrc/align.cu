// align.cu -- the registration loop of mp2p_icp::ICP::align as the reference
// drives it (LidarOdometry.cpp:869-871),
// resident on the device for the whole iteration loop.
//
// Per outer iteration, two launches over a table of independent jobs:
//   search kernel : per local point -- transform (A.2), radius-capped exact
//                   kNN on the grid index (A.3/A.4): the per-lane shell walk of
//                   tile_search.cuh over 32-query items, scheduled dynamically
//                   (results do not depend on the order), from the second
//                   iteration on bounded by the previous neighbours re-measured
//                   under the new pose -> the neighbour rows nn[job][point][K].
//   fit kernel    : per local point -- plane fit + gates (A.5) and the
//                   point-to-plane MOMENTS of the pairing, accumulated per chunk
//                   of items and reduced in a fixed tree over the data; the warp
//                   that completes a job's reduction runs the whole Gauss-Newton
//                   inner loop (A.6) on the 12x12 moment matrix, the SE(3)
//                   update, the convergence test (A.7) and sets the job's status
//                   flags -- no separate solver launch, no host round trip.
//
// Why moments: the point-to-plane residual r_i(T) = n_i.(R p_i + t - c_i) is
// LINEAR in theta = (R row-major | t) interleaved as 3 rows of (R_i0 R_i1 R_i2 t_i):
// r_i(T) = r_i(T0) + a_i.(theta - theta0), a_i = n_i (x) [p_i;1].  So
//   H = J^T (sum a a^T) J,  g = J^T (sum a r0 + (sum a a^T)(theta - theta0))
// for every inner GN iterate, with J = d theta / d eps (12x6) of the right
// perturbation T (+) exp(eps).  One pass over the pairings per OUTER iteration
// instead of one per inner iteration; the iterates equal the reference's
// per-pairing Gauss-Newton in exact arithmetic.
#include "icp_math.cuh"
#include "tile_search.cuh"
#include "sweep_search.cuh"
#include "item_sweep.cuh"
#include "runtime.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <map>

namespace b2
{
// ---- moments ---------------------------------------------------------------
// Every pairing contributes e e^T with the 16-vector
//   e = [ a (12) | r0 | 1 | 0 | 0 ],  a[4 i + j] = n_i * h_j,  h = (p_local, 1)
// so S = sum e e^T (16x16, symmetric) holds A = sum a a^T (12x12), sum a r0
// (column 12), sum r0^2 (S[12][12]) and the pairing count (S[13][13]).  S is
// accumulated on the FP64 tensor-core path (DMMA m8n8k4): per 4 pairings the
// warp loads two fragments from its staging buffer and issues three MMAs for
// the tiles C00 = E[0:8] E[0:8]^T, C01 = E[0:8] E[8:16]^T, C11 = E[8:16] E[8:16]^T
// (C10 = C01^T is not needed).  The accumulators stay in registers over all the
// items a CTA processes; the order of accumulation is fixed by the static
// item -> CTA assignment, so runs are bit-reproducible.
constexpr int kStageStride = 20;  // doubles per staged row: conflict-free fragment loads

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// element (a, b) of S from the three stored tiles (64 doubles each, row-major)
__host__ __device__ __forceinline__ double moment_at(const double* M, int a, int b)
{
    if (a > b)
    {
        const int t = a;
        a = b, b = t;
    }
    if (b < 8) return M[a * 8 + b];
    if (a < 8) return M[64 + a * 8 + (b - 8)];
    return M[128 + (a - 8) * 8 + (b - 8)];
}

// A = sum_i a_i a_i^T (12x12) and the pairing count out of S, for both matchers:
// point-to-plane stores it directly, point-to-point has A = I3 (x) sum h h^T
__host__ __device__ __forceinline__ double normal_matrix_at(const double* M, bool p2p, int a, int b)
{
    if (!p2p) return moment_at(M, a, b);
    return ((a >> 2) == (b >> 2)) ? moment_at(M, a & 3, b & 3) : 0.0;
}
__host__ __device__ __forceinline__ double pairing_count(const double* M, bool p2p)
{
    return p2p ? moment_at(M, 3, 3) : moment_at(M, 13, 13);
}

struct MatchOut
{
    uint8_t*  paired;
    uint32_t* nn_idx;
    uint32_t* nn_cnt;
    double*   centroid;
    double*   normal;
};

// A.2: q = fl32(R p + t), f64 accumulate in this fixed order (both stages use it)
__device__ __forceinline__ void transform_point(const double* Rt, const float4& pl, double& gx, double& gy,
                                                double& gz)
{
    const double px = pl.x, py = pl.y, pz = pl.z;
    gx = ((Rt[0] * px + Rt[1] * py) + Rt[2] * pz) + Rt[9];
    gy = ((Rt[3] * px + Rt[4] * py) + Rt[5] * pz) + Rt[10];
    gz = ((Rt[6] * px + Rt[7] * py) + Rt[8] * pz) + Rt[11];
}

__device__ __forceinline__ bool matcher_active(const IcpDevParams& P, uint32_t it)
{
    return (P.run_from_iteration <= it) && (P.run_up_to_iteration == 0 || it <= P.run_up_to_iteration);
}

// ============================================================== search stage
// What a search kernel does with the k best keys of one item.  Called by all 32
// lanes of one warp: lane = query, `has` false on lanes without a point.

// neighbour indices (and squared distances) to a buffer
// kNN call (by_orig): row = the query's original index, entries = ORIGINAL
// indices of the neighbours.  Matchers: row = job base + the query's sorted
// position, entries = SORTED POSITIONS of the neighbours in the global cloud --
// the fit stage and the next iteration's seeded search read pts[position]
// directly.
struct NnWriter
{
    uint32_t* idx;      // [rows * k]
    float*    d2;       // [rows * k] or null
    uint32_t  k;        // entries per row (<= K)
    uint32_t  by_orig;
    template <int K>
    __device__ __forceinline__ void operator()(JobDev& J, const CloudView& cvG, bool has, uint32_t pos,
                                               uint32_t orig, const uint64_t (&key)[K], uint64_t sent) const
    {
        if (!has) return;
        const size_t row = by_orig ? (size_t)orig : (size_t)J.pair_base + pos;
#pragma unroll
        for (int i = 0; i < K; i++)
            if ((uint32_t)i < k)
            {
                const bool ok = key[i] != sent;
                uint32_t   v = kInvalid;
                if (ok) v = by_orig ? key_idx(key[i]) : __ldg(cvG.rank + key_idx(key[i]));
                idx[row * k + i] = v;
                if (d2) d2[row * k + i] = ok ? key_d2(key[i]) : INFINITY;
            }
    }
};

// packed keys (d2 bits << 32 | index) for the multi-GPU arg-min merge; `map`
// renumbers shard-local indices to the caller's global ones
struct KeyWriter
{
    uint64_t*       keys;    // [rows * k], row = original index of the query (by_pos: its sorted position)
    const uint32_t* map;     // or null
    uint32_t        k;
    uint32_t        by_pos;  // sharded registration: rows in the local cloud's sorted order, as the fit stage reads them
    template <int K>
    __device__ __forceinline__ void operator()(JobDev&, const CloudView&, bool has, uint32_t pos, uint32_t orig,
                                               const uint64_t (&key)[K], uint64_t sent) const
    {
        if (!has) return;
        if (by_pos) orig = pos;
#pragma unroll
        for (int i = 0; i < K; i++)
            if ((uint32_t)i < k)
            {
                uint64_t v = B200ICP_NO_KEY;
                if (key[i] != sent)
                {
                    const uint32_t li = key_idx(key[i]);
                    v = (key[i] & 0xFFFFFFFF00000000ull) | (uint64_t)(map ? __ldg(map + li) : li);
                }
                keys[(size_t)orig * k + i] = v;
            }
    }
};

// The same keys stored into the gather buffers of all ranks of a sharded map
// (SURVEY 8e): peer[r] is rank r's buffer mapped over NVLink.  Rows go to
// [rank][query][k] of every buffer (search + all-gather in one kernel); with
// `amin` (k = 1) the key is folded into slot [query] of every buffer with a
// system-scope atomicMin (search + all-reduce(MIN) in one kernel).
// `per` > 0 (reduce-scatter form): the row of query q goes ONLY to the rank
// that owns q (owner = q / per), at [rank][q - owner * per][k] of its buffer;
// the owner merges its slice and broadcasts the merged rows
// (merge_scatter_kernel) -- 2 x nq x k keys of traffic per rank instead of
// world x nq x k.
struct KeyScatter
{
    uint64_t*       peer[8];
    const uint32_t* map;
    uint32_t        k, world, rank, nq, amin, per;
    template <int K>
    __device__ __forceinline__ void operator()(JobDev&, const CloudView&, bool has, uint32_t, uint32_t orig,
                                               const uint64_t (&key)[K], uint64_t sent) const
    {
        if (!has) return;
        uint64_t v[K];
#pragma unroll
        for (int i = 0; i < K; i++)
        {
            v[i] = B200ICP_NO_KEY;
            if ((uint32_t)i < k && key[i] != sent)
            {
                const uint32_t li = key_idx(key[i]);
                v[i] = (key[i] & 0xFFFFFFFF00000000ull) | (uint64_t)(map ? __ldg(map + li) : li);
            }
        }
        if (amin)
        {
            if (v[0] == B200ICP_NO_KEY) return;
            if (per)
            {  // only the owner of the query keeps the minimum; it broadcasts its slice afterwards
                const uint32_t owner = orig / per;
                atomicMin_system(reinterpret_cast<unsigned long long*>(peer[owner]) + (orig - owner * per),
                                 (unsigned long long)v[0]);
                return;
            }
            for (uint32_t r = 0; r < world; r++)
                atomicMin_system(reinterpret_cast<unsigned long long*>(peer[r]) + orig, (unsigned long long)v[0]);
            return;
        }
        if (per)
        {
            const uint32_t owner = orig / per;
            uint64_t*      dst = peer[owner] + ((size_t)rank * per + (orig - owner * per)) * k;
            store_row<K>(dst, v);
            return;
        }
        for (uint32_t r = 0; r < world; r++) store_row<K>(peer[r] + ((size_t)rank * nq + orig) * k, v);
    }
    // k keys to dst: 16-byte stores when the row allows it (k even)
    template <int K>
    __device__ __forceinline__ void store_row(uint64_t* dst, const uint64_t (&v)[K]) const
    {
        if ((k & 1u) == 0 && (K & 1) == 0)
        {
#pragma unroll
            for (int i = 0; i < K; i += 2)
                if ((uint32_t)i < k)
                    *reinterpret_cast<ulonglong2*>(dst + i) = make_ulonglong2(v[i], v[i + 1]);
        }
        else
        {
#pragma unroll
            for (int i = 0; i < K; i++)
                if ((uint32_t)i < k) dst[i] = v[i];
        }
    }
};

// QualityEvaluator_PairedRatio (row O / A.8): queries with a neighbour at d2 < thr2 (strict)
struct HitCounter
{
    float thr2;
    template <int K>
    __device__ __forceinline__ void operator()(JobDev& J, const CloudView&, bool has, uint32_t, uint32_t,
                                               const uint64_t (&key)[K], uint64_t sent) const
    {
        const bool     hit = has && (key[0] != sent) && (key_d2(key[0]) < thr2);
        const unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
        // integer count: any order gives the same sum
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&J.quality_count, (uint32_t)__popc(m));
    }
};

// Tile sweep (sweep_search.cuh): one CTA of WPI warps per item, items taken
// grid-stride.  gate 1: skip finished jobs and iterations at which the
// matcher does not run.
template <int K, int WPI, class Epi>
__global__ void __launch_bounds__(32 * WPI)
    search_sweep_kernel(const CloudView* __restrict__ clouds, JobDev* __restrict__ jobs, IcpDevParams P,
                        float cap_d2, int gate, Epi epi)
{
    JobDev& J = jobs[blockIdx.y];
    if (gate == 1 && (J.status != 0 || !matcher_active(P, J.iter))) return;  // matcher: running jobs only
    if (gate == 2 && (J.status == 0 || J.evaluated != 0)) return;            // quality: finished, once
    const CloudView cvL = clouds[J.to_cloud];
    const CloudView cvG = clouds[J.from_cloud];

    __shared__ float4  stage[WPI][kSweepCap];
    __shared__ GridDev sgrid;
    __shared__ double  sRt[12];
    __shared__ uint32_t s_items;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 12) sRt[tid] = (tid < 9) ? J.R[tid] : J.t[tid - 9];
    if (tid == 12) sgrid = *cvG.grid;
    if (tid == 13) s_items = cvL.grid->n_items;
    __syncthreads();
    const uint32_t n_items = (sgrid.n_valid > 0) ? s_items : 0u;
    const uint64_t sent = sentinel_key(cap_d2);

    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x)
    {
        const uint32_t first = __ldg(cvL.item_first + item);
        const uint32_t cnt = __ldg(cvL.item_first + item + 1) - first;
        const bool     has = (uint32_t)lane < cnt;
        float4         pl = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has) pl = __ldg(cvL.pts + first + lane);
        double gx, gy, gz;
        transform_point(sRt, pl, gx, gy, gz);
        const float qx = (float)gx, qy = (float)gy, qz = (float)gz;
        const bool  valid = has && (fabsf(qx) <= FLT_MAX) && (fabsf(qy) <= FLT_MAX) && (fabsf(qz) <= FLT_MAX);
        uint64_t    key[K];
#pragma unroll
        for (int i = 0; i < K; i++) key[i] = sent;
        sweep_search<K>(stage[warp], cvG, sgrid, valid, qx, qy, qz, cap_d2, warp, WPI, key);
        if (WPI > 1)
        {  // merge the warps' sorted lists into warp 0's
            uint64_t* mine = reinterpret_cast<uint64_t*>(stage[warp]);
            if (warp != 0)
            {
#pragma unroll
                for (int i = 0; i < K; i++) mine[i * 32 + lane] = key[i];
            }
            __syncthreads();
            if (warp == 0)
            {
                for (int w = 1; w < WPI; w++)
                {
                    const uint64_t* other = reinterpret_cast<const uint64_t*>(stage[w]);
#pragma unroll
                    for (int i = 0; i < K; i++)
                    {
                        const uint64_t kk = other[i * 32 + lane];
                        if (kk < key[K - 1]) topk_insert<K>(key, kk);
                    }
                }
            }
        }
        if (warp == 0) epi(J, cvG, has, first + lane, __float_as_uint(pl.w), key, sent);
        if (WPI > 1) __syncthreads();  // the staging buffers are reused by the next item
    }
}

// development probe (B200ICP_DBG_ITEMS=1, kNN call only): cycles spent per item, top bit = fallback search
__device__ uint32_t* g_dbg_item_cycles = nullptr;

// ---- the per-lane shell walk (tile_search.cuh) as the search stage ----------
// Neighbour rows of an earlier search of the same queries (the previous outer
// iteration's matcher run): every entry is a real point of the global cloud, so
// the k-th smallest of their distances under the CURRENT pose bounds the k-th
// nearest neighbour from above.  The search then runs with that bound as its
// radius cap -- all the seeds lie within it, hence so do the true k nearest --
// and gives exactly the result of the unseeded search (the seeds only bound,
// they are never inserted).  From the second outer iteration on the pose moves
// by centimetres: the bound is tight, the shell walk stops at the first shell
// and empty space is never walked.
struct SeedRows
{
    const uint32_t* rows;  // [job base + sorted position][k] sorted positions in the global cloud; null: none
    uint32_t        k;
};

template <int K>
__device__ __forceinline__ float seeded_cap(const CloudView& cvG, const uint32_t* __restrict__ row, uint32_t sk,
                                            float qx, float qy, float qz, float cap_d2)
{
    if (K == 1)
    {  // any seed bounds the nearest neighbour
        float best = cap_d2;
#pragma unroll
        for (int i = 0; i < B200ICP_MAX_KNN; i++)
            if ((uint32_t)i < sk)
            {
                const uint32_t p = row[i];
                if (p != kInvalid) best = fminf(best, dist2(qx, qy, qz, __ldg(cvG.pts + p)));
            }
        return best;
    }
    if (sk != (uint32_t)K) return cap_d2;
    float worst = 0.0f;
#pragma unroll
    for (int i = 0; i < K; i++)
    {
        const uint32_t p = row[i];
        if (p == kInvalid) return cap_d2;  // fewer than k within the radius last time: no bound
        worst = fmaxf(worst, dist2(qx, qy, qz, __ldg(cvG.pts + p)));
    }
    return fminf(cap_d2, worst);
}

// ---- the item sweep (item_sweep.cuh) as the search stage ---------------------
// Warps draw items from the job's counter (dynamic: the results do not depend on
// who searches what, and items differ in cost); the next item is drawn while the
// current one is searched.  Items are plain runs of kItem sorted points of the
// local cloud (cloud.cu).  A group of lanes too far apart for one box (a run
// that straddles a jump of the Morton curve) is split in two, down to single
// lanes, which search alone.
template <int K, class F>
__device__ __forceinline__ void for_each_item_sweep(SweepSmem& SW, const CloudView& cvL, const CloudView& cvG,
                                                    const GridDev& grid, const double* Rt, uint32_t n_items,
                                                    uint32_t n_local, uint32_t* next_item, float cap_d2,
                                                    bool bulk, const SeedRows& seed, F&& f)
{
    const int      lane = threadIdx.x & 31;
    const unsigned full = 0xFFFFFFFFu;
    uint32_t       par = 0;  // phase parities of the warp's two stage mbarriers (bulk-copy variant)
    uint32_t       item = 0;
    if (lane == 0) item = atomicAdd(next_item, 1u);
    item = __shfl_sync(full, item, 0);
#pragma unroll 1
    while (item < n_items)
    {
        uint32_t nxt = 0;
        if (lane == 0) nxt = atomicAdd(next_item, 1u);  // consumed at the end of this item
        B2_PHASE_DECL;
        const long long dbg_t0 = g_dbg_item_cycles ? clock64() : 0;
        const uint32_t first = item * (uint32_t)kItem;
        const uint32_t cnt = min((uint32_t)kItem, n_local - first);
        const bool     has = (uint32_t)lane < cnt;
        float4         pl = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has) pl = __ldg(cvL.pts + first + lane);
        double gx, gy, gz;
        transform_point(Rt, pl, gx, gy, gz);
        const float qx = (float)gx, qy = (float)gy, qz = (float)gz;
        // a query with a NaN / infinite coordinate has no neighbours
        const bool  hasq = has && (fabsf(qx) <= FLT_MAX) && (fabsf(qy) <= FLT_MAX) && (fabsf(qz) <= FLT_MAX);
        // this lane's radius cap: the caller's, or the tighter bound of its seeds
        float cap = cap_d2;
        if (seed.rows && hasq)
            cap = seeded_cap<K>(cvG, seed.rows + (size_t)(first + lane) * seed.k, seed.k, qx, qy, qz, cap_d2);
        uint64_t sent = sentinel_key(cap);
        uint64_t key[K];
#pragma unroll
        for (int i = 0; i < K; i++) key[i] = sent;
        B2_PHASE(0);
        uint32_t stack[6];
        int      sp = 0;
        stack[sp++] = __ballot_sync(full, hasq);
#pragma unroll 1
        while (sp > 0)
        {
            const uint32_t mask = stack[--sp];
            if (mask == 0u) continue;
            const bool in = (mask >> lane) & 1u;
            B2_COUNT(12, 1);
            if (item_sweep<K>(cvG, grid, SW, in, seed.rows != nullptr, bulk, par, qx, qy, qz, cap, key, sent) == 0) continue;
            const int n = __popc(mask);
            if (n == 1)
            {
                if (in) knn_search<K>(cvG, grid, qx, qy, qz, cap, key);
                continue;
            }
            // lanes too far apart for one box: lower half first
            const uint32_t cut = __fns(mask, 0u, n / 2 + 1);
            const uint32_t lower = mask & ((1u << cut) - 1u);
            stack[sp++] = mask & ~lower;
            stack[sp++] = lower;
        }
#ifdef B200ICP_DBG_PHASES
        b2_ph_t = clock64();
#endif
        f(has, first + lane, __float_as_uint(pl.w), key, sent);
        B2_PHASE(6);
        if (g_dbg_item_cycles && lane == 0)
            g_dbg_item_cycles[item] = (uint32_t)min((long long)0x7FFFFFFF, clock64() - dbg_t0);
        item = __shfl_sync(full, nxt, 0);
    }
}

struct ItemSmem
{
    SweepSmem sweep[kChunk / 32];
    GridDev   grid;
    double    Rt[12];
    uint32_t  n_items, n_local;
};

// The search stage.  seed_nn / seed_k: the launch's neighbour-row buffer of an
// earlier matcher search (rows of job j start at J.pair_base); used when the
// job says its rows are valid (JobDev::rows_valid, set by the solver after a
// matcher run).  Dynamic shared memory: sizeof(ItemSmem) (launch_search_k opts in).
#ifndef B200ICP_ITEM_MINB
#define B200ICP_ITEM_MINB 4
#endif
template <int K, class Epi>
__global__ void __launch_bounds__(kChunk, B200ICP_ITEM_MINB)
    search_item_kernel(const CloudView* __restrict__ clouds, JobDev* __restrict__ jobs, IcpDevParams P,
                       float cap_d2, int gate, const uint32_t* seed_nn, uint32_t seed_k, int bulk, Epi epi)
{
    JobDev& J = jobs[blockIdx.y];
    if (gate == 1 && (J.status != 0 || !matcher_active(P, J.iter))) return;  // matcher: running jobs only
    if (gate == 2 && (J.status == 0 || J.evaluated != 0)) return;            // quality: finished, once
    const CloudView cvL = clouds[J.to_cloud];
    const CloudView cvG = clouds[J.from_cloud];
    extern __shared__ __align__(16) unsigned char item_smem_raw[];
    ItemSmem& sm = *reinterpret_cast<ItemSmem*>(item_smem_raw);
    const int tid = threadIdx.x;
    if (bulk && (tid & 31) == 0)
    {
        mbar_init(&sm.sweep[tid >> 5].mbar[0], 1u);
        mbar_init(&sm.sweep[tid >> 5].mbar[1], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 12) sm.Rt[tid] = (tid < 9) ? J.R[tid] : J.t[tid - 9];
    if (tid == 12) sm.grid = *cvG.grid;
    if (tid == 13) sm.n_items = cvL.grid->n_items, sm.n_local = cvL.grid->n_valid;
    __syncthreads();
    const uint32_t n_items = (sm.grid.n_valid > 0) ? sm.n_items : 0u;
    SeedRows       seed = {nullptr, seed_k};
    if (seed_nn && J.rows_valid) seed.rows = seed_nn + (size_t)J.pair_base * seed_k;
    for_each_item_sweep<K>(sm.sweep[tid >> 5], cvL, cvG, sm.grid, sm.Rt, n_items, sm.n_local, &J.next_item, cap_d2,
                           bulk != 0, seed, [&](bool has, uint32_t pos, uint32_t orig, uint64_t (&key)[K], uint64_t sent) {
                               epi(J, cvG, has, pos, orig, key, sent);
                           });
}

// ---- the per-lane shell walk (tile_search.cuh) as the search stage: B200ICP_SEARCH=walk ------
// Warps draw items from the job's counter (dynamic: the results do not depend
// on who searches what, and items differ a lot in cost).  Per item (<= 32
// queries, one per lane): box of the queries' home cells -> tile -> search.
// Quality pass (k = 1, radius = the evaluator's threshold) right after a matcher search: most queries need no
// search at all.  A lane's seeds are the <= k nearest points the matcher found at the pose `cert_Rt`; here the
// question is only whether ANY point lies within the radius at the current pose.
//   * a seed within the radius  -> yes, and that seed is a real point: the search could only find it or a closer one;
//   * no seed within the radius -> every other point was at least `reach` away from the query at cert_Rt (the
//     farthest seed if the list is full; the matcher's own radius if it is not: then the list holds ALL points inside
//     it), and the query has moved by `mv` since: if reach - mv still exceeds the radius (2 mm of slack for the
//     float32 arithmetic of both distance evaluations), the answer is no.
// Lanes that are settled either way skip the search; the others search as usual.  An item whose lanes are all
// settled costs two loads per lane instead of a tile build and a walk.
// Returns 0: not settled, 1: a point within the radius (hit_key), 2: none.
template <int K>
__device__ __forceinline__ int settle_quality(const CloudView& cvG, const uint32_t* __restrict__ row, uint32_t sk,
                                               const double* cert_Rt, const float4& pl, float qx, float qy, float qz,
                                               float radius_d2, float rows_cap_d2, uint64_t& hit_key)
{
    double rx_, ry_, rz_;
    transform_point(cert_Rt, pl, rx_, ry_, rz_);
    const float rx = (float)rx_, ry = (float)ry_, rz = (float)rz_;
    if (!((fabsf(rx) <= FLT_MAX) && (fabsf(ry) <= FLT_MAX) && (fabsf(rz) <= FLT_MAX))) return 0;
    float    best = INFINITY, far = 0.0f;
    uint32_t best_idx = 0;
    bool     full = true;
#pragma unroll
    for (int i = 0; i < B200ICP_MAX_KNN; i++)
        if ((uint32_t)i < sk)
        {
            const uint32_t p = row[i];
            if (p == kInvalid)
            {
                full = false;
                continue;
            }
            const float4 c = __ldg(cvG.pts + p);
            const float  d = dist2(qx, qy, qz, c), dr = dist2(rx, ry, rz, c);
            if (d < best) best = d, best_idx = __float_as_uint(c.w);
            far = fmaxf(far, dr);
        }
    if (best < radius_d2)
    {
        hit_key = make_key(best, best_idx);
        return 1;
    }
    const float mv = sqrtf(dist2(qx, qy, qz, make_float4(rx, ry, rz, 0.f)));
    const float reach = full ? sqrtf(far) : sqrtf(rows_cap_d2);
    return (reach - mv - 2e-3f > sqrtf(radius_d2) * 1.001f) ? 2 : 0;
}

template <int K, class F>
__device__ __forceinline__ void for_each_item(WarpTile& W, const CloudView& cvL, const CloudView& cvG,
                                              const GridDev& grid, const double* Rt, uint32_t n_items,
                                              uint32_t* next_item, float cap_d2, const SeedRows& seed,
                                              const double* cert_Rt, float rows_cap_d2, F&& f)
{
    const int lane = threadIdx.x & 31;
    for (;;)
    {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(next_item, 1u);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        if (item >= n_items) break;
        const long long dbg_t0 = g_dbg_item_cycles ? clock64() : 0;
        const uint32_t first = __ldg(cvL.item_first + item);
        const uint32_t cnt = __ldg(cvL.item_first + item + 1) - first;
        const bool     has = (uint32_t)lane < cnt;
        float4         pl = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has) pl = __ldg(cvL.pts + first + lane);
        double gx, gy, gz;
        transform_point(Rt, pl, gx, gy, gz);
        const float qx = (float)gx, qy = (float)gy, qz = (float)gz;
        const bool  finite = has && (fabsf(qx) <= FLT_MAX) && (fabsf(qy) <= FLT_MAX) && (fabsf(qz) <= FLT_MAX);
        // quality pass: settled without a search?
        int      settle = 0;
        uint64_t hit_key = 0;
        if (K == 1 && cert_Rt && seed.rows && finite)
            settle = settle_quality<K>(cvG, seed.rows + (size_t)(first + lane) * seed.k, seed.k, cert_Rt, pl, qx, qy, qz,
                                        cap_d2, rows_cap_d2, hit_key);
        const bool settled = settle != 0;
        // this lane's radius cap: the caller's, or the tighter bound of its seeds
        float cap = cap_d2;
        if (seed.rows && finite && !settled)
            cap = seeded_cap<K>(cvG, seed.rows + (size_t)(first + lane) * seed.k, seed.k, qx, qy, qz, cap_d2);
        // shells the widest lane of the item may need (warp-uniform)
        float capmax = (finite && !settled) ? cap : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) capmax = fmaxf(capmax, __shfl_xor_sync(0xFFFFFFFFu, capmax, o));
        const int       S = search_shells(grid, capmax);
        const QueryCell qc = locate_query(grid, S, qx, qy, qz);
        const bool      hasq = has && qc.valid && !settled;
        int lo[3] = {hasq ? qc.hx : INT_MAX, hasq ? qc.hy : INT_MAX, hasq ? qc.hz : INT_MAX};
        int hi[3] = {hasq ? qc.hx : INT_MIN, hasq ? qc.hy : INT_MIN, hasq ? qc.hz : INT_MIN};
#pragma unroll
        for (int d = 0; d < 3; d++)
        {
            lo[d] = __reduce_min_sync(0xFFFFFFFFu, lo[d]);
            hi[d] = __reduce_max_sync(0xFFFFFFFFu, hi[d]);
        }
        TileGeom       G;
        const bool     tiled = warp_tile_build(W, G, cvG, lo, hi, S);
        const uint64_t sent = sentinel_key(cap);
        uint64_t       key[K];
#pragma unroll
        for (int i = 0; i < K; i++) key[i] = sent;
        if (settle == 1) key[0] = hit_key;  // a real point within the radius
        if (hasq)
        {
            if (tiled)
                tile_knn<K>(W, G, cvG, grid, qc, S, qx, qy, qz, key);
            else
                knn_search<K>(cvG, grid, qx, qy, qz, cap, key);
        }
        f(has, first + lane, __float_as_uint(pl.w), key, sent);
        if (g_dbg_item_cycles && lane == 0)
            g_dbg_item_cycles[item] = (uint32_t)min((long long)0x7FFFFFFF, clock64() - dbg_t0) | (tiled ? 0u : 0x80000000u);
#ifdef B200ICP_DBG_COUNT
        if (g_dbg_lane_cand && g_dbg_item_cycles)
        {   // per item: max and sum over lanes of the candidates scanned
            uint32_t* slot = g_dbg_lane_cand + (blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
            const uint32_t c = *slot;
            *slot = 0;
            const uint32_t mx = __reduce_max_sync(0xFFFFFFFFu, c), sm = __reduce_add_sync(0xFFFFFFFFu, c);
            if (lane == 0) g_dbg_item_cycles[n_items + 1 + 2 * item] = mx, g_dbg_item_cycles[n_items + 2 + 2 * item] = sm;
        }
#endif
    }
}

struct SearchSmem
{
    WarpTile tile[kChunk / 32];
    GridDev  grid;
    double   Rt[12];
    double   Rt_rows[12];  // pose of the matcher search whose rows seed this one (quality pass)
    uint32_t n_items;
};

// seed_nn / seed_k: the launch's neighbour-row buffer of an earlier matcher
// search (rows of job j start at J.pair_base); used when the job says its rows
// are valid (JobDev::rows_valid, set by the solver after a matcher run).
#ifndef B200ICP_SEARCH_MINB
#define B200ICP_SEARCH_MINB 4
#endif
template <int K, class Epi>
__global__ void __launch_bounds__(kChunk, B200ICP_SEARCH_MINB)
    search_tile_kernel(const CloudView* __restrict__ clouds, JobDev* __restrict__ jobs, IcpDevParams P,
                       float cap_d2, int gate, const uint32_t* seed_nn, uint32_t seed_k, Epi epi)
{
    JobDev& J = jobs[blockIdx.y];
    if (gate == 1 && (J.status != 0 || !matcher_active(P, J.iter))) return;  // matcher: running jobs only
    if (gate == 2 && (J.status == 0 || J.evaluated != 0)) return;            // quality: finished, once
    const CloudView cvL = clouds[J.to_cloud];
    const CloudView cvG = clouds[J.from_cloud];
    __shared__ SearchSmem sm;
    const int tid = threadIdx.x;
    if (tid < 12) sm.Rt[tid] = (tid < 9) ? J.R[tid] : J.t[tid - 9];
    if (tid == 12) sm.grid = *cvG.grid;
    if (tid == 13) sm.n_items = cvL.grid->n_items;
    if (tid >= 32 && tid < 44) sm.Rt_rows[tid - 32] = J.rows_Rt[tid - 32];
    __syncthreads();
    const uint32_t n_items = (sm.grid.n_valid > 0) ? sm.n_items : 0u;
    SeedRows       seed = {nullptr, seed_k};
    if (seed_nn && J.rows_valid) seed.rows = seed_nn + (size_t)J.pair_base * seed_k;
    // the quality pass after a matcher run of this registration: settle_quality
    const double* cert = (gate == 2 && K == 1 && seed.rows && J.rows_pose_valid) ? sm.Rt_rows : nullptr;
    for_each_item<K>(sm.tile[tid >> 5], cvL, cvG, sm.grid, sm.Rt, n_items, &J.next_item, cap_d2, seed, cert, P.thr2,
                     [&](bool has, uint32_t pos, uint32_t orig, uint64_t (&key)[K], uint64_t sent) {
                         epi(J, cvG, has, pos, orig, key, sent);
                     });
}

// ================================================================= fit stage
// Adds e e^T of the warp's (<= 32) pairings to the DMMA accumulators: the 16
// doubles of every paired lane go through the warp's staging buffer, half a
// warp at a time; groups of four lanes without a pairing are skipped.
__device__ __forceinline__ void accumulate_moments(double* st, bool paired, const double (&e)[16],
                                                   double (&c00)[2], double (&c01)[2], double (&c11)[2])
{
    const int      lane = threadIdx.x & 31;
    const unsigned any = __ballot_sync(0xFFFFFFFFu, paired);
    if (!any) return;
#pragma unroll
    for (int half = 0; half < 2; half++)
    {
        if (((any >> (16 * half)) & 0xFFFFu) == 0) continue;
        __syncwarp();
        if ((lane >> 4) == half)
        {
            double* d = st + (lane & 15) * kStageStride;
#pragma unroll
            for (int i = 0; i < 16; i++) d[i] = paired ? e[i] : 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int ks = 0; ks < 4; ks++)
        {
            if (((any >> (16 * half + 4 * ks)) & 0xFu) == 0) continue;
            const double* row = st + (4 * ks + (lane & 3)) * kStageStride + (lane >> 2);
            const double  f0 = row[0], f8 = row[8];
            dmma_8x8x4(c00[0], c00[1], f0, f0);
            dmma_8x8x4(c01[0], c01[1], f0, f8);
            dmma_8x8x4(c11[0], c11[1], f8, f8);
        }
    }
}

// ---- moments: chunk partials and their fixed reduction tree ------------------
// The local cloud's items are taken in CHUNKS of `chunk_items` consecutive items.
// One warp accumulates a chunk (DMMA accumulators in registers) and stores its
// 192 doubles as partial[chunk]; the partials of kFitGroup consecutive chunks
// are summed, in chunk order, into gpartial[group]; the group partials are summed
// in group order into the job's moment matrix.  The tree is defined over the
// DATA (positions in the local cloud), not over who executes what: any launch
// shape, any scheduling -- and any split of the chunks over several GPUs --
// gives the same bits.  Who runs a reduction step is decided by arrival
// counters: the last warp to deliver a partial of a group reduces the group, the
// last group reducer reduces the job and -- tail_mode 1 -- runs the solver on the
// result right away (no separate launch, no second pass over the partials).
constexpr int kFitGroup = 32;  // chunk partials per first-level sum

struct FitBuffers
{
    double*   partials;   // [job chunk base + chunk][192], fragment layout (see frag_slots)
    double*   gpartials;  // [job group base + group][192]
    uint32_t* tickets;    // [job group base + group] arrival counters; zero between launches
};

// Scratch of the solver (one warp), one per CTA: only one warp of a job's grid row ever gets there.
struct SolveSmem
{
    double S[kNumMoments];
    double A[144];
    double G0[12];   // sum a r0
    double T0[12];   // theta0 layout: [R_i0 R_i1 R_i2 t_i] x 3
    double R[9], t[3];
    double X[12], G12[12], Jm[72], B[72], H[36], g[6];
    int    stop;
};

// the six elements of S a lane holds after the DMMAs: rows r = lane / 4, columns 2 (lane % 4) and + 1 of each tile
__device__ __forceinline__ void frag_slots(int lane, int (&idx)[6])
{
    const int r = lane >> 2, c = 2 * (lane & 3);
    idx[0] = r * 8 + c, idx[1] = idx[0] + 1;
    idx[2] = 64 + r * 8 + c, idx[3] = idx[2] + 1;
    idx[4] = 128 + r * 8 + c, idx[5] = idx[4] + 1;
}

// Sum of `count` consecutive 192-double records, this lane's six slots, in a
// FIXED order: four interleaved running sums over the records in sequence, then
// a balanced tree.  L2 loads (the records were written by other SMs in this launch).
__device__ __forceinline__ void sum_records(const double* base, uint32_t count, const int (&idx)[6], double (&v)[6])
{
    double   a[4][6];
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int i = 0; i < 6; i++) a[u][i] = 0.0;
    uint32_t c = 0;
    for (; c + 4 <= count; c += 4)
    {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 6; i++) a[u][i] += __ldcg(base + (size_t)(c + u) * kNumMoments + idx[i]);
    }
    for (uint32_t u = 0; c < count; c++, u++)
#pragma unroll
        for (int i = 0; i < 6; i++)
        {   // the up-to-three records left go to the running sums 0, 1, 2
            const double x = __ldcg(base + (size_t)c * kNumMoments + idx[i]);
            if (u == 0) a[0][i] += x;
            else if (u == 1) a[1][i] += x;
            else a[2][i] += x;
        }
#pragma unroll
    for (int i = 0; i < 6; i++) v[i] = (a[0][i] + a[1][i]) + (a[2][i] + a[3][i]);
}

// Called by the whole warp that has just accumulated chunk `chunk` (of n_chunks)
// of job J.  Returns true on the ONE warp of the job that then holds the fully
// reduced moments: in sS[192] (shared) and in J.M.
__device__ __forceinline__ bool deliver_chunk(JobDev& J, uint32_t chunk, uint32_t n_chunks, const double (&c00)[2],
                                              const double (&c01)[2], const double (&c11)[2],
                                              const FitBuffers& fb, double* sS, bool groups_only)
{
    const int lane = threadIdx.x & 31;
    int       idx[6];
    frag_slots(lane, idx);
    double* pp = fb.partials + (size_t)(J.chunk_base + chunk) * kNumMoments;
    pp[idx[0]] = c00[0], pp[idx[1]] = c00[1], pp[idx[2]] = c01[0];
    pp[idx[3]] = c01[1], pp[idx[4]] = c11[0], pp[idx[5]] = c11[1];
    __threadfence();
    __syncwarp();
    const uint32_t g = chunk / kFitGroup, n_groups = (n_chunks + kFitGroup - 1) / kFitGroup;
    const uint32_t in_group = min((uint32_t)kFitGroup, n_chunks - g * kFitGroup);
    uint32_t       t = 0;
    if (lane == 0) t = atomicAdd(fb.tickets + J.group_base + g, 1u);
    t = __shfl_sync(0xFFFFFFFFu, t, 0);
    if (t != in_group - 1) return false;
    // last of its group: first-level sum
    __threadfence();
    double v[6];
    sum_records(fb.partials + (size_t)(J.chunk_base + g * kFitGroup) * kNumMoments, in_group, idx, v);
    double* gp = fb.gpartials + (size_t)(J.group_base + g) * kNumMoments;
#pragma unroll
    for (int i = 0; i < 6; i++) gp[idx[i]] = v[i];
    if (lane == 0) fb.tickets[J.group_base + g] = 0;  // ready for the next launch
    if (groups_only) return false;  // sharded registration: the group partials of all ranks are summed after the exchange
    __threadfence();
    __syncwarp();
    if (lane == 0) t = atomicAdd(&J.groups_done, 1u);
    t = __shfl_sync(0xFFFFFFFFu, t, 0);
    if (t != n_groups - 1) return false;
    // last group: the job's moments
    if (lane == 0) J.groups_done = 0;
    __threadfence();
    const long long t0 = clock64();
    sum_records(fb.gpartials + (size_t)J.group_base * kNumMoments, n_groups, idx, v);
    if (lane == 0) J.dbg_tail[0] = (uint32_t)(clock64() - t0);
#pragma unroll
    for (int i = 0; i < 6; i++) J.Mprev[idx[i]] = J.M[idx[i]], sS[idx[i]] = v[i], J.M[idx[i]] = v[i];
    __syncwarp();
    return true;
}

struct FitSmem
{
    double    stage[kChunk / 32][16 * kStageStride];  // half a warp of e-vectors per round
    SolveSmem solve;
    double    Rt[12];
    uint32_t  n_items;
};

__device__ void solve_job_warp(JobDev& J, SolveSmem& ss, const IcpDevParams& P, uint32_t* n_active);

// what a fit kernel does once a warp holds the job's reduced moments (or there is nothing to reduce)
__device__ __forceinline__ void fit_job_tail(JobDev& J, FitSmem& sm, int tail_mode, const IcpDevParams& P,
                                             uint32_t* n_active)
{
    if (tail_mode == 1) solve_job_warp(J, sm.solve, P, n_active);
}

// a job whose matcher does not run at this iteration (or that has nothing to match): zero moments, then the tail
__device__ __forceinline__ void fit_job_without_chunks(JobDev& J, FitSmem& sm, int tail_mode, const IcpDevParams& P,
                                                       uint32_t* n_active)
{
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < kNumMoments; i += 32) sm.solve.S[i] = 0.0, J.M[i] = 0.0;
    __syncwarp();
    fit_job_tail(J, sm, tail_mode, P, n_active);
}

// Matcher_Point2Plane after the search (rows H, J) + the per-pairing part of
// optimal_tf_gauss_newton (rows K, L).  One thread per local point, in the
// cloud's sorted order; a warp takes whole chunks of `chunk_items` items (see
// "chunk partials" above: the summation order of the moments never depends on
// scheduling or on the launch shape).  The warp that completes the job's
// reduction runs the solver (tail_mode 1) or leaves the moments in J.M (0).
// grid = (CTAs per job, jobs).  nn = [job base + sorted position][K].
template <int K, bool WRITE>
__global__ void __launch_bounds__(kChunk, 4)
    fit_plane_kernel(const CloudView* __restrict__ clouds, JobDev* __restrict__ jobs,
                     const uint32_t* __restrict__ nn, FitBuffers fb, uint32_t chunk_items, int tail_mode,
                     IcpDevParams P, MatchOut out, PairRec* __restrict__ pairs, uint32_t* __restrict__ n_active,
                     uint32_t chunk_begin, uint32_t chunk_end)
{
    const uint32_t job = blockIdx.y;
    JobDev&        J = jobs[job];
    if (J.status != 0) return;
    const CloudView cvL = clouds[J.to_cloud];
    const CloudView cvG = clouds[J.from_cloud];
    PairRec*        prec = pairs ? pairs + J.pair_base : nullptr;
    const uint32_t* nnj = nn + (size_t)J.pair_base * K;

    __shared__ FitSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 12) sm.Rt[tid] = (tid < 9) ? J.R[tid] : J.t[tid - 9];
    if (tid == 32) sm.n_items = (cvG.grid->n_valid > 0) ? cvL.grid->n_items : 0u;
    __syncthreads();
    const uint32_t n_items = matcher_active(P, J.iter) ? sm.n_items : 0u;
    const uint32_t n_chunks = (n_items + chunk_items - 1) / chunk_items;
    double*        st = sm.stage[warp];
    if (n_chunks == 0)
    {
        if (blockIdx.x == 0 && warp == 0 && tail_mode != 2) fit_job_without_chunks(J, sm, tail_mode, P, n_active);
        return;
    }

    // the delivery that completes a job is always some warp's LAST chunk: the tail runs after the loop, when
    // nothing of the loop is live any more.  [chunk_begin, chunk_end): the chunks this launch is responsible for
    // (all of them, or this rank's slice of a sharded registration -- tail_mode 2: group partials only)
    bool completes_job = false;
    for (uint32_t chunk = chunk_begin + item_warp_id(); chunk < min(n_chunks, chunk_end); chunk += item_warp_count())
    {
      double c00[2] = {0, 0}, c01[2] = {0, 0}, c11[2] = {0, 0};
      for (uint32_t item = chunk * chunk_items; item < min(n_items, (chunk + 1) * chunk_items); item++)
      {
        const uint32_t first = __ldg(cvL.item_first + item);
        const uint32_t cnt = __ldg(cvL.item_first + item + 1) - first;
        const bool     has = (uint32_t)lane < cnt;
        const uint32_t pos = first + lane;
        float4         pl = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has) pl = __ldg(cvL.pts + pos);
        double gx, gy, gz;
        transform_point(sm.Rt, pl, gx, gy, gz);
        const float  qx = (float)gx, qy = (float)gy, qz = (float)gz;
        const double px = pl.x, py = pl.y, pz = pl.z;
        bool         paired = false;
        double       nrm[3] = {0, 0, 0}, r0 = 0;
        double       cen[3] = {0, 0, 0};
        if (has)
        {
            uint32_t nb[K];
            {
                const uint2* row = reinterpret_cast<const uint2*>(nnj + (size_t)pos * K);
#pragma unroll
                for (int i = 0; i < K / 2; i++)
                {
                    const uint2 v = __ldg(row + i);
                    nb[2 * i] = v.x, nb[2 * i + 1] = v.y;
                }
            }
            // neighbours kept after the distance cut; K may exceed the configured knn
            uint32_t m = 0;
#pragma unroll
            for (int i = 0; i < K; i++)
                if ((uint32_t)i < P.knn && nb[i] != kInvalid) m++;

            const uint32_t orig = __float_as_uint(pl.w);
            if (WRITE)
            {   // the rows hold sorted positions; the caller gets original indices
                if (out.nn_cnt) out.nn_cnt[orig] = m;
                if (out.nn_idx)
#pragma unroll
                    for (int i = 0; i < K; i++)
                        if ((uint32_t)i < P.knn)
                            out.nn_idx[(size_t)orig * P.knn + i] =
                                ((uint32_t)i < m) ? __float_as_uint(__ldg(cvG.pts + nb[i]).w) : kInvalid;
            }

            if (m >= P.min_plane_points && m > 0)
            {
                // row J: mean and covariance (1/m) of the neighbours in f64,
                // accumulated in neighbour order
                double nx_[K], ny_[K], nz_[K];
                double sx = 0, sy = 0, sz = 0;
#pragma unroll
                for (int i = 0; i < K; i++)
                    if ((uint32_t)i < m)
                    {
                        const float4 pn = __ldg(cvG.pts + nb[i]);
                        nx_[i] = (double)pn.x, ny_[i] = (double)pn.y, nz_[i] = (double)pn.z;
                        sx += nx_[i], sy += ny_[i], sz += nz_[i];
                    }
                const double inv = 1.0 / (double)m;
                const double cx = sx * inv, cy = sy * inv, cz = sz * inv;
                double c00_ = 0, c01_ = 0, c02_ = 0, c11_ = 0, c12_ = 0, c22_ = 0;
#pragma unroll
                for (int i = 0; i < K; i++)
                    if ((uint32_t)i < m)
                    {
                        const double dx = nx_[i] - cx, dy = ny_[i] - cy, dz = nz_[i] - cz;
                        c00_ += dx * dx, c01_ += dx * dy, c02_ += dx * dz;
                        c11_ += dy * dy, c12_ += dy * dz, c22_ += dz * dz;
                    }
                double C[9] = {c00_ * inv, c01_ * inv, c02_ * inv, c01_ * inv, c11_ * inv,
                               c12_ * inv, c02_ * inv, c12_ * inv, c22_ * inv};
                double ev[3], V[9];
                jacobi3(C, ev, V);
                if (!(ev[0] > P.plane_eigen_threshold * ev[2]))
                {
                    double nx = V[0], ny = V[3], nz = V[6];
                    const double lead = (nx != 0.0) ? nx : ((ny != 0.0) ? ny : nz);
                    if (lead < 0) nx = -nx, ny = -ny, nz = -nz;
                    const double dist = fabs((nx * ((double)qx - cx) + ny * ((double)qy - cy)) +
                                             nz * ((double)qz - cz));
                    if (!(dist > P.distance_threshold))
                    {
                        paired = true;
                        nrm[0] = nx, nrm[1] = ny, nrm[2] = nz;
                        cen[0] = cx, cen[1] = cy, cen[2] = cz;
                        // residual at T0 with the f64 transformed point (row L)
                        r0 = (nx * (gx - cx) + ny * (gy - cy)) + nz * (gz - cz);
                        if (WRITE)
                        {
                            if (out.centroid)
                                out.centroid[(size_t)orig * 3] = cx, out.centroid[(size_t)orig * 3 + 1] = cy,
                                                          out.centroid[(size_t)orig * 3 + 2] = cz;
                            if (out.normal)
                                out.normal[(size_t)orig * 3] = nx, out.normal[(size_t)orig * 3 + 1] = ny,
                                                        out.normal[(size_t)orig * 3 + 2] = nz;
                        }
                    }
                }
            }
            if (WRITE && out.paired) out.paired[orig] = paired ? 1 : 0;
            if (prec)
            {  // what the closed-form solver consumes: the plane centroid (A.10)
                PairRec r;
                r.q[0] = cen[0], r.q[1] = cen[1], r.q[2] = cen[2];
                r.paired = paired ? 1u : 0u, r.pad = 0u;
                prec[pos] = r;
            }
        }

        // ---- moments on the FP64 tensor-core path ---------------------------
        double e[16];
        const double f1 = paired ? 1.0 : 0.0;
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            const double ni = nrm[i];  // zero when unpaired
            e[4 * i + 0] = ni * px, e[4 * i + 1] = ni * py, e[4 * i + 2] = ni * pz, e[4 * i + 3] = ni;
        }
        e[12] = r0, e[13] = f1, e[14] = 0.0, e[15] = 0.0;
        accumulate_moments(st, paired, e, c00, c01, c11);
      }
      completes_job = deliver_chunk(J, chunk, n_chunks, c00, c01, c11, fb, sm.solve.S, tail_mode == 2);
    }
    if (completes_job) fit_job_tail(J, sm, tail_mode, P, n_active);
}

// Matcher_Points_DistanceThreshold as the ICP matcher: 1-NN with d2 < thr^2
// (strict, A.8 / orc_match_points).  Moments of the point-to-point residual
// r = R p + t - q, linear in theta like the point-to-plane one:
//   e = [ h (4) | r0 (3) | 0 ... ],  h = (p_local, 1)
// so tile C00 holds sum h h^T, sum h r0^T, sum |r0|^2 (trace of the r0 block)
// and the pairing count (S[3][3]).  nn = [job base + sorted position][1].
template <bool WRITE>
__global__ void __launch_bounds__(kChunk)
    fit_p2p_kernel(const CloudView* __restrict__ clouds, JobDev* __restrict__ jobs,
                   const uint32_t* __restrict__ nn, FitBuffers fb, uint32_t chunk_items, int tail_mode,
                   IcpDevParams P, MatchOut out, PairRec* __restrict__ pairs, uint32_t* __restrict__ n_active,
                   uint32_t chunk_begin, uint32_t chunk_end)
{
    const uint32_t job = blockIdx.y;
    JobDev&        J = jobs[job];
    if (J.status != 0) return;
    const CloudView cvL = clouds[J.to_cloud];
    const CloudView cvG = clouds[J.from_cloud];
    PairRec*        prec = pairs ? pairs + J.pair_base : nullptr;
    const uint32_t* nnj = nn + (size_t)J.pair_base;

    __shared__ FitSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 12) sm.Rt[tid] = (tid < 9) ? J.R[tid] : J.t[tid - 9];
    if (tid == 32) sm.n_items = (cvG.grid->n_valid > 0) ? cvL.grid->n_items : 0u;
    __syncthreads();
    const uint32_t n_items = matcher_active(P, J.iter) ? sm.n_items : 0u;
    const uint32_t n_chunks = (n_items + chunk_items - 1) / chunk_items;
    double*        st = sm.stage[warp];
    if (n_chunks == 0)
    {
        if (blockIdx.x == 0 && warp == 0 && tail_mode != 2) fit_job_without_chunks(J, sm, tail_mode, P, n_active);
        return;
    }

    // the delivery that completes a job is always some warp's LAST chunk: the tail runs after the loop, when
    // nothing of the loop is live any more.  [chunk_begin, chunk_end): the chunks this launch is responsible for
    // (all of them, or this rank's slice of a sharded registration -- tail_mode 2: group partials only)
    bool completes_job = false;
    for (uint32_t chunk = chunk_begin + item_warp_id(); chunk < min(n_chunks, chunk_end); chunk += item_warp_count())
    {
      double c00[2] = {0, 0}, c01[2] = {0, 0}, c11[2] = {0, 0};
      for (uint32_t item = chunk * chunk_items; item < min(n_items, (chunk + 1) * chunk_items); item++)
      {
        const uint32_t first = __ldg(cvL.item_first + item);
        const uint32_t cnt = __ldg(cvL.item_first + item + 1) - first;
        const bool     has = (uint32_t)lane < cnt;
        const uint32_t pos = first + lane;
        float4         pl = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has) pl = __ldg(cvL.pts + pos);
        double gx, gy, gz;
        transform_point(sm.Rt, pl, gx, gy, gz);
        bool   paired = false;
        double q[3] = {0, 0, 0};
        if (has)
        {
            const uint32_t nb = __ldg(nnj + pos);
            const uint32_t orig = __float_as_uint(pl.w);
            uint32_t       nb_orig = kInvalid;
            if (nb != kInvalid)
            {
                const float4 pn = __ldg(cvG.pts + nb);  // the rows hold sorted positions
                nb_orig = __float_as_uint(pn.w);
                // the search keeps d2 <= thr^2; the matcher pairs d2 < thr^2
                paired = dist2((float)gx, (float)gy, (float)gz, pn) < P.thr2;
                if (paired) q[0] = (double)pn.x, q[1] = (double)pn.y, q[2] = (double)pn.z;
            }
            if (WRITE)
            {
                if (out.nn_cnt) out.nn_cnt[orig] = paired ? 1u : 0u;
                if (out.nn_idx) out.nn_idx[orig] = paired ? nb_orig : kInvalid;
                if (out.paired) out.paired[orig] = paired ? 1 : 0;
                if (paired && out.centroid)
                    out.centroid[(size_t)orig * 3] = q[0], out.centroid[(size_t)orig * 3 + 1] = q[1],
                                              out.centroid[(size_t)orig * 3 + 2] = q[2];
            }
            if (prec)
            {
                PairRec r;
                r.q[0] = q[0], r.q[1] = q[1], r.q[2] = q[2];
                r.paired = paired ? 1u : 0u, r.pad = 0u;
                prec[pos] = r;
            }
        }
        double e[16];
        e[0] = pl.x, e[1] = pl.y, e[2] = pl.z, e[3] = 1.0;
        e[4] = gx - q[0], e[5] = gy - q[1], e[6] = gz - q[2];
#pragma unroll
        for (int i = 7; i < 16; i++) e[i] = 0.0;
        accumulate_moments(st, paired, e, c00, c01, c11);
      }
      completes_job = deliver_chunk(J, chunk, n_chunks, c00, c01, c11, fb, sm.solve.S, tail_mode == 2);
    }
    if (completes_job) fit_job_tail(J, sm, tail_mode, P, n_active);
}

// ------------------------------------------------------------------- solver
// End of one outer iteration of ICP::align (row G / A.7), one thread: adopt
// the solver's pose, test the step delta = log(Tprev^-1 * Tnew) split into
// (xyz, rot), update the job's status.
// One outer iteration maps the pose to the next pose, pose' = f(pose): the search is exact whatever bounds it,
// the moments are summed in an order fixed by the data, the solver is deterministic -- f is a pure function
// of the pose (as long as the matcher is active at every iteration).  ICP on coarse clouds often ends in a
// 2-cycle between two pairing sets, which the reference's loop runs up to maxIterations: 100 iterations that
// alternate between two poses.  Inside such a cycle the pairing sets repeat exactly and each solve lands on the
// minimiser of the same quadratic, so the poses repeat up to the rounding of the Gauss-Newton iterates
// (~1e-14); they repeat BITWISE only by luck.  The cycle is recognised when, on two consecutive iterations,
// the new pose equals the pose of two iterations before within 1e-11 (m / rad; `detect_cycles` 1: bitwise
// only) and the pairing count equals the one of two iterations before.  The rest of the run is then known: poses,
// moments and pairing counts alternate, no step is ever below the tolerances (these were not), the loop ends
// with MaxIterations in the state its parity gives.  The job jumps there: same iteration count, termination
// reason and pairing count as running it out, pose within 1e-11 of it (the stated tolerance is 1e-5 m / 1e-6
// rad; B200ICP_CYCLE=0 runs every iteration).
// Returns 0: nothing special; 1: fast-forwarded, state as is; 2: fast-forwarded, and the caller must swap the
// moments with the previous ones (done by the whole warp).
__device__ int finish_outer_iteration(JobDev& J, const Pose& Tn, uint32_t npair, uint32_t inner,
                                      const IcpDevParams& P, uint32_t* n_active, bool allow_cycle)
{
    Pose T0, dT;
    for (int i = 0; i < 9; i++) T0.R[i] = J.R[i];
    for (int i = 0; i < 3; i++) T0.t[i] = J.t[i];
    pose_inverse_compose(T0, Tn, dT);
    double d[6];
    se3_log(dT, d);
    const double dxyz = sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
    const double drot = sqrt((d[3] * d[3] + d[4] * d[4]) + d[5] * d[5]);
    bool same_as_two_back = allow_cycle && P.detect_cycles && J.iter >= 2 && P.run_from_iteration == 0 &&
                            P.run_up_to_iteration == 0 && npair == J.npair_prev;  // npair_prev: two iterations back
    if (same_as_two_back)
    {
        const double tol = (P.detect_cycles >= 2) ? 1e-11 : 0.0;
        for (int i = 0; i < 9; i++) same_as_two_back = same_as_two_back && (fabs(Tn.R[i] - J.Rprev[i]) <= tol);
        for (int i = 0; i < 3; i++) same_as_two_back = same_as_two_back && (fabs(Tn.t[i] - J.tprev[i]) <= tol);
    }
    J.cycle_hits = same_as_two_back ? J.cycle_hits + 1 : 0;
    same_as_two_back = J.cycle_hits >= 2;
    const uint32_t npair_before = J.n_pairings, inner_before = J.inner_prev;
    for (int i = 0; i < 9; i++) J.Rprev[i] = T0.R[i], J.R[i] = Tn.R[i];
    for (int i = 0; i < 3; i++) J.tprev[i] = T0.t[i], J.t[i] = Tn.t[i];
    J.npair_prev = npair_before;
    J.n_pairings = npair;
    J.inner_iters_total += inner;
    J.inner_prev = inner;
    if (dxyz < P.min_abs_step_trans && drot < P.min_abs_step_rot)
    {
        J.status = 1;
        J.term_reason = B200ICP_TERM_STALLED;
        atomicSub(n_active, 1u);
        return 0;
    }
    const uint32_t it = J.iter + 1;  // iterations completed
    J.iter = it;
    if (it >= P.max_iterations)
    {
        J.status = 1;
        J.term_reason = B200ICP_TERM_MAX_ITERATIONS;
        atomicSub(n_active, 1u);
        return 0;
    }
    if (!same_as_two_back)
    {
        // Longer cycles (period 3 .. kCycleMax; a 3-cycle between three pairing sets is as common on coarse clouds
        // as the 2-cycle).  Same recognition -- the new pose and pairing count equal those of `p` iterations before,
        // p times in a row -- but a simpler jump: whole periods are skipped (the state after them is this one) and
        // the remaining (left mod p) iterations simply run; the loop then ends with MaxIterations by itself.
        if (allow_cycle && P.detect_cycles && J.cycle_at == 0 && P.run_from_iteration == 0 && P.run_up_to_iteration == 0)
        {
            const double tol = (P.detect_cycles >= 2) ? 1e-11 : 0.0;
            uint32_t     found = 0;
            for (uint32_t p = 3; p <= (uint32_t)kCycleMax; p++)
            {
                bool same = it > p && J.hist_npair[(it - p) % kCycleMax] == npair;
                if (same)
                {
                    const double* h = J.hist_pose[(it - p) % kCycleMax];
                    for (int i = 0; i < 9; i++) same = same && (fabs(Tn.R[i] - h[i]) <= tol);
                    for (int i = 0; i < 3; i++) same = same && (fabs(Tn.t[i] - h[9 + i]) <= tol);
                }
                J.cyc_hits[p] = same ? J.cyc_hits[p] + 1 : 0;
                if (!found && J.cyc_hits[p] >= p) found = p;
            }
            double* h = J.hist_pose[it % kCycleMax];
            for (int i = 0; i < 9; i++) h[i] = Tn.R[i];
            for (int i = 0; i < 3; i++) h[9 + i] = Tn.t[i];
            J.hist_npair[it % kCycleMax] = npair;
            J.hist_inner[it % kCycleMax] = inner;
            if (found)
            {
                const uint32_t left = P.max_iterations - it, periods = left / found;
                uint32_t       inner_per_period = 0;
                for (uint32_t k = 0; k < found; k++) inner_per_period += J.hist_inner[(it - k) % kCycleMax];
                J.cycle_at = it;
                J.iter = it + periods * found;
                J.inner_iters_total += periods * inner_per_period;
                if (J.iter >= P.max_iterations)
                {   // a whole number of periods was left: this state is the final one
                    J.status = 1;
                    J.term_reason = B200ICP_TERM_MAX_ITERATIONS;
                    atomicSub(n_active, 1u);
                }
            }
        }
        return 0;
    }
    // period 2 from here on: `left` more iterations would run; an even number leaves this state, an odd number
    // the one of the previous iteration (pose_i, moments and pairings of iteration i - 1)
    const uint32_t left = P.max_iterations - it;
    J.cycle_at = it;
    J.iter = P.max_iterations;
    J.status = 1;
    J.term_reason = B200ICP_TERM_MAX_ITERATIONS;
    J.inner_iters_total += (left / 2) * (inner + inner_before) + ((left & 1u) ? inner_before : 0u);
    atomicSub(n_active, 1u);
    if ((left & 1u) == 0) return 1;
    for (int i = 0; i < 9; i++) J.R[i] = T0.R[i], J.Rprev[i] = Tn.R[i];
    for (int i = 0; i < 3; i++) J.t[i] = T0.t[i], J.tprev[i] = Tn.t[i];
    J.n_pairings = npair_before;
    J.npair_prev = npair;
    return 2;
}

// optimal_tf_gauss_newton on the reduced moments (A.6) + the end of the outer
// iteration (A.7), by ONE warp: the one that completed the job's reduction in the
// fit kernel.  ss.S holds the moments.
__device__ void solve_job_warp(JobDev& J, SolveSmem& ss, const IcpDevParams& P, uint32_t* n_active)
{
    const int lane = threadIdx.x & 31;
    if (lane == 0)
    {
        J.next_item = 0;                                   // the next search draws items from 0 again
        if (matcher_active(P, J.iter))
        {   // this iteration's search wrote the neighbour rows, at the pose the job still holds
            J.rows_valid = 1;
            J.rows_pose_valid = 1;
            for (int i = 0; i < 9; i++) J.rows_Rt[i] = J.R[i];
            for (int i = 0; i < 3; i++) J.rows_Rt[9 + i] = J.t[i];
        }
    }
    const bool     p2p = (P.matcher_kind == B200ICP_MATCHER_POINTS_DISTANCE);
    const uint32_t npair = (uint32_t)(pairing_count(ss.S, p2p) + 0.5);
    if (npair == 0)
    {
        if (lane == 0)
        {
            J.n_pairings = 0;
            J.status = 1;
            J.term_reason = B200ICP_TERM_NO_PAIRINGS;
            atomicSub(n_active, 1u);
        }
        return;
    }
    for (int e = lane; e < 144; e += 32) ss.A[e] = normal_matrix_at(ss.S, p2p, e / 12, e % 12);
    if (lane < 12)
    {
        const int    i = lane >> 2, j = lane & 3;
        const double v = (j < 3) ? J.R[i * 3 + j] : J.t[i];
        ss.T0[lane] = v;
        ss.G0[lane] = p2p ? moment_at(ss.S, j, 4 + i) : moment_at(ss.S, lane, 12);
        if (j < 3)
            ss.R[i * 3 + j] = v;
        else
            ss.t[i] = v;
    }
    __syncwarp();

    const long long t_gn = clock64();
    uint32_t inner = 0;
    for (uint32_t iter = 0; iter < P.solver_max_iterations; iter++)
    {
        // x = theta(T) - theta0
        if (lane < 12)
        {
            const int i = lane >> 2, j = lane & 3;
            ss.X[lane] = ((j < 3) ? ss.R[i * 3 + j] : ss.t[i]) - ss.T0[lane];
        }
        // J = d theta / d eps (12 x 6): columns v0..2 then w0..2
        for (int e = lane; e < 72; e += 32)
        {
            const int a = e / 6, c = e % 6, i = a >> 2, l = a & 3;
            double    v = 0.0;
            if (c < 3)
                v = (l == 3) ? ss.R[i * 3 + c] : 0.0;
            else if (l < 3)
            {
                const int j = c - 3;
                // (R [e_j]x)[i][l]
                if (j == 0)
                    v = (l == 1) ? ss.R[i * 3 + 2] : (l == 2 ? -ss.R[i * 3 + 1] : 0.0);
                else if (j == 1)
                    v = (l == 0) ? -ss.R[i * 3 + 2] : (l == 2 ? ss.R[i * 3 + 0] : 0.0);
                else
                    v = (l == 0) ? ss.R[i * 3 + 1] : (l == 1 ? -ss.R[i * 3 + 0] : 0.0);
            }
            ss.Jm[e] = v;
        }
        __syncwarp();
        // g12 = s + A x
        if (lane < 12)
        {
            double acc = ss.G0[lane];
#pragma unroll
            for (int b = 0; b < 12; b++) acc += ss.A[lane * 12 + b] * ss.X[b];
            ss.G12[lane] = acc;
        }
        // B = A J
        for (int e = lane; e < 72; e += 32)
        {
            const int a = e / 6, c = e % 6;
            double    acc = 0;
#pragma unroll
            for (int b = 0; b < 12; b++) acc += ss.A[a * 12 + b] * ss.Jm[b * 6 + c];
            ss.B[e] = acc;
        }
        __syncwarp();
        // H = J^T B, g = J^T g12
        for (int e = lane; e < 36; e += 32)
        {
            const int r = e / 6, c = e % 6;
            double    acc = 0;
#pragma unroll
            for (int a = 0; a < 12; a++) acc += ss.Jm[a * 6 + r] * ss.B[a * 6 + c];
            ss.H[e] = acc;
        }
        if (lane < 6)
        {
            double acc = 0;
#pragma unroll
            for (int a = 0; a < 12; a++) acc += ss.Jm[a * 6 + lane] * ss.G12[a];
            ss.g[lane] = acc;
        }
        __syncwarp();
        if (lane == 0)
        {
            double Hm[36], mg[6], delta[6];
            for (int i = 0; i < 36; i++) Hm[i] = 0.5 * (ss.H[i] + ss.H[(i % 6) * 6 + i / 6]);
            for (int i = 0; i < 6; i++) mg[i] = -ss.g[i];
            solve6_spd(Hm, mg, delta);
            Pose T, dT, Tn;
            for (int i = 0; i < 9; i++) T.R[i] = ss.R[i];
            for (int i = 0; i < 3; i++) T.t[i] = ss.t[i];
            se3_exp(delta, dT);
            pose_compose(T, dT, Tn);
            for (int i = 0; i < 9; i++) ss.R[i] = Tn.R[i];
            for (int i = 0; i < 3; i++) ss.t[i] = Tn.t[i];
            double nd = 0;
            for (int i = 0; i < 6; i++) nd += delta[i] * delta[i];
            ss.stop = (sqrt(nd) < P.gn_min_delta) ? 1 : 0;
        }
        __syncwarp();
        inner++;
        // every lane takes its copy before lane 0 may reuse ss.stop below (racecheck: read here / write after the loop)
        const int stop_now = ss.stop;
        __syncwarp();
        if (stop_now) break;
    }
    if (lane == 0)
    {
        const long long t_fin = clock64();
        Pose Tn;
        for (int i = 0; i < 9; i++) Tn.R[i] = ss.R[i];
        for (int i = 0; i < 3; i++) Tn.t[i] = ss.t[i];
        ss.stop = finish_outer_iteration(J, Tn, npair, inner, P, n_active, true);
        J.dbg_tail[1] = (uint32_t)(t_fin - t_gn), J.dbg_tail[2] = (uint32_t)(clock64() - t_fin), J.dbg_tail[3] = inner;
    }
    __syncwarp();
    if (ss.stop == 2)
        for (int i = lane; i < kNumMoments; i += 32)
        {   // the run ends one iteration "earlier" in the cycle: its moments are the previous ones
            const double m = J.M[i];
            J.M[i] = J.Mprev[i], J.Mprev[i] = m;
        }
}

// ------------------------------------------------------------ Horn (row N)
// optimal_tf_horn on the pairings the matcher left in the pair buffer, with the
// pairings-weight rules of row M.  Two passes over the pairings (A.10):
//   phase 0: centroids of ALL pairings (sum p, sum q, count)
//   phase 1: per pairing b = p - pc, a = q - qc; scale-outlier rule; optional
//            robust weight; S += w b a^T; pairs used
// then N(S) (4x4), its top eigenvector = quaternion, t = qc - R pc.
constexpr int kHornVals = 16;   // doubles per CTA partial of either phase
constexpr int kHornThreads = 256;

// centroids from the phase-0 partials, in one fixed order (used by every CTA
// of phase 1 and by the final solve: all must see the same bits)
__device__ __forceinline__ void horn_centroids(const double* __restrict__ part0, uint32_t n_part,
                                               double* pc, double* qc, double& count)
{
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (uint32_t c = 0; c < n_part; c++)
#pragma unroll
        for (int i = 0; i < 7; i++) acc[i] += part0[(size_t)c * kHornVals + i];
    count = acc[6];
    const double n = acc[6] > 0 ? acc[6] : 1.0;
    for (int d = 0; d < 3; d++) pc[d] = acc[d] / n, qc[d] = acc[3 + d] / n;
}

__global__ void __launch_bounds__(kHornThreads)
    horn_sum_kernel(const CloudView* __restrict__ clouds, const JobDev* __restrict__ jobs,
                    const PairRec* __restrict__ pairs, const double* __restrict__ part0,
                    double* __restrict__ part_out, IcpDevParams P, int phase)
{
    const uint32_t job = blockIdx.y;
    const JobDev&  J = jobs[job];
    if (J.status != 0) return;
    const CloudView cvL = clouds[J.to_cloud];
    const PairRec*  prec = pairs + J.pair_base;
    const uint32_t  n = cvL.grid->n_valid;
    __shared__ double sc[8];
    __shared__ double sred[kHornThreads / 32][kHornVals];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (phase == 1)
    {
        if (tid == 0)
        {
            double pc[3], qc[3], cnt;
            horn_centroids(part0 + (size_t)job * gridDim.x * kHornVals, gridDim.x, pc, qc, cnt);
            for (int d = 0; d < 3; d++) sc[d] = pc[d], sc[3 + d] = qc[d];
        }
        __syncthreads();
    }
    double acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t pos = blockIdx.x * kHornThreads + tid; pos < n; pos += gridDim.x * kHornThreads)
    {
        const PairRec r = prec[pos];
        if (!r.paired) continue;
        const float4 pl = __ldg(cvL.pts + pos);
        const double p[3] = {(double)pl.x, (double)pl.y, (double)pl.z};
        if (phase == 0)
        {
            acc[0] += p[0], acc[1] += p[1], acc[2] += p[2];
            acc[3] += r.q[0], acc[4] += r.q[1], acc[5] += r.q[2];
            acc[6] += 1.0;
            continue;
        }
        const double b[3] = {p[0] - sc[0], p[1] - sc[1], p[2] - sc[2]};           // local, centroid-relative
        const double a[3] = {r.q[0] - sc[3], r.q[1] - sc[4], r.q[2] - sc[5]};     // global
        const double bn = sqrt((b[0] * b[0] + b[1] * b[1]) + b[2] * b[2]);
        const double an = sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]);
        double       w = 1.0;
        if (P.use_scale_outlier_detector)
        {  // row M
            const double mx = bn > an ? bn : an, mn = bn > an ? an : bn;
            if (!(mn > 0.0) || mx / mn > P.scale_outlier_threshold) continue;
        }
        if (P.use_robust_kernel && bn > 0 && an > 0)
        {
            const double bu[3] = {b[0] / bn, b[1] / bn, b[2] / bn};
            double       rb[3];
            mat3_vec(J.R, bu, rb);
            double cs = (rb[0] * a[0] + rb[1] * a[1] + rb[2] * a[2]) / an;
            cs = cs > 1 ? 1 : (cs < -1 ? -1 : cs);
            const double ang = acos(cs);
            if (ang > P.robust_kernel_param)
            {
                const double e = ang - P.robust_kernel_param;
                w *= 1.0 / (1.0 + P.robust_kernel_scale * e * e);
            }
        }
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int k = 0; k < 3; k++) acc[i * 3 + k] += w * b[i] * a[k];
        acc[9] += 1.0;
    }
#pragma unroll
    for (int i = 0; i < 10; i++)
    {
        double v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if (lane == 0) sred[warp][i] = v;
    }
    __syncthreads();
    if (tid < kHornVals)
    {
        double v = 0;
        if (tid < 10)
            for (int w = 0; w < kHornThreads / 32; w++) v += sred[w][tid];
        part_out[((size_t)job * gridDim.x + blockIdx.x) * kHornVals + tid] = v;
    }
}

constexpr int kHornSolveThreads = 32;
__global__ void __launch_bounds__(kHornSolveThreads)
    horn_solve_kernel(JobDev* __restrict__ jobs, const double* __restrict__ part0,
                      const double* __restrict__ part1, uint32_t n_hpart, IcpDevParams P,
                      uint32_t* __restrict__ n_active)
{
    const uint32_t job = blockIdx.x;
    JobDev&        J = jobs[job];
    if (threadIdx.x == 0) J.next_item = 0;  // the next search draws items from 0 again
    if (J.status != 0) return;
    if (threadIdx.x == 0 && matcher_active(P, J.iter))
        J.rows_valid = 1, J.rows_pose_valid = 0;  // this iteration's search wrote them (pose not recorded on this path)
    const int tid = threadIdx.x;
    __shared__ double sH[kHornVals];
    // the matcher's moments (needed for the covariance at the end) are already in J.M: the fit kernel's tail
    if (tid < 10)
    {
        double        v = 0;
        const double* p = part1 + (size_t)job * n_hpart * kHornVals + tid;
        for (uint32_t c = 0; c < n_hpart; c++) v += p[(size_t)c * kHornVals];
        sH[tid] = v;
    }
    __syncthreads();
    if (tid != 0) return;
    double pc[3], qc[3], cnt;
    horn_centroids(part0 + (size_t)job * n_hpart * kHornVals, n_hpart, pc, qc, cnt);
    const uint32_t npair = (uint32_t)(cnt + 0.5);
    if (npair == 0)
    {
        J.n_pairings = 0;
        J.status = 1;
        J.term_reason = B200ICP_TERM_NO_PAIRINGS;
        atomicSub(n_active, 1u);
        return;
    }
    const uint32_t used = (uint32_t)(sH[9] + 0.5);
    if (used < 3)
    {  // Solver error: the pose is left as it was
        J.n_pairings = npair;
        J.status = 1;
        J.term_reason = B200ICP_TERM_SOLVER_ERROR;
        atomicSub(n_active, 1u);
        return;
    }
    double q[4];
    horn_quaternion(sH, q);
    Pose Tn;
    quaternion_to_R(q, Tn.R);
    double Rp[3];
    mat3_vec(Tn.R, pc, Rp);
    for (int d = 0; d < 3; d++) Tn.t[d] = qc[d] - Rp[d];
    finish_outer_iteration(J, Tn, npair, 1u, P, n_active, false);
}

// --------------------------------------------------------------- covariance
// mp2p_icp::covariance (row P / A.9): forward-difference Jacobian of the
// stacked residuals wrt (x,y,z,yaw,pitch,roll) at the solution, H = J^T J,
// cov = H^-1.  With linear-in-theta residuals J^T J = dTheta^T A dTheta.
// Runs right after the quality search of the same batch, on finished jobs
// that have not been evaluated yet, and marks them evaluated (the quality count
// must be taken once).
__global__ void __launch_bounds__(64) covariance_kernel(JobDev* __restrict__ jobs, IcpDevParams P)
{
    JobDev&   J = jobs[blockIdx.x];
    const int tid = threadIdx.x;
    __shared__ double sA[144], sD[72], sH[36];
    __shared__ int    s_go;
    if (tid == 0) s_go = (J.status != 0 && J.evaluated == 0) ? 1 : 0;
    __syncthreads();
    if (!s_go) return;
    if (tid == 0) J.evaluated = 1;
    const bool p2p = (P.matcher_kind == B200ICP_MATCHER_POINTS_DISTANCE);
    if (J.n_pairings == 0)
    {
        if (tid < 36) J.cov[tid] = 0.0;
        if (tid == 0) J.cov_singular = 1;
        return;
    }
    for (int e = tid; e < 144; e += blockDim.x) sA[e] = normal_matrix_at(J.M, p2p, e / 12, e % 12);
    if (tid < 6)
    {
        Pose T;
        for (int i = 0; i < 9; i++) T.R[i] = J.R[i];
        for (int i = 0; i < 3; i++) T.t[i] = J.t[i];
        double x0[6], xj[6];
        pose_to_ypr(T, x0);
        Pose T0, Tj;
        pose_from_ypr(x0, T0);
        for (int i = 0; i < 6; i++) xj[i] = x0[i];
        xj[tid] += P.cov_fd_step;
        pose_from_ypr(xj, Tj);
        for (int a = 0; a < 12; a++)
        {
            const int    i = a >> 2, l = a & 3;
            const double v1 = (l < 3) ? Tj.R[i * 3 + l] : Tj.t[i];
            const double v0 = (l < 3) ? T0.R[i * 3 + l] : T0.t[i];
            sD[a * 6 + tid] = (v1 - v0) / P.cov_fd_step;
        }
    }
    __syncthreads();
    if (tid < 36)
    {
        const int r = tid / 6, c = tid % 6;
        double    acc = 0;
        for (int a = 0; a < 12; a++)
        {
            double row = 0;
            for (int b = 0; b < 12; b++) row += sA[a * 12 + b] * sD[b * 6 + c];
            acc += sD[a * 6 + r] * row;
        }
        sH[tid] = acc;
    }
    __syncthreads();
    if (tid == 0)
    {
        double Hm[36], Ci[36];
        for (int i = 0; i < 36; i++) Hm[i] = 0.5 * (sH[i] + sH[(i % 6) * 6 + i / 6]);
        const int rank = inverse6_spd(Hm, Ci);
        for (int i = 0; i < 36; i++) J.cov[i] = (rank == 6) ? Ci[i] : 0.0;
        J.cov_singular = (rank == 6) ? 0u : 1u;
    }
}

__global__ void fill_u32_kernel(uint32_t* p, size_t n, uint32_t v)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void fill_f32_kernel(float* p, size_t n, float v)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void fill_u64_kernel(uint64_t* p, size_t n, uint64_t v)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// Arg-min merge of the partial results of disjoint map shards: one
// thread per query takes the k smallest of `parts` ascending key lists.
template <int K>
__global__ void __launch_bounds__(256)
    merge_keys_kernel(const uint64_t* __restrict__ parts, uint32_t nparts, size_t part_stride, size_t nq,
                      uint32_t k, uint64_t* __restrict__ out)
{
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    uint64_t key[K];
#pragma unroll
    for (int i = 0; i < K; i++) key[i] = ~0ull;
    for (uint32_t p = 0; p < nparts; p++)
    {
        const uint64_t* row = parts + (size_t)p * part_stride + q * k;
        for (uint32_t i = 0; i < k; i++)
        {
            const uint64_t kk = __ldg(row + i);
            if (!(kk < key[K - 1])) break;  // the lists are ascending
            topk_insert<K>(key, kk);
        }
    }
#pragma unroll
    for (int i = 0; i < K; i++)
        if ((uint32_t)i < k) out[q * k + i] = (key[i] == ~0ull) ? B200ICP_NO_KEY : key[i];
}

// =========================================================== host orchestration
static int wait_cloud(Workspace* ws, const b200icp_cloud* c)
{
    if (!c->indexed)
    {
        set_error("this cloud holds coordinates only (b200icp_cloud_upload_raw): it can feed b200icp_voxel_decimate, "
                  "not a search or a registration");
        return B200ICP_ERR_BAD_ARG;
    }
    B2_CUDA_TRY(cudaStreamWaitEvent(ws->stream, c->ready, 0));
    return B200ICP_OK;
}

// Switches (environment, read once).  The search stage is chosen per launch: the
// per-lane shell walk (tile_search.cuh) for single registrations and queries -- one
// wave of items, the kernel lasts as long as its slowest item, and the walk's items
// are the shortest -- and the item sweep (item_sweep.cuh) when kBatchJobs or more
// jobs share a launch (Monte-Carlo loop, loop-closure candidates): many waves, so
// throughput counts, and the sweep executes fewer, converged instructions
// (measured: C4 4263 against 2359 registrations/s).
// B200ICP_SEARCH=walk|item forces one of them, B200ICP_SEARCH=sweep selects the
// radius-wide tile sweep (sweep_search.cuh); all give the same results.
// B200ICP_WPI=1|2|4: the warps that share one item in the tile sweep.
constexpr size_t kBatchJobs = 8;  // jobs per launch from which the item sweep is the search stage
struct SearchConfig
{
    bool sweep = false;
    int  wpi = 0;  // 0 = automatic
    bool graphs = true;  // B200ICP_GRAPH=0: no CUDA-graph replay of single registrations
    bool tma = false;    // B200ICP_TMA=1 (with B200ICP_SEARCH=item): runs staged with cp.async.bulk + mbarriers
    bool seeds = true;   // B200ICP_SEED=0: searches never take their bound from the previous neighbour rows
    bool walk = false;   // B200ICP_SEARCH=walk: always the per-lane shell walk
    bool item = false;   // B200ICP_SEARCH=item: always the item sweep
};
static const SearchConfig& search_config()
{
    static const SearchConfig cfg = [] {
        SearchConfig c;
        if (const char* s = getenv("B200ICP_SEARCH")) c.sweep = (strcmp(s, "sweep") == 0), c.walk = (strcmp(s, "walk") == 0), c.item = (strcmp(s, "item") == 0);
        if (const char* w = getenv("B200ICP_WPI")) c.wpi = atoi(w);
        if (const char* g = getenv("B200ICP_GRAPH")) c.graphs = atoi(g) != 0;
        if (const char* g = getenv("B200ICP_SEED")) c.seeds = atoi(g) != 0;
        if (const char* g = getenv("B200ICP_TMA")) c.tma = atoi(g) != 0;
        return c;
    }();
    return cfg;
}

// CTAs per job of the FIT stage: one warp per chunk up to a full resident wave for a single job, fewer per job
// when many jobs share a launch.  Any value gives the same results (chunk partials).
static uint32_t fit_ctas_per_job(const ::b200icp* ctx, size_t max_points, size_t njobs, uint32_t chunk_items)
{
    static int resident = 0;
    if (resident == 0)
    {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fit_plane_kernel<6, false>, kChunk, 0) !=
                cudaSuccess || occ < 1)
            occ = 3;
        resident = occ;
    }
    const size_t items = std::max<size_t>(1, (max_points + kItem - 1) / kItem);
    const size_t chunks = (items + chunk_items - 1) / chunk_items;
    const size_t warps_per_cta = kChunk / 32;
    const size_t wave = (size_t)ctx->sm_count * resident;
    const size_t cap = std::max<size_t>(2, (2 * wave + njobs - 1) / njobs);
    return (uint32_t)std::max<size_t>(1, std::min({(chunks + warps_per_cta - 1) / warps_per_cta, cap, wave}));
}

// search stage over a table of jobs: grid = (items of the largest local cloud, jobs)
template <int K, class Epi>
static void launch_search_k(const ::b200icp* ctx, Workspace* ws, size_t max_points, size_t njobs,
                            const CloudView* d_clouds, JobDev* d_jobs, const IcpDevParams& D, float cap_d2,
                            int gate, const Epi& epi, const uint32_t* seed_nn = nullptr, uint32_t seed_k = 0)
{
    cudaStream_t        s = ws->stream;
    const SearchConfig& cfg = search_config();
    const size_t        items = std::max<size_t>(1, (max_points + kItem - 1) / kItem);
    const bool use_item = cfg.item || (!cfg.walk && !cfg.sweep && njobs >= kBatchJobs);
    if (use_item)
    {
        // resident CTAs on one SM (registers and the opted-in dynamic shared memory); every instantiation opts in once
        static int  resident = 0;
        static bool opted = false;  // per instantiation of this template
        const auto  kern = search_item_kernel<K, Epi>;
        if (!opted)
        {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ItemSmem));
            opted = true;
        }
        if (resident == 0)
        {
            int occ = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kChunk, sizeof(ItemSmem)) != cudaSuccess ||
                occ < 1)
                occ = 3;
            resident = occ;
        }
        const size_t   warps = kChunk / 32;
        const size_t   wave = (size_t)ctx->sm_count * resident;
        const size_t   cap = std::max<size_t>(4, (2 * wave + njobs - 1) / njobs);
        const uint32_t G = (uint32_t)std::max<size_t>(1, std::min({(items + warps - 1) / warps, cap, wave}));
        kern<<<dim3(G, (unsigned)njobs), kChunk, sizeof(ItemSmem), s>>>(d_clouds, d_jobs, D, cap_d2, gate, seed_nn,
                                                                       seed_k, cfg.tma ? 1 : 0, epi);
    }
    else if (!cfg.sweep)
    {
        // resident CTAs of the walk on one SM (registers and shared memory)
        static int resident = 0;
        if (resident == 0)
        {
            int occ = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, search_tile_kernel<6, NnWriter>, kChunk, 0) !=
                    cudaSuccess || occ < 1)
                occ = 4;
            resident = occ;
        }
        const size_t   warps = kChunk / 32;
        const size_t   wave = (size_t)ctx->sm_count * resident;
        const size_t   cap = std::max<size_t>(4, (2 * wave + njobs - 1) / njobs);
        const uint32_t G = (uint32_t)std::max<size_t>(1, std::min({(items + warps - 1) / warps, cap, wave}));
        search_tile_kernel<K, Epi><<<dim3(G, (unsigned)njobs), kChunk, 0, s>>>(d_clouds, d_jobs, D, cap_d2, gate,
                                                                              seed_nn, seed_k, epi);
    }
    else
    {
        // one CTA per item up to a few waves per launch; more items are taken grid-stride
        const size_t   budget = std::max<size_t>(64, ((size_t)ctx->sm_count * 64 + njobs - 1) / njobs);
        const uint32_t G = (uint32_t)std::min(items, budget);
        int            wpi = cfg.wpi;
        if (wpi != 1 && wpi != 2 && wpi != 4) wpi = (items * njobs >= (size_t)ctx->sm_count * 256) ? 1 : 4;
        const dim3 grid(G, (unsigned)njobs);
        if (wpi == 1)
            search_sweep_kernel<K, 1, Epi><<<grid, 32, 0, s>>>(d_clouds, d_jobs, D, cap_d2, gate, epi);
        else if (wpi == 2)
            search_sweep_kernel<K, 2, Epi><<<grid, 64, 0, s>>>(d_clouds, d_jobs, D, cap_d2, gate, epi);
        else
            search_sweep_kernel<K, 4, Epi><<<grid, 128, 0, s>>>(d_clouds, d_jobs, D, cap_d2, gate, epi);
    }
    ws->launches++;
}

// K of the matcher's search for the configured matcher / knn
static int matcher_k(const IcpDevParams& D)
{
    if (D.matcher_kind == B200ICP_MATCHER_POINTS_DISTANCE) return 1;
    return D.knn <= 4 ? 4 : (D.knn <= 6 ? 6 : 8);
}

// the matcher's search: neighbour rows [job base + sorted position][K]
static void launch_match_search(const ::b200icp* ctx, Workspace* ws, const IcpDevParams& D, size_t max_points,
                                size_t njobs, const CloudView* d_clouds, JobDev* d_jobs, uint32_t* d_nn)
{
    const int           K = matcher_k(D);
    const bool          seeds = search_config().seeds;  // B200ICP_SEED=0: every search from the radius cap
    const NnWriter      w = {d_nn, nullptr, (uint32_t)K, 0u};
    if (K == 1)
        launch_search_k<1>(ctx, ws, max_points, njobs, d_clouds, d_jobs, D, D.thr2, 1, w, seeds ? d_nn : nullptr,
                           (uint32_t)K);
    else if (K == 4)
        launch_search_k<4>(ctx, ws, max_points, njobs, d_clouds, d_jobs, D, D.thr2, 1, w, seeds ? d_nn : nullptr,
                           (uint32_t)K);
    else if (K == 6)
        launch_search_k<6>(ctx, ws, max_points, njobs, d_clouds, d_jobs, D, D.thr2, 1, w, seeds ? d_nn : nullptr,
                           (uint32_t)K);
    else
        launch_search_k<8>(ctx, ws, max_points, njobs, d_clouds, d_jobs, D, D.thr2, 1, w, seeds ? d_nn : nullptr,
                           (uint32_t)K);
}

// tail_mode 1: the warp that completes a job's moment reduction runs the Gauss-Newton solver and the end of the
// outer iteration; 0: the reduced moments are left in JobDev::M (Horn path, matcher hook)
template <bool WRITE>
static void launch_fit(Workspace* ws, const IcpDevParams& D, dim3 grid, const CloudView* d_clouds, JobDev* d_jobs,
                       const uint32_t* d_nn, const FitBuffers& fb, uint32_t chunk_items, int tail_mode,
                       const MatchOut& mo, PairRec* d_pairs, uint32_t* d_active, uint32_t chunk_begin = 0,
                       uint32_t chunk_end = 0xFFFFFFFFu)
{
    cudaStream_t s = ws->stream;
    const int    K = matcher_k(D);
    if (K == 1)
        fit_p2p_kernel<WRITE><<<grid, kChunk, 0, s>>>(d_clouds, d_jobs, d_nn, fb, chunk_items, tail_mode, D, mo, d_pairs,
                                                      d_active, chunk_begin, chunk_end);
    else if (K == 4)
        fit_plane_kernel<4, WRITE><<<grid, kChunk, 0, s>>>(d_clouds, d_jobs, d_nn, fb, chunk_items, tail_mode, D, mo,
                                                           d_pairs, d_active, chunk_begin, chunk_end);
    else if (K == 6)
        fit_plane_kernel<6, WRITE><<<grid, kChunk, 0, s>>>(d_clouds, d_jobs, d_nn, fb, chunk_items, tail_mode, D, mo,
                                                           d_pairs, d_active, chunk_begin, chunk_end);
    else
        fit_plane_kernel<8, WRITE><<<grid, kChunk, 0, s>>>(d_clouds, d_jobs, d_nn, fb, chunk_items, tail_mode, D, mo,
                                                           d_pairs, d_active, chunk_begin, chunk_end);
    ws->launches++;
}

// Sizes baked into a replayed launch sequence (grids, offsets inside the scratch allocation) are taken from a
// BUCKET of the point count -- four steps per octave -- so that scans whose size changes from frame to frame
// (real LiDAR, decimated clouds) reuse one graph; the kernels read the true sizes from the device records.
static size_t size_bucket(size_t n)
{
    if (n <= 1024) return 1024;
    size_t step = 1;
    while ((step << 3) < n) step <<= 1;  // step = 2^(floor(log2(n - 1)) - 2)
    return (n + step - 1) / step * step;
}

// items of the local cloud per chunk partial: 2 for a single registration (all warps of the machine get work:
// 120k points -> 1875 chunks), more when many jobs share a launch (parallelism comes from the jobs and the
// partials are 1.5 KB each)
static uint32_t fit_chunk_items(size_t max_points, size_t njobs)
{
    if (njobs == 1) return 2;
    const size_t items = (max_points + kItem - 1) / kItem;
    return (uint32_t)std::min<size_t>(16, std::max<size_t>(2, items / 64));
}

static int check_supported(const IcpDevParams& P)
{
    if ((P.matcher_kind != B200ICP_MATCHER_POINT2PLANE && P.matcher_kind != B200ICP_MATCHER_POINTS_DISTANCE) ||
        (P.solver_kind != B200ICP_SOLVER_GAUSS_NEWTON && P.solver_kind != B200ICP_SOLVER_HORN))
    {
        set_error("unknown matcher_kind=%d / solver_kind=%d", P.matcher_kind, P.solver_kind);
        return B200ICP_ERR_UNSUPPORTED;
    }
    if (P.matcher_kind == B200ICP_MATCHER_POINT2PLANE && (P.knn < 1 || P.knn > B200ICP_MAX_KNN))
    {
        set_error("knn=%u outside [1,%d]", P.knn, B200ICP_MAX_KNN);
        return B200ICP_ERR_UNSUPPORTED;
    }
    if (P.use_robust_kernel && P.solver_kind == B200ICP_SOLVER_GAUSS_NEWTON)
    {
        set_error("use_robust_kernel=true is not available with Solver_GaussNewton");
        return B200ICP_ERR_UNSUPPORTED;
    }
    return B200ICP_OK;
}

// Jobs are processed in waves that bound gridDim.y.
static int run_wave(::b200icp* ctx, Workspace* ws, const IcpDevParams& D, size_t n,
                    const b200icp_cloud* const* from, const b200icp_cloud* const* to, const double* guesses,
                    b200icp_result_t* out)
{
    cudaStream_t        s = ws->stream;
    // unique cloud table
    std::map<const b200icp_cloud*, uint32_t> cmap;
    std::vector<CloudView>                   views;
    size_t                                   max_points = 1;
    uint64_t                                 total_queries = 0;
    auto add = [&](const b200icp_cloud* c) -> uint32_t {
        auto it = cmap.find(c);
        if (it != cmap.end()) return it->second;
        const uint32_t id = (uint32_t)views.size();
        views.push_back(c->view());
        cmap[c] = id;
        return id;
    };
    std::vector<JobDev> hjobs(n);
    for (size_t j = 0; j < n; j++)
    {
        JobDev& J = hjobs[j];
        memset(&J, 0, sizeof(J));
        Pose T;
        pose_from_ypr(guesses + 6 * j, T);
        memcpy(J.R, T.R, sizeof(T.R)), memcpy(J.t, T.t, sizeof(T.t));
        memcpy(J.Rprev, T.R, sizeof(T.R)), memcpy(J.tprev, T.t, sizeof(T.t));
        J.from_cloud = add(from[j]);
        J.to_cloud = add(to[j]);
        J.pair_base = (uint32_t)total_queries;
        max_points = std::max(max_points, to[j]->n);
        total_queries += to[j]->n;
    }
    // chunk partials of the fit stage: where each job's partials / group partials / tickets start
    const uint32_t chunk_items = fit_chunk_items(max_points, n);
    uint64_t       total_chunks = 0, total_groups = 0;
    for (size_t j = 0; j < n; j++)
    {
        const uint64_t items = (to[j]->n + kItem - 1) / kItem;
        const uint64_t chunks = (items + chunk_items - 1) / chunk_items;
        hjobs[j].chunk_base = (uint32_t)total_chunks;
        hjobs[j].group_base = (uint32_t)total_groups;
        total_chunks += chunks;
        total_groups += (chunks + kFitGroup - 1) / kFitGroup;
    }
    if (total_chunks >= 0xFFFFFFFFull)
    {
        set_error("too many points in one batch");
        return B200ICP_ERR_BAD_ARG;
    }
    // a single registration lays its buffers out for the size BUCKET of its clouds: offsets and grids then stay
    // the same from scan to scan and the replayed graph with them
    const size_t lay_points = (n == 1) ? size_bucket(max_points) : max_points;
    const size_t lay_queries = (n == 1) ? size_bucket(total_queries) : (size_t)total_queries;
    const size_t lay_chunks = (n == 1) ? ((lay_points + kItem - 1) / kItem + chunk_items - 1) / chunk_items
                                       : (size_t)total_chunks;
    const size_t lay_groups = (n == 1) ? (lay_chunks + kFitGroup - 1) / kFitGroup : (size_t)total_groups;
    const bool horn = (D.solver_kind == B200ICP_SOLVER_HORN);
    if (horn && total_queries >= 0xFFFFFFFFull)
    {
        set_error("too many points in one batch for the Horn pair buffer");
        return B200ICP_ERR_BAD_ARG;
    }
    for (auto& kv : cmap)
        if (int r = wait_cloud(ws, kv.first)) return r;
    const uint32_t G = fit_ctas_per_job(ctx, lay_points, n, chunk_items);
    const int      K = matcher_k(D);

    const uint32_t GH = std::max<uint32_t>(1, std::min<uint32_t>(G, (uint32_t)((lay_points + 1023) / 1024)));
    PairRec*       d_pairs = nullptr;
    double *       d_h0 = nullptr, *d_h1 = nullptr;
    uint32_t*      d_nn = nullptr;
    FitBuffers     fb = {nullptr, nullptr, nullptr};
    Carver sz(nullptr);
    auto layout = [&](Carver& k, CloudView*& dc, JobDev*& dj, uint32_t*& da) {
        dc = k.take<CloudView>(views.size());
        dj = k.take<JobDev>(n);
        da = k.take<uint32_t>(4);
        fb.tickets = k.take<uint32_t>(lay_groups ? lay_groups : 1);
        fb.gpartials = k.take<double>((lay_groups ? lay_groups : 1) * (size_t)kNumMoments);
        fb.partials = k.take<double>((lay_chunks ? lay_chunks : 1) * (size_t)kNumMoments);
        d_nn = k.take<uint32_t>((lay_queries ? lay_queries : 1) * (size_t)K);
        if (horn)
        {
            d_pairs = k.take<PairRec>(lay_queries ? lay_queries : 1);
            d_h0 = k.take<double>((size_t)n * GH * kHornVals);
            d_h1 = k.take<double>((size_t)n * GH * kHornVals);
        }
    };
    CloudView* d_clouds;
    JobDev*    d_jobs;
    uint32_t*  d_active;
    layout(sz, d_clouds, d_jobs, d_active);
    if (int r = ws->reserve_device(sz.off)) return r;
    Carver real(ws->d_scratch);
    layout(real, d_clouds, d_jobs, d_active);
    const size_t ticket_bytes = (lay_groups ? lay_groups : 1) * sizeof(uint32_t);
    // pinned staging: views | jobs | active flags (2 slots) | n_active init
    const size_t off_jobs = align_up(views.size() * sizeof(CloudView));
    const size_t off_flags = off_jobs + align_up(n * sizeof(JobDev));
    if (int r = ws->reserve_pinned(off_flags + 256)) return r;
    char*      hp = (char*)ws->h_pinned;
    CloudView* h_views = (CloudView*)hp;
    JobDev*    h_jobs = (JobDev*)(hp + off_jobs);
    uint32_t*  h_flags = (uint32_t*)(hp + off_flags);
    memcpy(h_views, views.data(), views.size() * sizeof(CloudView));
    memcpy(h_jobs, hjobs.data(), n * sizeof(JobDev));
    h_flags[0] = h_flags[1] = 0xFFFFFFFFu;
    h_flags[2] = (uint32_t)n;
    cudaError_t setup_err = cudaSuccess;
    auto enqueue_setup = [&]() {
        cudaError_t e = cudaMemcpyAsync(d_clouds, h_views, views.size() * sizeof(CloudView), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_jobs, h_jobs, n * sizeof(JobDev), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_active, h_flags + 2, sizeof(uint32_t), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemsetAsync(fb.tickets, 0, ticket_bytes, s);  // arrival counters of the fit stage
        if (e != cudaSuccess) setup_err = e;
    };

    const bool prof = ctx->profile_on;
    if (prof)
        if (int r = ws->reserve_prof_events(2 * (size_t)D.max_iterations + 2)) return r;

    const dim3     mgrid(G, (unsigned)n);
    const MatchOut no_out = {nullptr, nullptr, nullptr, nullptr, nullptr};
    // Iterations are enqueued in batches and the host looks at the active-job
    // counter one batch behind, so the stream never drains; launches enqueued
    // after the last job has finished exit at once but still cost ~2.5 us each.
    // A single registration (the odometry call) therefore sizes its first batch
    // from the previous call on this ICP object -- consecutive scans need about
    // the same number of iterations -- and checks it right away.
    cudaError_t eval_err = cudaSuccess;
    auto evaluate = [&]() {
        launch_search_k<1>(ctx, ws, lay_points, n, d_clouds, d_jobs, D, D.q_thr2, 2, HitCounter{D.q_thr2},
                           search_config().seeds ? d_nn : nullptr, (uint32_t)K);
        covariance_kernel<<<(unsigned)n, 64, 0, s>>>(d_jobs, D);
        ws->launches++;
        const cudaError_t e = cudaMemcpyAsync(h_jobs, d_jobs, n * sizeof(JobDev), cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) eval_err = e;
    };
    const uint32_t kBatch = 4;
    const bool     predict = (n == 1) && ctx->expected_runs.load() > 0;
    // one more than the previous registration needed, rounded up to a few fixed sizes: every distinct size is
    // a graph of its own to instantiate (milliseconds), an iteration enqueued in vain costs two empty launches
    auto first_bucket = [](int runs) -> uint32_t {
        static const int sizes[] = {3, 4, 5, 6, 8, 10, 12, 16, 20, 24};
        for (int sz : sizes)
            if (runs <= sz) return (uint32_t)sz;
        return 24u;
    };
    const uint32_t first = predict ? first_bucket(ctx->expected_runs.load() + 1) : kBatch;
    uint32_t       enq = 0, batch = 0;
    bool           finished = false;

    // A single predicted registration without profiling replays its first batch -- job upload, `first`
    // iterations, evaluation, record and flag download -- as ONE CUDA graph launch: the ~25 driver calls per
    // registration become one, which is what the host side of a short registration costs.  The graph is
    // cached per workspace and keyed by everything baked into it.
    bool graphed = false;
    if (predict && !prof && !horn && search_config().graphs && first <= D.max_iterations)
    {
        uint64_t phash = 1469598103934665603ull;  // the parameter block is baked into the kernel nodes
        for (size_t b = 0; b < sizeof(D); b++) phash = (phash ^ reinterpret_cast<const unsigned char*>(&D)[b]) * 1099511628211ull;
        const AlignGraphKey key = {ws->d_scratch, ws->h_pinned, (uint64_t)lay_points, (uint64_t)lay_queries,
                                   (uint32_t)views.size(), G, first, (uint32_t)matcher_k(D), phash};
        cudaGraphExec_t exec = ws->find_align_graph(key);
        if (!exec && !ws->graph_failed)
        {
            cudaGraph_t graph = nullptr;
            const uint64_t launches_before = ws->launches;
            bool ok = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok)
            {
                enqueue_setup();
                for (uint32_t i = 0; i < first; i++)
                {
                    launch_match_search(ctx, ws, D, lay_points, n, d_clouds, d_jobs, d_nn);
                    launch_fit<false>(ws, D, mgrid, d_clouds, d_jobs, d_nn, fb, chunk_items, 1, no_out, d_pairs, d_active);
                }
                evaluate();
                cudaMemcpyAsync(h_flags, d_active, sizeof(uint32_t), cudaMemcpyDeviceToHost, s);
                ok = cudaStreamEndCapture(s, &graph) == cudaSuccess && graph != nullptr &&
                     setup_err == cudaSuccess && eval_err == cudaSuccess;
                if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
                if (graph) cudaGraphDestroy(graph);
            }
            const uint64_t per_launch = ws->launches - launches_before;
            ws->launches = launches_before;  // captured, not launched
            if (ok)
                ws->store_align_graph(key, exec, per_launch);
            else
            {
                cudaGetLastError();
                exec = nullptr;
                ws->graph_failed = true;  // this workspace keeps to plain launches
                setup_err = eval_err = cudaSuccess;
            }
        }
        if (exec)
        {
            B2_CUDA_TRY(cudaGraphLaunch(exec, s));
            ws->launches += ws->align_graph_launches(key);
            {
                std::lock_guard<std::mutex> lk(ctx->mtx);
                ctx->prof.graph_replays++;
            }
            B2_CUDA_TRY(cudaStreamSynchronize(s));
            enq = first, batch = 1, graphed = true;
            finished = (h_flags[0] == 0);
        }
    }
    if (!graphed)
    {
        enqueue_setup();
        B2_CUDA_TRY(setup_err);
    }
    while (enq < D.max_iterations && !finished)
    {
        const uint32_t want = (batch == 0) ? first : (predict ? 2u : kBatch);
        const uint32_t todo = std::min(want, D.max_iterations - enq);
        for (uint32_t i = 0; i < todo; i++, enq++)
        {
            if (prof) B2_CUDA_TRY(cudaEventRecord(ws->prof_ev[4 * enq + 0], s));
            launch_match_search(ctx, ws, D, lay_points, n, d_clouds, d_jobs, d_nn);
            if (prof) B2_CUDA_TRY(cudaEventRecord(ws->prof_ev[4 * enq + 1], s));
            // Gauss-Newton: the fit kernel's last warp per job solves and ends the iteration; Horn: the fit leaves
            // the matcher's moments in the job record and the closed-form solver follows
            launch_fit<false>(ws, D, mgrid, d_clouds, d_jobs, d_nn, fb, chunk_items, horn ? 0 : 1, no_out, d_pairs,
                              d_active);
            if (prof) B2_CUDA_TRY(cudaEventRecord(ws->prof_ev[4 * enq + 2], s));
            if (horn)
            {
                const dim3 hgrid(GH, (unsigned)n);
                horn_sum_kernel<<<hgrid, kHornThreads, 0, s>>>(d_clouds, d_jobs, d_pairs, nullptr, d_h0, D, 0);
                horn_sum_kernel<<<hgrid, kHornThreads, 0, s>>>(d_clouds, d_jobs, d_pairs, d_h0, d_h1, D, 1);
                horn_solve_kernel<<<(unsigned)n, kHornSolveThreads, 0, s>>>(d_jobs, d_h0, d_h1, GH, D, d_active);
                ws->launches += 3;
            }
            if (prof) B2_CUDA_TRY(cudaEventRecord(ws->prof_ev[4 * enq + 3], s));
        }
        // Single registration: the evaluation of the job if this batch has finished it
        // (QualityEvaluator_PairedRatio: one more search, k = 1, its own radius; then the covariance) and the
        // job record are enqueued without waiting for the host to learn that it finished -- an unfinished job
        // makes these launches exit at once.  Saves one host round trip per registration.
        if (n == 1) evaluate();
        // the host stays one batch ahead of the device: it only looks at the
        // active-job counter of the PREVIOUS batch, so the stream never drains
        const int slot = batch & 1;
        B2_CUDA_TRY(cudaMemcpyAsync(h_flags + slot, d_active, sizeof(uint32_t),
                                    cudaMemcpyDeviceToHost, s));
        B2_CUDA_TRY(cudaEventRecord(ws->ev[slot], s));
        if (batch >= 1)
        {
            B2_CUDA_TRY(cudaEventSynchronize(ws->ev[slot ^ 1]));
            if (h_flags[slot ^ 1] == 0) finished = true;
        }
        else if (predict)
        {  // most likely the whole registration: look at it now
            B2_CUDA_TRY(cudaEventSynchronize(ws->ev[slot]));
            if (h_flags[slot] == 0) finished = true;
        }
        batch++;
    }
    if (n != 1) evaluate();  // batches: once, after the last job has finished
    B2_CUDA_TRY(eval_err);
    B2_CUDA_TRY(cudaStreamSynchronize(s));
    B2_CUDA_TRY(cudaGetLastError());

    uint32_t max_runs = 0;
    for (size_t j = 0; j < n; j++)
    {
        const JobDev&     J = h_jobs[j];
        b200icp_result_t& r = out[j];
        memset(&r, 0, sizeof(r));
        Pose T;
        memcpy(T.R, J.R, sizeof(T.R)), memcpy(T.t, J.t, sizeof(T.t));
        pose_to_ypr(T, r.pose);
        memcpy(r.R, J.R, sizeof(r.R)), memcpy(r.t, J.t, sizeof(r.t));
        memcpy(r.cov, J.cov, sizeof(r.cov));
        r.quality = to[j]->n ? (double)J.quality_count / (double)to[j]->n : 0.0;
        r.n_iterations = J.iter;
        r.termination_reason = J.term_reason;
        r.n_pairings = J.n_pairings;
        r.cov_singular = J.cov_singular;
        const uint32_t runs = std::min(J.iter + (J.term_reason == B200ICP_TERM_MAX_ITERATIONS ? 0u : 1u),
                                       D.max_iterations);
        max_runs = std::max(max_runs, runs);
    }
    // the next registration's first batch is sized from this one -- unless this one ended in a recognised cycle: an
    // outlier that says nothing about the next scan pair
    if (n == 1 && h_jobs[0].cycle_at == 0) ctx->expected_runs.store((int)max_runs);
    if (getenv("B200ICP_DBG_TAIL"))
        fprintf(stderr, "[dbg tail] job 0: final moment sum %u cycles, GN loop %u cycles (%u inner), end of iteration %u "
                        "cycles; %u outer iterations, %u inner in total\n",
                h_jobs[0].dbg_tail[0], h_jobs[0].dbg_tail[1], h_jobs[0].dbg_tail[3], h_jobs[0].dbg_tail[2],
                h_jobs[0].iter, h_jobs[0].inner_iters_total);
    if (prof)
    {
        double mm = 0, fm = 0, sm = 0;
        for (uint32_t i = 0; i < std::min(max_runs, enq); i++)
        {
            float a = 0, b = 0, c = 0;
            B2_CUDA_TRY(cudaEventElapsedTime(&a, ws->prof_ev[4 * i + 0], ws->prof_ev[4 * i + 1]));
            B2_CUDA_TRY(cudaEventElapsedTime(&c, ws->prof_ev[4 * i + 1], ws->prof_ev[4 * i + 2]));
            B2_CUDA_TRY(cudaEventElapsedTime(&b, ws->prof_ev[4 * i + 2], ws->prof_ev[4 * i + 3]));
            mm += a, fm += c, sm += b;
        }
        std::lock_guard<std::mutex> lk(ctx->mtx);
        ctx->prof.match_launches += std::min(max_runs, enq);
        ctx->prof.match_ms += mm;
        // queries examined by those launches (single job: exact; batches: upper bound)
        ctx->prof.match_queries += (uint64_t)std::min(max_runs, enq) * total_queries;
        ctx->prof.fit_launches += std::min(max_runs, enq);
        ctx->prof.fit_ms += fm;
        ctx->prof.solve_launches += std::min(max_runs, enq);
        ctx->prof.solve_ms += sm;
    }
    return B200ICP_OK;
}

// mp2p_icp::Parameters of ONE call over the object's own (LidarOdometry.cpp:869-871 hands in.icp_params to
// align() next to the shared ICP object, whose matchers / solvers / quality evaluators stay as configured)
static IcpDevParams merged_params(const ::b200icp* ctx, const b200icp_call_params_t* call)
{
    IcpDevParams D = ctx->D;
    if (call)
    {
        D.max_iterations = call->max_iterations;
        D.min_abs_step_trans = call->min_abs_step_trans;
        D.min_abs_step_rot = call->min_abs_step_rot;
        D.use_scale_outlier_detector = call->use_scale_outlier_detector;
        D.scale_outlier_threshold = call->scale_outlier_threshold;
        D.use_robust_kernel = call->use_robust_kernel;
        D.robust_kernel_param = call->robust_kernel_param;
        D.robust_kernel_scale = call->robust_kernel_scale;
    }
    return D;
}

int run_align_batch(::b200icp* ctx, size_t n, const b200icp_cloud* const* from,
                    const b200icp_cloud* const* to, const double* guesses, const b200icp_call_params_t* call,
                    b200icp_result_t* out)
{
    const IcpDevParams D = merged_params(ctx, call);
    if (int r = check_supported(D)) return r;
    if (n == 0) return B200ICP_OK;
    Lease L(ctx);
    if (!L.ws) return B200ICP_ERR_CUDA;
    // waves bound gridDim.y (<= 65535) and the neighbour buffer (K x 4 B per query)
    const size_t kWaveJobs = 32768, kWaveQueries = (size_t)1 << 28;
    for (size_t i = 0; i < n;)
    {
        size_t cnt = 0, queries = 0;
        while (i + cnt < n && cnt < kWaveJobs && (cnt == 0 || queries + to[i + cnt]->n <= kWaveQueries))
            queries += to[i + cnt]->n, cnt++;
        if (int r = run_wave(ctx, L.ws, D, cnt, from + i, to + i, guesses + 6 * i, out + i)) return r;
        i += cnt;
    }
    return B200ICP_OK;
}

// one (from, to, pose) job laid out and uploaded; `extra` carves the caller's
// own arrays out of the same scratch allocation
struct SingleJob
{
    CloudView* d_clouds = nullptr;
    JobDev*    d_jobs = nullptr;
    Pose       T;
};

template <typename F>
static int single_job_setup(Workspace* ws, const b200icp_cloud* from, const b200icp_cloud* to,
                            const double* pose6, uint32_t iter, SingleJob& sj, F&& extra)
{
    cudaStream_t s = ws->stream;
    if (int r = wait_cloud(ws, from)) return r;
    if (int r = wait_cloud(ws, to)) return r;
    auto layout = [&](Carver& c) {
        sj.d_clouds = c.take<CloudView>(2);
        sj.d_jobs = c.take<JobDev>(1);
        extra(c);
    };
    Carver sz(nullptr);
    layout(sz);
    if (int r = ws->reserve_device(sz.off)) return r;
    Carver real(ws->d_scratch);
    layout(real);
    const size_t off_job = align_up(2 * sizeof(CloudView));
    if (int r = ws->reserve_pinned(off_job + align_up(sizeof(JobDev)) + 64)) return r;
    CloudView* hv = (CloudView*)ws->h_pinned;
    JobDev*    hj = (JobDev*)((char*)ws->h_pinned + off_job);
    hv[0] = from->view(), hv[1] = to->view();
    memset(hj, 0, sizeof(JobDev));
    const double ident[6] = {0, 0, 0, 0, 0, 0};
    pose_from_ypr(pose6 ? pose6 : ident, sj.T);
    memcpy(hj->R, sj.T.R, sizeof(sj.T.R)), memcpy(hj->t, sj.T.t, sizeof(sj.T.t));
    hj->from_cloud = 0, hj->to_cloud = 1;
    hj->iter = iter;
    B2_CUDA_TRY(cudaMemcpyAsync(sj.d_clouds, hv, 2 * sizeof(CloudView), cudaMemcpyHostToDevice, s));
    B2_CUDA_TRY(cudaMemcpyAsync(sj.d_jobs, hj, sizeof(JobDev), cudaMemcpyHostToDevice, s));
    return B200ICP_OK;
}

// Uncapped search (max_dist = +inf), second pass: the rows the radius-capped search left incomplete -- queries with
// fewer than k points within the index's block edge -- are completed EXACTLY by one warp per query over the whole
// reference cloud: lane l takes every 32nd group of 8 sorted points, tests the group's box against the best bound
// the warp has (the smallest k-th-best any lane holds: that lane alone has k points within it) and evaluates the
// groups that survive; the 32 sorted lists are merged with k warp minima.  Same keys, same tie rule.
template <int K>
__global__ void __launch_bounds__(256)
    knn_complete_kernel(const CloudView* __restrict__ clouds, const JobDev* __restrict__ jobs, uint32_t k,
                        uint32_t* __restrict__ idx, float* __restrict__ d2)
{
    const JobDev&   J = jobs[0];
    const CloudView cvL = clouds[J.to_cloud];
    const CloudView cvG = clouds[J.from_cloud];
    const uint32_t  nq = cvL.grid->n_valid, nr = cvG.grid->n_valid;
    const int       lane = threadIdx.x & 31;
    const unsigned  full = 0xFFFFFFFFu;
    const uint32_t  warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t  kk_need = min(k, nr);
    double          Rt[12];
#pragma unroll
    for (int i = 0; i < 12; i++) Rt[i] = (i < 9) ? J.R[i] : J.t[i - 9];
    for (uint32_t pos = warp; pos < nq; pos += nwarps)
    {
        const float4   pl = __ldg(cvL.pts + pos);
        const uint32_t orig = __float_as_uint(pl.w);
        if (kk_need == 0 || idx[(size_t)orig * k + (kk_need - 1)] != kInvalid) continue;  // complete already (uniform)
        double gx, gy, gz;
        transform_point(Rt, pl, gx, gy, gz);
        const float qx = (float)gx, qy = (float)gy, qz = (float)gz;
        if (!((fabsf(qx) <= FLT_MAX) && (fabsf(qy) <= FLT_MAX) && (fabsf(qz) <= FLT_MAX))) continue;
        const uint64_t sent = sentinel_key(INFINITY);
        uint64_t       key[K];
#pragma unroll
        for (int i = 0; i < K; i++) key[i] = sent;
        const uint32_t ngroups = (nr + kGroup - 1) / kGroup;
        float          bound = INFINITY;  // warp-wide bound of the k-th neighbour (k = the launch's K here)
        for (uint32_t g0 = 0; g0 < ngroups; g0 += 32)
        {
            const uint32_t g = g0 + lane;
            if (g < ngroups)
            {
                const float4 lo = __ldg(cvG.gbox + 2 * g), hi = __ldg(cvG.gbox + 2 * g + 1);
                if (!(box_lower_d2(qx, qy, qz, lo, hi) > fminf(bound, key_d2(key[K - 1]))))
                {
                    const uint32_t base = g * kGroup;
#pragma unroll
                    for (int u = 0; u < kGroup; u++)
                        if (base + u < nr)
                        {
                            const float4   c = __ldg(cvG.pts + base + u);
                            const uint64_t kk = make_key(dist2(qx, qy, qz, c), __float_as_uint(c.w));
                            if (kk < key[K - 1]) topk_insert<K>(key, kk);
                        }
                }
            }
            // a lane whose list is full bounds the k-th neighbour of the union (ties: the bound is not strict)
            bound = __uint_as_float(__reduce_min_sync(full, __float_as_uint(key_d2(key[K - 1]))));
        }
        // merge: the smallest head of the 32 sorted lists, K times (keys are unique: they hold the point's index)
        uint64_t res[K];
#pragma unroll
        for (int i = 0; i < K; i++)
        {
            const uint32_t hi = __reduce_min_sync(full, (uint32_t)(key[0] >> 32));
            const uint32_t lo = __reduce_min_sync(full, ((uint32_t)(key[0] >> 32) == hi) ? (uint32_t)key[0] : 0xFFFFFFFFu);
            const uint64_t m = ((uint64_t)hi << 32) | lo;
            res[i] = m;
            if (key[0] == m && m != sent)
            {
#pragma unroll
                for (int j = 0; j + 1 < K; j++) key[j] = key[j + 1];
                key[K - 1] = sent;
            }
        }
        if (lane == 0)
        {
#pragma unroll
            for (int i = 0; i < K; i++)
                if ((uint32_t)i < k)
                {
                    const bool ok = res[i] != sent;
                    idx[(size_t)orig * k + i] = ok ? key_idx(res[i]) : kInvalid;
                    if (d2) d2[(size_t)orig * k + i] = ok ? key_d2(res[i]) : INFINITY;
                }
        }
    }
}

int run_knn(::b200icp* ctx, const b200icp_cloud* ref, const b200icp_cloud* q, const double* pose6,
            uint32_t k, float max_dist, uint32_t* idx_out, float* d2_out)
{
    if (k < 1 || k > B200ICP_MAX_KNN)
    {
        set_error("k=%u outside [1,%d]", k, B200ICP_MAX_KNN);
        return B200ICP_ERR_BAD_ARG;
    }
    if (!(max_dist > 0))  // NaN included
    {
        set_error("max_dist must be positive (+inf = uncapped search)");
        return B200ICP_ERR_BAD_ARG;
    }
    // uncapped: the capped search at the reference index's own radius first, the incomplete rows completed after
    const bool uncapped = std::isinf(max_dist);
    if (uncapped) max_dist = ref->cell_req / 1.002f;  // the radius the reference cloud was indexed for
    Lease L(ctx);
    if (!L.ws) return B200ICP_ERR_CUDA;
    Workspace*   ws = L.ws;
    cudaStream_t s = ws->stream;
    const size_t nq = q->n;
    if (nq == 0) return B200ICP_OK;
    uint32_t* d_idx = nullptr;
    float*    d_d2 = nullptr;
    SingleJob sj;
    if (int r = single_job_setup(ws, ref, q, pose6, 0, sj, [&](Carver& c) {
            d_idx = c.take<uint32_t>(nq * k);
            d_d2 = c.take<float>(nq * k);
        }))
        return r;
    const float cap_d2 = max_dist * max_dist;
    const int   fb = (int)((nq * k + 255) / 256);
    fill_u32_kernel<<<fb, 256, 0, s>>>(d_idx, nq * k, kInvalid);
    fill_f32_kernel<<<fb, 256, 0, s>>>(d_d2, nq * k, INFINITY);
    ws->launches += 2;
    const bool prof = ctx->profile_on;
    if (prof)
    {
        if (int r = ws->reserve_prof_events(1)) return r;
        B2_CUDA_TRY(cudaEventRecord(ws->prof_ev[0], s));
    }
    const NnWriter w = {d_idx, d_d2, k, 1u};
    uint32_t*      d_dbg = nullptr;
    const size_t   dbg_items = (nq + kItem - 1) / kItem + 1;
    if (getenv("B200ICP_DBG_ITEMS"))
    {
        B2_CUDA_TRY(cudaMalloc(&d_dbg, 3 * dbg_items * sizeof(uint32_t)));
        B2_CUDA_TRY(cudaMemset(d_dbg, 0, 3 * dbg_items * sizeof(uint32_t)));
        B2_CUDA_TRY(cudaMemcpyToSymbol(g_dbg_item_cycles, &d_dbg, sizeof(d_dbg)));
#ifdef B200ICP_DBG_COUNT
        {
            uint32_t* d_cnt = nullptr;
            B2_CUDA_TRY(cudaMalloc(&d_cnt, (size_t)1 << 22));
            B2_CUDA_TRY(cudaMemset(d_cnt, 0, (size_t)1 << 22));
            B2_CUDA_TRY(cudaMemcpyToSymbol(g_dbg_lane_cand, &d_cnt, sizeof(d_cnt)));
        }
#endif
    }
    if (k == 1)
        launch_search_k<1>(ctx, ws, nq, 1, sj.d_clouds, sj.d_jobs, ctx->D, cap_d2, 0, w);
    else if (k <= 4)
        launch_search_k<4>(ctx, ws, nq, 1, sj.d_clouds, sj.d_jobs, ctx->D, cap_d2, 0, w);
    else if (k <= 6)
        launch_search_k<6>(ctx, ws, nq, 1, sj.d_clouds, sj.d_jobs, ctx->D, cap_d2, 0, w);
    else
        launch_search_k<8>(ctx, ws, nq, 1, sj.d_clouds, sj.d_jobs, ctx->D, cap_d2, 0, w);
    if (uncapped)
    {
        const int grid = ctx->sm_count * 8;
        if (k == 1)
            knn_complete_kernel<1><<<grid, 256, 0, s>>>(sj.d_clouds, sj.d_jobs, k, d_idx, d_d2);
        else if (k <= 4)
            knn_complete_kernel<4><<<grid, 256, 0, s>>>(sj.d_clouds, sj.d_jobs, k, d_idx, d_d2);
        else if (k <= 6)
            knn_complete_kernel<6><<<grid, 256, 0, s>>>(sj.d_clouds, sj.d_jobs, k, d_idx, d_d2);
        else
            knn_complete_kernel<8><<<grid, 256, 0, s>>>(sj.d_clouds, sj.d_jobs, k, d_idx, d_d2);
        ws->launches++;
    }
    if (prof) B2_CUDA_TRY(cudaEventRecord(ws->prof_ev[1], s));
    B2_CUDA_TRY(cudaGetLastError());
#ifdef B200ICP_DBG_PHASES
    {
        unsigned long long h[16], z[16] = {0};
        B2_CUDA_TRY(cudaStreamSynchronize(s));
        B2_CUDA_TRY(cudaMemcpyFromSymbol(h, g_dbg_phase, sizeof(h)));
        B2_CUDA_TRY(cudaMemcpyToSymbol(g_dbg_phase, z, sizeof(z)));
        fprintf(stderr, "[dbg phases] Mcycles: setup %.1f - %.1f - %.1f enum %.1f sweep %.1f alone %.1f epi %.1f | "
                        "loaded %llu kept %llu lanes searched singly %llu passes %llu groups %llu alone %llu\n",
                h[0] * 1e-6, h[1] * 1e-6, h[2] * 1e-6, h[3] * 1e-6, h[4] * 1e-6, h[5] * 1e-6, h[6] * 1e-6, h[8], h[9],
                h[10], h[11], h[12], h[13]);
    }
#endif
    if (d_dbg)
    {
        std::vector<uint32_t> h(3 * dbg_items);
        B2_CUDA_TRY(cudaMemcpy(h.data(), d_dbg, 3 * dbg_items * sizeof(uint32_t), cudaMemcpyDeviceToHost));
#ifdef B200ICP_DBG_COUNT
        {   // the 25 slowest items: cycles, candidates of the busiest lane, mean candidates per lane
            const size_t ni = (nq + kItem - 1) / kItem;
            std::vector<size_t> ord(ni);
            for (size_t i = 0; i < ni; i++) ord[i] = i;
            std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return (h[a] & 0x7FFFFFFFu) > (h[b] & 0x7FFFFFFFu); });
            double tot = 0, totmax = 0;
            for (size_t i = 0; i < ni; i++) tot += h[ni + 2 + 2 * i], totmax += 32.0 * h[ni + 1 + 2 * i];
            fprintf(stderr, "[dbg items] candidates: sum over lanes %.3g, 32 x max lane %.3g (lane efficiency %.2f)\n", tot, totmax, tot / totmax);
            for (size_t r = 0; r < 25 && r < ni; r++)
            {
                const size_t i = ord[r];
                fprintf(stderr, "[dbg items] #%zu item %zu cycles %u %s max-lane %u mean-lane %.0f\n", r, i, h[i] & 0x7FFFFFFFu,
                        (h[i] & 0x80000000u) ? "fallback" : "tiled", h[ni + 1 + 2 * i], h[ni + 2 + 2 * i] / 32.0);
            }
            h.resize(ni);
        }
#else
        h.resize(dbg_items);
#endif
        uint32_t* null_ptr = nullptr;
        B2_CUDA_TRY(cudaMemcpyToSymbol(g_dbg_item_cycles, &null_ptr, sizeof(null_ptr)));
        cudaFree(d_dbg);
        std::vector<uint32_t> tiled, fb;
        for (uint32_t v : h)
            if (v) ((v & 0x80000000u) ? fb : tiled).push_back(v & 0x7FFFFFFFu);
        auto stats = [](std::vector<uint32_t>& v, const char* name) {
            if (v.empty()) { fprintf(stderr, "[dbg items] %s: none\n", name); return; }
            std::sort(v.begin(), v.end());
            double sum = 0;
            for (uint32_t x : v) sum += x;
            fprintf(stderr, "[dbg items] %s: n=%zu cycles mean=%.0f p50=%u p90=%u p99=%u max=%u sum=%.3g\n", name,
                    v.size(), sum / v.size(), v[v.size() / 2], v[v.size() * 9 / 10], v[v.size() * 99 / 100],
                    v.back(), sum);
        };
        stats(tiled, "tiled");
        stats(fb, "fallback");
    }
    if (idx_out)
        B2_CUDA_TRY(cudaMemcpyAsync(idx_out, d_idx, nq * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    if (d2_out)
        B2_CUDA_TRY(cudaMemcpyAsync(d2_out, d_d2, nq * k * sizeof(float), cudaMemcpyDeviceToHost, s));
    B2_CUDA_TRY(cudaStreamSynchronize(s));
    if (prof)
    {
        float ms = 0;
        B2_CUDA_TRY(cudaEventElapsedTime(&ms, ws->prof_ev[0], ws->prof_ev[1]));
        std::lock_guard<std::mutex> lk(ctx->mtx);
        ctx->prof.knn_launches++;
        ctx->prof.knn_ms += ms;
        ctx->prof.knn_queries += nq;
    }
    return B200ICP_OK;
}

// Owner side of the reduce-scatter exchange: thread ql merges the `world` lists
// of local query ql (global q = rank * per + ql) and stores the merged row into
// the result region of EVERY rank (P2P stores over NVLink).
struct PeerPtrs
{
    uint64_t* p[8];
};
template <int K>
__global__ void __launch_bounds__(256)
    merge_scatter_kernel(const uint64_t* __restrict__ gather, uint32_t nparts, uint32_t world, uint32_t rank,
                         uint32_t per, uint32_t nq, uint32_t k, PeerPtrs result)
{
    const uint32_t ql = blockIdx.x * blockDim.x + threadIdx.x;
    if (ql >= per) return;
    const size_t q = (size_t)rank * per + ql;
    if (q >= nq) return;
    uint64_t key[K];
#pragma unroll
    for (int i = 0; i < K; i++) key[i] = ~0ull;
    for (uint32_t p = 0; p < nparts; p++)
    {
        const uint64_t* row = gather + ((size_t)p * per + ql) * k;
        for (uint32_t i = 0; i < k; i++)
        {
            const uint64_t kk = row[i];  // written by peers: no read-only path
            if (!(kk < key[K - 1])) break;
            topk_insert<K>(key, kk);
        }
    }
#pragma unroll
    for (int i = 0; i < K; i++)
        if (key[i] == ~0ull) key[i] = B200ICP_NO_KEY;
    for (uint32_t r = 0; r < world; r++)
    {
        uint64_t* dst = result.p[r] + q * k;
        if ((k & 1u) == 0 && (K & 1) == 0)
        {
#pragma unroll
            for (int i = 0; i < K; i += 2)
                if ((uint32_t)i < k) *reinterpret_cast<ulonglong2*>(dst + i) = make_ulonglong2(key[i], key[i + 1]);
        }
        else
        {
#pragma unroll
            for (int i = 0; i < K; i++)
                if ((uint32_t)i < k) dst[i] = key[i];
        }
    }
}

int run_knn_keys(::b200icp* ctx, const b200icp_cloud* ref, const b200icp_cloud* q, const double* pose6,
                 uint32_t k, float max_dist, const uint32_t* d_index_map, uint64_t* d_keys_out)
{
    if (k < 1 || k > B200ICP_MAX_KNN || !(max_dist > 0) || !std::isfinite(max_dist))
    {
        set_error("k=%u outside [1,%d] or max_dist not positive and finite", k, B200ICP_MAX_KNN);
        return B200ICP_ERR_BAD_ARG;
    }
    Lease L(ctx);
    if (!L.ws) return B200ICP_ERR_CUDA;
    Workspace*   ws = L.ws;
    cudaStream_t s = ws->stream;
    const size_t nq = q->n;
    if (nq == 0) return B200ICP_OK;
    SingleJob sj;
    if (int r = single_job_setup(ws, ref, q, pose6, 0, sj, [&](Carver&) {})) return r;
    const float cap_d2 = max_dist * max_dist;
    fill_u64_kernel<<<(int)((nq * k + 255) / 256), 256, 0, s>>>(d_keys_out, nq * k, B200ICP_NO_KEY);
    ws->launches++;
    const bool prof = ctx->profile_on;
    if (prof)
    {
        if (int r = ws->reserve_prof_events(1)) return r;
        B2_CUDA_TRY(cudaEventRecord(ws->prof_ev[0], s));
    }
    const KeyWriter w = {d_keys_out, d_index_map, k};
    if (k == 1)
        launch_search_k<1>(ctx, ws, nq, 1, sj.d_clouds, sj.d_jobs, ctx->D, cap_d2, 0, w);
    else if (k <= 4)
        launch_search_k<4>(ctx, ws, nq, 1, sj.d_clouds, sj.d_jobs, ctx->D, cap_d2, 0, w);
    else if (k <= 6)
        launch_search_k<6>(ctx, ws, nq, 1, sj.d_clouds, sj.d_jobs, ctx->D, cap_d2, 0, w);
    else
        launch_search_k<8>(ctx, ws, nq, 1, sj.d_clouds, sj.d_jobs, ctx->D, cap_d2, 0, w);
    if (prof) B2_CUDA_TRY(cudaEventRecord(ws->prof_ev[1], s));
    B2_CUDA_TRY(cudaGetLastError());
    B2_CUDA_TRY(cudaStreamSynchronize(s));
    if (prof)
    {
        float ms = 0;
        B2_CUDA_TRY(cudaEventElapsedTime(&ms, ws->prof_ev[0], ws->prof_ev[1]));
        std::lock_guard<std::mutex> lk(ctx->mtx);
        ctx->prof.knn_launches++;
        ctx->prof.knn_ms += ms;
        ctx->prof.knn_queries += nq;
    }
    return B200ICP_OK;
}

// enqueues the scatter search on ws->stream (no host synchronisation)
static int enqueue_scatter(::b200icp* ctx, Workspace* ws, const b200icp_cloud* ref, const b200icp_cloud* q,
                           const double* pose6, uint32_t k, float max_dist, const uint32_t* d_index_map,
                           uint64_t* const* d_gather, uint32_t world, uint32_t rank, int atomic_min, uint32_t per,
                           bool time_it)
{
    if (k < 1 || k > B200ICP_MAX_KNN || !(max_dist > 0) || !std::isfinite(max_dist) || (atomic_min && k != 1))
    {
        set_error("k=%u outside [1,%d], max_dist not positive and finite, or atomic_min with k != 1", k,
                  B200ICP_MAX_KNN);
        return B200ICP_ERR_BAD_ARG;
    }
    cudaStream_t s = ws->stream;
    const size_t nq = q->n;
    SingleJob    sj;
    if (int r = single_job_setup(ws, ref, q, pose6, 0, sj, [&](Carver&) {})) return r;
    const float cap_d2 = max_dist * max_dist;
    KeyScatter  w;
    memset(&w, 0, sizeof(w));
    for (uint32_t r = 0; r < world; r++) w.peer[r] = d_gather[r];
    w.map = d_index_map, w.k = k, w.world = world, w.rank = rank, w.nq = (uint32_t)nq, w.amin = atomic_min ? 1u : 0u;
    w.per = per;
    if (!atomic_min)
    {
        // rows of non-finite queries are never visited by the search: this rank's
        // slice of EVERY buffer is reset first
        const size_t rows = per ? per : nq;
        for (uint32_t r = 0; r < world; r++)
            fill_u64_kernel<<<(int)((rows * k + 255) / 256), 256, 0, s>>>(d_gather[r] + (size_t)rank * rows * k,
                                                                            rows * k, B200ICP_NO_KEY);
        ws->launches += world;
    }
    if (time_it)
    {
        if (int r = ws->reserve_prof_events(1)) return r;
        B2_CUDA_TRY(cudaEventRecord(ws->prof_ev[0], s));
    }
    if (k == 1)
        launch_search_k<1>(ctx, ws, nq, 1, sj.d_clouds, sj.d_jobs, ctx->D, cap_d2, 0, w);
    else if (k <= 4)
        launch_search_k<4>(ctx, ws, nq, 1, sj.d_clouds, sj.d_jobs, ctx->D, cap_d2, 0, w);
    else if (k <= 6)
        launch_search_k<6>(ctx, ws, nq, 1, sj.d_clouds, sj.d_jobs, ctx->D, cap_d2, 0, w);
    else
        launch_search_k<8>(ctx, ws, nq, 1, sj.d_clouds, sj.d_jobs, ctx->D, cap_d2, 0, w);
    if (time_it) B2_CUDA_TRY(cudaEventRecord(ws->prof_ev[1], s));
    B2_CUDA_TRY(cudaGetLastError());
    return B200ICP_OK;
}

static int account_knn_time(::b200icp* ctx, Workspace* ws, size_t nq)
{
    float ms = 0;
    B2_CUDA_TRY(cudaEventElapsedTime(&ms, ws->prof_ev[0], ws->prof_ev[1]));
    std::lock_guard<std::mutex> lk(ctx->mtx);
    ctx->prof.knn_launches++;
    ctx->prof.knn_ms += ms;
    ctx->prof.knn_queries += nq;
    return B200ICP_OK;
}

int run_knn_keys_scatter(::b200icp* ctx, const b200icp_cloud* ref, const b200icp_cloud* q, const double* pose6,
                         uint32_t k, float max_dist, const uint32_t* d_index_map, uint64_t* const* d_gather,
                         uint32_t world, uint32_t rank, int atomic_min)
{
    if (q->n == 0) return B200ICP_OK;
    Lease L(ctx);
    if (!L.ws) return B200ICP_ERR_CUDA;
    const bool prof = ctx->profile_on;
    if (int r = enqueue_scatter(ctx, L.ws, ref, q, pose6, k, max_dist, d_index_map, d_gather, world, rank,
                                atomic_min, 0u, prof))
        return r;
    B2_CUDA_TRY(cudaStreamSynchronize(L.ws->stream));
    if (prof) return account_knn_time(ctx, L.ws, q->n);
    return B200ICP_OK;
}

// Barrier of the ranks of one node over peer memory (see b200icp_peer_barrier).
struct PeerFlags
{
    uint64_t* p[8];
};
__global__ void peer_barrier_kernel(PeerFlags flags, uint32_t world, uint32_t rank, unsigned long long epoch,
                                    uint32_t* __restrict__ timed_out)
{
    const uint32_t r = threadIdx.x;
    if (r >= world) return;
    __threadfence_system();  // everything this stream wrote before is visible system-wide first
    volatile unsigned long long* mine = reinterpret_cast<volatile unsigned long long*>(flags.p[rank]);
    volatile unsigned long long* theirs = reinterpret_cast<volatile unsigned long long*>(flags.p[r]);
    theirs[rank] = epoch;
    __threadfence_system();
    const long long t0 = clock64();
    while (mine[r] < epoch)
    {
        if (clock64() - t0 > 4000000000ll)
        {  // ~2 s: a peer is missing; give up instead of hanging the GPU
            *timed_out = 1u;
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

int run_peer_barrier(::b200icp* ctx, uint64_t* const* d_flags, uint32_t world, uint32_t rank, uint64_t epoch)
{
    Lease L(ctx);
    if (!L.ws) return B200ICP_ERR_CUDA;
    Workspace*   ws = L.ws;
    cudaStream_t s = ws->stream;
    uint32_t* d_to = ws->d_flag;
    uint32_t* h_to = ws->h_flag;
    B2_CUDA_TRY(cudaMemsetAsync(d_to, 0, sizeof(uint32_t), s));
    PeerFlags f;
    memset(&f, 0, sizeof(f));
    for (uint32_t r = 0; r < world; r++) f.p[r] = d_flags[r];
    peer_barrier_kernel<<<1, 32, 0, s>>>(f, world, rank, (unsigned long long)epoch, d_to);
    ws->launches++;
    B2_CUDA_TRY(cudaMemcpyAsync(h_to, d_to, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    B2_CUDA_TRY(cudaStreamSynchronize(s));
    B2_CUDA_TRY(cudaGetLastError());
    if (*h_to)
    {
        set_error("peer barrier %llu timed out on rank %u: a rank of the node did not arrive",
                  (unsigned long long)epoch, rank);
        return B200ICP_ERR_CUDA;
    }
    return B200ICP_OK;
}

// One call, one stream, one host synchronisation (b200icp_knn_keys_exchange):
// reset -> barrier -> search with scatter into every rank's buffer -> barrier ->
// merge into d_out.  d_bases[r]: rank r's exchange buffer, flags at the base,
// keys kPeerHeader bytes further.
constexpr size_t kPeerHeader = 256;
int run_knn_exchange(::b200icp* ctx, const b200icp_cloud* ref, const b200icp_cloud* q, const double* pose6,
                     uint32_t k, float max_dist, const uint32_t* d_index_map, uint64_t* const* d_bases,
                     uint32_t world, uint32_t rank, uint64_t* epoch_io, uint64_t* d_out)
{
    const size_t nq = q->n;
    if (nq == 0) return B200ICP_OK;
    if (k < 1 || k > B200ICP_MAX_KNN)
    {
        set_error("k=%u outside [1,%d]", k, B200ICP_MAX_KNN);
        return B200ICP_ERR_BAD_ARG;
    }
    Lease L(ctx);
    if (!L.ws) return B200ICP_ERR_CUDA;
    Workspace*   ws = L.ws;
    cudaStream_t s = ws->stream;
    uint64_t*    data[8];
    PeerFlags    f;
    memset(&f, 0, sizeof(f));
    for (uint32_t r = 0; r < world; r++)
    {
        f.p[r] = d_bases[r];
        data[r] = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(d_bases[r]) + kPeerHeader);
    }
    const bool     amin = (k == 1);
    const bool     prof = ctx->profile_on;
    const uint32_t per = (uint32_t)((nq + world - 1) / world);  // queries owned by one rank
    uint64_t       epoch = *epoch_io;
    auto barrier = [&]() {
        peer_barrier_kernel<<<1, 32, 0, s>>>(f, world, rank, (unsigned long long)++epoch, ws->d_flag);
        ws->launches++;
    };
    B2_CUDA_TRY(cudaMemsetAsync(ws->d_flag, 0, sizeof(uint32_t), s));
    // Buffer: gather region [world][per][k] (k = 1: [per], one folded slot per owned query), then the result
    // region [world*per][k].
    PeerPtrs res;
    memset(&res, 0, sizeof(res));
    for (uint32_t r = 0; r < world; r++) res.p[r] = data[r] + (size_t)world * per * k;
    const int mblocks = (int)((per + 255) / 256);
    if (amin)
    {
        // k = 1: every rank folds its key into the OWNER's slot of the query (atomicMin over NVLink: search +
        // reduce-scatter(MIN) in one kernel); the owner then stores its slice into every rank's result region
        fill_u64_kernel<<<mblocks, 256, 0, s>>>(data[rank], per, B200ICP_NO_KEY);
        ws->launches++;
        barrier();
        if (int r = enqueue_scatter(ctx, ws, ref, q, pose6, k, max_dist, d_index_map, data, world, rank, 1, per, prof))
            return r;
        barrier();
        merge_scatter_kernel<1><<<mblocks, 256, 0, s>>>(data[rank], 1u, world, rank, per, (uint32_t)nq, k, res);
        ws->launches++;
        barrier();
    }
    else
    {
        // k > 1, reduce-scatter form: 1. rows of query q go to owner(q) only; 2. the owner merges its slice and
        // stores the merged rows into every rank's result region; 3. the local result region is the answer.
        barrier();
        if (int r = enqueue_scatter(ctx, ws, ref, q, pose6, k, max_dist, d_index_map, data, world, rank, 0, per, prof))
            return r;
        barrier();
        if (k <= 4)
            merge_scatter_kernel<4><<<mblocks, 256, 0, s>>>(data[rank], world, world, rank, per, (uint32_t)nq, k, res);
        else if (k <= 6)
            merge_scatter_kernel<6><<<mblocks, 256, 0, s>>>(data[rank], world, world, rank, per, (uint32_t)nq, k, res);
        else
            merge_scatter_kernel<8><<<mblocks, 256, 0, s>>>(data[rank], world, world, rank, per, (uint32_t)nq, k, res);
        ws->launches++;
        barrier();
    }
    B2_CUDA_TRY(cudaMemcpyAsync(d_out, res.p[rank], nq * k * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
    B2_CUDA_TRY(cudaMemcpyAsync(ws->h_flag, ws->d_flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    B2_CUDA_TRY(cudaStreamSynchronize(s));
    B2_CUDA_TRY(cudaGetLastError());
    *epoch_io = epoch;
    if (*ws->h_flag)
    {
        set_error("peer barrier timed out on rank %u: a rank of the node did not arrive", rank);
        return B200ICP_ERR_CUDA;
    }
    if (prof) return account_knn_time(ctx, ws, nq);
    return B200ICP_OK;
}

int run_fill_no_key(::b200icp* ctx, uint64_t* d_keys, size_t n)
{
    if (n == 0) return B200ICP_OK;
    Lease L(ctx);
    if (!L.ws) return B200ICP_ERR_CUDA;
    cudaStream_t s = L.ws->stream;
    fill_u64_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(d_keys, n, B200ICP_NO_KEY);
    L.ws->launches++;
    B2_CUDA_TRY(cudaGetLastError());
    B2_CUDA_TRY(cudaStreamSynchronize(s));
    return B200ICP_OK;
}

int run_merge_keys(::b200icp* ctx, const uint64_t* d_parts, uint32_t parts, size_t part_stride, size_t nq,
                   uint32_t k, uint64_t* d_out)
{
    if (k < 1 || k > B200ICP_MAX_KNN)
    {
        set_error("k=%u outside [1,%d]", k, B200ICP_MAX_KNN);
        return B200ICP_ERR_BAD_ARG;
    }
    if (nq == 0) return B200ICP_OK;
    Lease L(ctx);
    if (!L.ws) return B200ICP_ERR_CUDA;
    cudaStream_t s = L.ws->stream;
    const int    blocks = (int)((nq + 255) / 256);
    if (k == 1)
        merge_keys_kernel<1><<<blocks, 256, 0, s>>>(d_parts, parts, part_stride, nq, k, d_out);
    else if (k <= 4)
        merge_keys_kernel<4><<<blocks, 256, 0, s>>>(d_parts, parts, part_stride, nq, k, d_out);
    else if (k <= 6)
        merge_keys_kernel<6><<<blocks, 256, 0, s>>>(d_parts, parts, part_stride, nq, k, d_out);
    else
        merge_keys_kernel<8><<<blocks, 256, 0, s>>>(d_parts, parts, part_stride, nq, k, d_out);
    L.ws->launches++;
    B2_CUDA_TRY(cudaGetLastError());
    B2_CUDA_TRY(cudaStreamSynchronize(s));
    return B200ICP_OK;
}

int run_match(::b200icp* ctx, const b200icp_cloud* from, const b200icp_cloud* to,
              const double* pose6, uint8_t* paired, uint32_t* nn_idx, uint32_t* nn_cnt,
              double* centroid, double* normal, uint32_t* n_pairings)
{
    if (int r = check_supported(ctx->D)) return r;
    Lease L(ctx);
    if (!L.ws) return B200ICP_ERR_CUDA;
    Workspace*          ws = L.ws;
    cudaStream_t        s = ws->stream;
    const IcpDevParams& D = ctx->D;
    const size_t        n = to->n, k = (D.matcher_kind == B200ICP_MATCHER_POINTS_DISTANCE) ? 1 : D.knn;
    if (n_pairings) *n_pairings = 0;
    if (n == 0) return B200ICP_OK;
    const uint32_t chunk_items = fit_chunk_items(n, 1);
    const uint32_t G = fit_ctas_per_job(ctx, n, 1, chunk_items);
    const size_t   chunks = ((n + kItem - 1) / kItem + chunk_items - 1) / chunk_items;
    const size_t   groups = (chunks + kFitGroup - 1) / kFitGroup;
    FitBuffers     fb = {nullptr, nullptr, nullptr};
    uint32_t*      d_nn = nullptr;
    MatchOut       mo;
    SingleJob      sj;
    // the matcher is active at iteration run_from_iteration
    if (int r = single_job_setup(ws, from, to, pose6, D.run_from_iteration, sj, [&](Carver& c) {
            fb.tickets = c.take<uint32_t>(groups ? groups : 1);
            fb.gpartials = c.take<double>((groups ? groups : 1) * (size_t)kNumMoments);
            fb.partials = c.take<double>((chunks ? chunks : 1) * (size_t)kNumMoments);
            d_nn = c.take<uint32_t>(n * (size_t)matcher_k(D));
            mo.paired = c.take<uint8_t>(n);
            mo.nn_idx = c.take<uint32_t>(n * k);
            mo.nn_cnt = c.take<uint32_t>(n);
            mo.centroid = c.take<double>(n * 3);
            mo.normal = c.take<double>(n * 3);
        }))
        return r;
    B2_CUDA_TRY(cudaMemsetAsync(fb.tickets, 0, (groups ? groups : 1) * sizeof(uint32_t), s));
    B2_CUDA_TRY(cudaMemsetAsync(mo.paired, 0, n, s));
    B2_CUDA_TRY(cudaMemsetAsync(mo.nn_cnt, 0, n * sizeof(uint32_t), s));
    B2_CUDA_TRY(cudaMemsetAsync(mo.nn_idx, 0xFF, n * k * sizeof(uint32_t), s));
    B2_CUDA_TRY(cudaMemsetAsync(mo.centroid, 0, n * 3 * sizeof(double), s));
    B2_CUDA_TRY(cudaMemsetAsync(mo.normal, 0, n * 3 * sizeof(double), s));
    launch_match_search(ctx, ws, D, n, 1, sj.d_clouds, sj.d_jobs, d_nn);
    launch_fit<true>(ws, D, dim3(G, 1), sj.d_clouds, sj.d_jobs, d_nn, fb, chunk_items, 0, mo, nullptr, nullptr);
    B2_CUDA_TRY(cudaGetLastError());
    JobDev* hj = (JobDev*)((char*)ws->h_pinned + align_up(2 * sizeof(CloudView)));  // the reduced moments come back
    B2_CUDA_TRY(cudaMemcpyAsync(hj, sj.d_jobs, sizeof(JobDev), cudaMemcpyDeviceToHost, s));
    if (paired) B2_CUDA_TRY(cudaMemcpyAsync(paired, mo.paired, n, cudaMemcpyDeviceToHost, s));
    if (nn_idx)
        B2_CUDA_TRY(cudaMemcpyAsync(nn_idx, mo.nn_idx, n * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    if (nn_cnt)
        B2_CUDA_TRY(cudaMemcpyAsync(nn_cnt, mo.nn_cnt, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    if (centroid)
        B2_CUDA_TRY(cudaMemcpyAsync(centroid, mo.centroid, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (normal)
        B2_CUDA_TRY(cudaMemcpyAsync(normal, mo.normal, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    B2_CUDA_TRY(cudaStreamSynchronize(s));
    if (n_pairings)
        *n_pairings = (uint32_t)(pairing_count(hj->M, D.matcher_kind == B200ICP_MATCHER_POINTS_DISTANCE) + 0.5);
    return B200ICP_OK;
}

#include "sharded.inl"

}  // namespace b2
