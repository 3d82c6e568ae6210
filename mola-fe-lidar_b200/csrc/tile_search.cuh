// tile_search.cuh -- warp-autonomous exact k-nearest-neighbour search.
//
// Work unit ("item") = kItem (32) consecutive points of the LOCAL cloud's own
// cell-sorted array (built once with the cloud's index, cloud.cu).  ONE WARP owns an item: under the current pose its
// points land in a small box of the GLOBAL cloud's grid.  The warp
//   1. reduces the box of fine cells its queries fall in and grows it by the
//      search radius S (in cells) -> the TILE,
//   2. probes the global cloud's block hash ONCE per block of the tile (lanes
//      in parallel) instead of one chain of dependent probes per query and
//      shell, and keeps the block records in shared memory,
//   3. copies each present block's run of fine-cell offsets next to them and
//      folds the 64-bit occupancy masks into one bit-row per (y,z) of the tile
//      plus one row-occupancy word per z layer,
// and then every lane walks the shells around its own home cell using shared
// memory only: an empty layer costs one AND, an empty row one more, a cell
// one popc.  Candidate coordinates stay in global memory (L1/L2 resident
// float4; eight independent loads in flight on runs of >= 8 points, else four).  There is no CTA-wide barrier:
// the warps of a CTA work on different items.
//
// Results are identical to knn_search<K> (knn_search.cuh), which remains the
// fallback for items whose tile would not fit (very spread queries, exotic
// radius / cell ratios).
#pragma once
#include "knn_search.cuh"

namespace b2
{
constexpr int kTileBlocks = 192;   // blocks probed per tile
constexpr int kTileRows = 512;     // (y,z) rows per tile
constexpr int kTileMaxDim = 32;    // cells per axis (one bit-row word)
constexpr int kTileFs = 640;       // cached fine-cell offsets per tile
constexpr uint32_t kFsNone = 0xFFFFFFFFu;

struct WarpTile
{
    uint4    rec[kTileBlocks];     // BlockRec of every probed block (mask 0 when absent)
    uint32_t fsoff[kTileBlocks];   // where the block's fine_start run sits in fs[] (kFsNone: read global)
    uint32_t fs[kTileFs];          // fine_start[fine_base .. fine_base + popc(mask)] per cached block
    uint32_t rowmask[kTileRows];   // occupancy along x of row z * ny + y
    uint32_t ymask[kTileMaxDim];   // per z layer: which rows y hold any point
};

// warp-uniform geometry of a built tile
struct TileGeom
{
    int t0x, t0y, t0z;  // origin in fine-cell coordinates (may be negative)
    int nx, ny, nz;     // extent in cells
    int bx0, by0, bz0;  // first block per axis
    int nbx, nby;       // blocks per axis (x, y)
};

// A query's position in the global grid: home fine cell (absolute), fractional
// position inside it, distance to the nearest face. Same arithmetic as the
// head of knn_search<K>.
struct QueryCell
{
    int   hx, hy, hz;
    float fx, fy, fz;
    float gmin;
    bool  valid;  // false: NaN position -> no neighbours
};

__device__ __forceinline__ int search_shells(const GridDev& g, float cap_d2)
{
    return max(1, (int)ceilf(sqrtf(cap_d2) * g.inv_cell * 1.0005f));
}

__device__ __forceinline__ QueryCell locate_query(const GridDev& g, int S, float qx, float qy, float qz)
{
    QueryCell   q;
    const float inv = g.inv_cell;
    const float lim_lo = -(float)(S + 2), lim_hi = (float)(kFineMax + S + 3);
    float       ux = (qx - g.ox) * inv, uy = (qy - g.oy) * inv, uz = (qz - g.oz) * inv;
    q.valid = (ux == ux) && (uy == uy) && (uz == uz);
    ux = fminf(fmaxf(ux, lim_lo), lim_hi);
    uy = fminf(fmaxf(uy, lim_lo), lim_hi);
    uz = fminf(fmaxf(uz, lim_lo), lim_hi);
    const float flx = floorf(ux), fly = floorf(uy), flz = floorf(uz);
    q.hx = (int)flx, q.hy = (int)fly, q.hz = (int)flz;
    q.fx = ux - flx, q.fy = uy - fly, q.fz = uz - flz;
    q.gmin = fmaxf(fminf(fminf(fminf(q.fx, 1.0f - q.fx), fminf(q.fy, 1.0f - q.fy)),
                         fminf(q.fz, 1.0f - q.fz)) - g.slack, 0.0f);
    if (!q.valid) q.hx = q.hy = q.hz = 0;
    return q;
}

// Warp-collective. lo/hi: box of the home cells of ALL the item's valid
// queries (already reduced over the warp; lo > hi when there is none).
// Returns true when the tile is usable (uniform over the warp).
__device__ __forceinline__ bool warp_tile_build(WarpTile& W, TileGeom& G, const CloudView& cv,
                                                const int (&lo)[3], const int (&hi)[3], int S)
{
    const int lane = threadIdx.x & 31;
    if (lo[0] > hi[0]) return false;  // nobody searches
    G.t0x = lo[0] - S, G.t0y = lo[1] - S, G.t0z = lo[2] - S;
    G.nx = hi[0] - lo[0] + 1 + 2 * S, G.ny = hi[1] - lo[1] + 1 + 2 * S, G.nz = hi[2] - lo[2] + 1 + 2 * S;
    if (G.nx > kTileMaxDim || G.ny > kTileMaxDim || G.nz > kTileMaxDim || G.ny * G.nz > kTileRows)
        return false;
    G.bx0 = G.t0x >> 2, G.by0 = G.t0y >> 2, G.bz0 = G.t0z >> 2;
    G.nbx = ((G.t0x + G.nx - 1) >> 2) - G.bx0 + 1;
    G.nby = ((G.t0y + G.ny - 1) >> 2) - G.by0 + 1;
    const int nbz = ((G.t0z + G.nz - 1) >> 2) - G.bz0 + 1;
    const int nb = G.nbx * G.nby * nbz;
    if (nb > kTileBlocks) return false;
    __syncwarp();  // the previous item's searches are over
    const int nrow = G.ny * G.nz;
    for (int i = lane; i < nrow; i += 32) W.rowmask[i] = 0u;
    W.ymask[lane] = 0u;
    __syncwarp();
    const uint32_t xmask = (G.nx >= 32) ? 0xFFFFFFFFu : ((1u << G.nx) - 1u);
    uint32_t       fs_used = 0;
    for (int i0 = 0; i0 < nb; i0 += 32)
    {
        const int i = i0 + lane;
        uint4     rec = make_uint4(0u, 0u, 0u, 0u);
        int       bxi = 0, byi = 0, bzi = 0;
        if (i < nb)
        {
            bxi = i % G.nbx, byi = (i / G.nbx) % G.nby, bzi = i / (G.nbx * G.nby);
            const int bx = G.bx0 + bxi, by = G.by0 + byi, bz = G.bz0 + bzi;
            if (bx >= 0 && by >= 0 && bz >= 0 && bx <= kGridMax && by <= kGridMax && bz <= kGridMax)
            {
                const uint32_t bkey =
                    (uint32_t)bx | ((uint32_t)by << kGridBits) | ((uint32_t)bz << (2 * kGridBits));
                uint4 r;
                if (block_lookup(cv, bkey, r)) rec = r;
            }
        }
        const uint64_t occ = ((uint64_t)rec.w << 32) | (uint64_t)rec.z;
        // where this block's run of fine-cell offsets goes in fs[]
        const uint32_t need = occ ? (uint32_t)__popcll(occ) + 1u : 0u;
        uint32_t       incl = need;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += v;
        }
        const uint32_t off = fs_used + incl - need;
        const bool     cached = need && (off + need <= (uint32_t)kTileFs);
        fs_used += __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (i < nb)
        {
            W.rec[i] = rec;
            W.fsoff[i] = cached ? off : kFsNone;
        }
        if (occ)
        {
            if (cached)
                for (uint32_t e = 0; e < need; e++) W.fs[off + e] = __ldg(cv.fine_start + rec.y + e);
            // fold the 4x4x4 mask into the tile's bit rows
            const int cbx = 4 * (G.bx0 + bxi) - G.t0x, cby = 4 * (G.by0 + byi) - G.t0y,
                      cbz = 4 * (G.bz0 + bzi) - G.t0z;
#pragma unroll 4
            for (int r = 0; r < 16; r++)
            {
                const uint32_t nib = (uint32_t)(occ >> (4 * r)) & 0xFu;
                if (nib == 0) continue;
                const int y = cby + (r & 3), z = cbz + (r >> 2);
                if (y < 0 || z < 0 || y >= G.ny || z >= G.nz) continue;
                const uint32_t bits = ((cbx >= 0) ? (nib << cbx) : (nib >> (-cbx))) & xmask;
                if (bits == 0) continue;
                atomicOr(&W.rowmask[z * G.ny + y], bits);
                atomicOr(&W.ymask[z], 1u << y);
            }
        }
    }
    __syncwarp();
    return true;
}

// [beg, end) of the points of tile cell (x, y, z); the cell must be occupied
__device__ __forceinline__ uint2 tile_cell_range(const WarpTile& W, const TileGeom& G,
                                                 const CloudView& cv, int x, int y, int z)
{
    const int      ax = G.t0x + x, ay = G.t0y + y, az = G.t0z + z;
    const int      b = ((ax >> 2) - G.bx0) + G.nbx * (((ay >> 2) - G.by0) + G.nby * ((az >> 2) - G.bz0));
    const uint4    rec = W.rec[b];
    const uint64_t occ = ((uint64_t)rec.w << 32) | (uint64_t)rec.z;
    const int      bit = (ax & 3) | ((ay & 3) << 2) | ((az & 3) << 4);
    const uint32_t k = (uint32_t)__popcll(occ & ((1ull << bit) - 1ull));
    const uint32_t so = W.fsoff[b];
    if (so != kFsNone) return make_uint2(W.fs[so + k], W.fs[so + k + 1]);
    return make_uint2(__ldg(cv.fine_start + rec.y + k), __ldg(cv.fine_start + rec.y + k + 1));
}

// Per-lane search inside a built tile. keys must hold sentinel_key(cap_d2).
template <int K>
__device__ __forceinline__ void tile_knn(const WarpTile& W, const TileGeom& G, const CloudView& cv,
                                         const GridDev& g, const QueryCell& q, int S, float qx,
                                         float qy, float qz, uint64_t (&key)[K])
{
    const int   ny = G.ny;
    const int   hx = q.hx - G.t0x, hy = q.hy - G.t0y, hz = q.hz - G.t0z;
    const float slack = g.slack;
    const float to_cells2 = g.inv_cell * g.inv_cell * 1.0001f;
    float       worst = key_d2(key[K - 1]) * to_cells2;
    for (int s = 0; s <= S; s++)
    {
        if (s >= 1)
        {  // every cell of shell s is at least (s - 1 + gmin) cells away
            const float b = (float)(s - 1) + q.gmin;
            if (b * b > worst) break;
        }
        const uint32_t xfull = (2u << (hx + s)) - (1u << (hx - s));
        const uint32_t xends = (1u << (hx - s)) | (1u << (hx + s));
        const uint32_t yfull = (2u << (hy + s)) - (1u << (hy - s));
        for (int dz = -s; dz <= s; dz++)
        {
            // rows of this layer that hold anything inside the shell's y range
            uint32_t ym = W.ymask[hz + dz] & yfull;
            if (ym == 0) continue;
            const float gz = axis_gap(q.fz, dz, slack);
            const float gz2 = gz * gz;
            if (gz2 > worst) continue;
            const bool zface = (dz == -s) || (dz == s);
            const int  rowz = (hz + dz) * ny;
            while (ym)
            {
                const int y = __ffs((int)ym) - 1;
                ym &= ym - 1;
                const int  dy = y - hy;
                const bool face = zface || (dy == -s) || (dy == s);
                uint32_t   m = W.rowmask[rowz + y] & (face ? xfull : xends);
                if (m == 0) continue;
                const float gy = axis_gap(q.fy, dy, slack);
                const float gyz = gz2 + gy * gy;
                if (gyz > worst) continue;
                while (m)
                {
                    const int x = __ffs((int)m) - 1;
                    m &= m - 1;
                    const float gx = axis_gap(q.fx, x - hx, slack);
                    if (gyz + gx * gx > worst) continue;
                    const uint2 c = tile_cell_range(W, G, cv, x, y, hz + dz);
                    scan_range<K>(cv.pts, cv.gbox, c.x, c.y, qx, qy, qz, key);
                    worst = key_d2(key[K - 1]) * to_cells2;
                }
            }
        }
    }
}

// global warp id / warp count of a (G, jobs) launch of kChunk-thread CTAs
__device__ __forceinline__ uint32_t item_warp_id() { return blockIdx.x * (kChunk / 32) + (threadIdx.x >> 5); }
__device__ __forceinline__ uint32_t item_warp_count() { return gridDim.x * (kChunk / 32); }

}  // namespace b2
