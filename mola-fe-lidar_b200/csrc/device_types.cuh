// device_types.cuh -- HBM-resident data layout of the ICP path.
//
// A cloud (one point layer of the reference's mp2p_icp::metric_map_t, i.e. MRPT
// CPointsMap SoA float buffers -- SURVEY.md Appendix A.1) lives in HBM as:
//   x[n], y[n], z[n]      original order (what the caller uploaded)
//   pts[n_valid] float4   (x, y, z, bitcast original index), sorted by
//                         (30-bit Morton code of the point's BLOCK, 6-bit fine
//                         cell inside the block, 3-bit octant inside the fine
//                         cell) -- 8 consecutive points of a dense cell are
//                         spatially compact
//   gbox[2*ceil(n/8)]     float4 pairs: (min xyz, max xyz) of every GROUP of 8
//                         consecutive sorted points; the search tests a group's
//                         box against its current k-th best before loading it
//   rank[n]               original index -> sorted position
//   hkeys/hrecs/hrange    open-addressing hash: linear block key -> BlockRec
//                         {first point, first fine-cell ordinal, 64-bit
//                         occupancy mask of its 4x4x4 fine cells} and the
//                         block's point range [first, end) for the tile sweep
//   fine_start[n_fine+1]  first point of every occupied fine cell, in order
//   item_first[n_items+1] work items of the cloud when it is the LOCAL (query)
//                         side: runs of kItem consecutive sorted points:
//                         one warp, one round
//   GridDev               grid origin / cell size / counts, written on device
// Two levels: a BLOCK (edge >= the search radius, so 27 blocks always cover
// the radius) holds 4x4x4 FINE cells; dense regions are pruned at fine-cell
// granularity while empty space costs one hash probe per block.
// The grid replaces the lazily built nanoflann kd-tree (row I).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2
{
constexpr uint32_t kInvalid = 0xFFFFFFFFu;
constexpr uint32_t kEmptyKey = 0xFFFFFFFFu;
constexpr int kGridBits = 10;  // blocks per axis = 1024 (30-bit Morton / linear keys)
constexpr int kGridMax = (1 << kGridBits) - 1;
constexpr int kFineMax = 4 * (kGridMax + 1) - 1;  // fine cells per axis - 1
constexpr int kChunk = 128;    // threads per CTA in the search kernels (4 warps, 4 items in flight)
constexpr int kItem = 32;      // queries per work item = one warp round
constexpr int kNumMoments = 192; // three 8x8 tiles of the 16x16 moment matrix S (align.cu)
constexpr int kMicroBits = 3;   // octant of the point inside its fine cell: low bits of the sort key
constexpr int kGroup = 8;       // sorted points per bounding-box group
constexpr int kCycleMax = 6;    // longest period of a pose cycle the iteration loop recognises
constexpr int kSortBits = 30 + 6 + kMicroBits + 1;  // Morton30 | fine6 | octant3, +1 for the invalid key

struct GridDev
{
    float    ox, oy, oz;  // origin = bbox min of the finite points
    float    cell;        // FINE cell edge actually used (block edge = 4 * cell)
    float    inv_cell;
    float    slack;       // in cell units: guards float rounding in cell assignment
    uint32_t n_valid;     // finite points (sorted first)
    uint32_t n_cells;     // occupied fine cells
    uint32_t n_blocks;    // occupied blocks
    float    bmin[3], bmax[3];
    uint32_t n_items;     // work items (see item_first)
};

struct CloudView
{
    const float4*   pts;
    const float4*   gbox;        // [2 * groups]: (lo.xyz, -), (hi.xyz, -) of sorted points 8g .. 8g+7
    const uint32_t* rank;
    const GridDev*  grid;
    const uint32_t* hkeys;
    const uint4*    hrecs;       // BlockRec as (start, fine_base, mask.lo, mask.hi)
    const uint2*    hrange;      // [first, last+1) sorted positions of the block's points
    const uint32_t* fine_start;
    const uint32_t* item_first;
    uint32_t        hshift;  // 32 - log2(capacity)
    uint32_t        hmask;   // capacity - 1
    uint32_t        n;       // total points (incl. non-finite)
};

// One registration job, resident in HBM for the whole iteration loop.
struct JobDev
{
    double   R[9], t[3];          // current solution (to wrt from)
    double   Rprev[9], tprev[3];  // solution of the previous outer iteration
    double   M[kNumMoments];      // reduced moments of the last matcher run
    double   Mprev[kNumMoments];  // ... and of the one before (period-2 fast-forward, align.cu)
    double   cov[36];
    uint32_t iter;         // outer iterations completed
    uint32_t status;       // 0 running, 1 finished
    uint32_t term_reason;
    uint32_t n_pairings;
    uint32_t quality_count;
    uint32_t cov_singular;
    uint32_t from_cloud, to_cloud;  // indices into the launch's CloudView table
    uint32_t inner_iters_total;
    uint32_t pair_base;    // first row of this job in the launch's neighbour / pair buffers
    uint32_t next_item;    // work counter of the search stage (reset by the solver)
    uint32_t evaluated;    // quality + covariance taken (once, after the job finished)
    uint32_t rows_valid;   // the job's neighbour rows hold a matcher search of this registration (seeds)
    uint32_t rows_pose_valid;  // ... and rows_Rt is the pose that search ran at (quality certificate, align.cu)
    double   rows_Rt[12];
    uint32_t chunk_base;   // first chunk partial / first group partial + ticket of this job in the launch's
    uint32_t group_base;   //   fit buffers (align.cu, "chunk partials")
    uint32_t groups_done;  // arrival counter of the job's group reductions; zero between launches
    uint32_t npair_prev;   // n_pairings / inner iterations of the previous outer iteration
    uint32_t inner_prev;
    uint32_t cycle_hits;   // consecutive iterations whose pose repeated the pose of two iterations before
    uint32_t cycle_at;     // outer iteration at which a period-2 cycle was recognised (0: none)
    // longer cycles (period 3 .. kCycleMax, align.cu): the poses, pairing counts and inner iterations of the last
    // kCycleMax outer iterations (slot = iteration % kCycleMax) and, per period, how many consecutive iterations
    // repeated the iteration `period` back
    double   hist_pose[kCycleMax][12];
    uint32_t hist_npair[kCycleMax], hist_inner[kCycleMax];
    uint32_t cyc_hits[kCycleMax + 1];
    uint32_t dbg_tail[4];  // development probe (last outer iteration): cycles of the final moment sum, of the
                           // Gauss-Newton loop, of the end-of-iteration step; inner iterations
};

struct IcpDevParams
{
    uint32_t max_iterations;
    double   min_abs_step_trans, min_abs_step_rot;
    uint32_t solver_max_iterations;
    double   gn_min_delta;
    int32_t  matcher_kind;
    float    thr;       // (float)distanceThreshold
    float    thr2;      // thr*thr in float (A.5)
    double   distance_threshold;
    double   plane_eigen_threshold;
    uint32_t knn;
    uint32_t min_plane_points;
    uint32_t run_from_iteration, run_up_to_iteration;
    float    q_thr2;    // quality: (float)thresholdDistance squared in float (A.8)
    float    q_thr;
    double   cov_fd_step;
    int32_t  solver_kind;
    // pairingsWeightParameters (row M): consumed by the Horn solver
    int32_t  use_scale_outlier_detector;
    double   scale_outlier_threshold;
    int32_t  use_robust_kernel;
    double   robust_kernel_param, robust_kernel_scale;
    uint32_t detect_cycles;  // period-2 fast-forward: 0 off, 1 bitwise repeats only, 2 repeats within 1e-11 (default)
};

// One pairing as the closed-form (Horn) solver consumes it: the global-side
// point (nearest neighbour, or the plane centroid for Matcher_Point2Plane);
// the local point is pts[position]. 32 bytes, indexed by sorted position.
struct PairRec
{
    double   q[3];
    uint32_t paired;
    uint32_t pad;
};

__host__ __device__ inline uint32_t hash_slot(uint32_t key, uint32_t shift)
{
    return (key * 2654435761u) >> shift;
}

// 10 bits -> every third bit (Morton interleave helper)
__host__ __device__ inline uint32_t spread10(uint32_t v)
{
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// Sort key of a point from its HALF-cell coordinates (hx,hy,hz in
// [0, 2*kFineMax+1]; fine cell = h >> 1, octant bit = h & 1): Morton code of the
// block in the high bits, fine cell inside the block in the next 6 bits, octant
// inside the fine cell in the low 3 -- the memory order of the index, and the
// order queries are binned in.  key >> kMicroBits identifies the fine cell,
// key >> (6 + kMicroBits) the block.
constexpr unsigned long long kInvalidSortKey = 1ull << (36 + kMicroBits);
__host__ __device__ inline unsigned long long point_sort_key(uint32_t hx, uint32_t hy, uint32_t hz)
{
    const uint32_t fx = hx >> 1, fy = hy >> 1, fz = hz >> 1;
    const uint32_t mort = spread10(fx >> 2) | (spread10(fy >> 2) << 1) | (spread10(fz >> 2) << 2);
    const uint32_t sub = (fx & 3u) | ((fy & 3u) << 2) | ((fz & 3u) << 4);
    const uint32_t oct = (hx & 1u) | ((hy & 1u) << 1) | ((hz & 1u) << 2);
    return ((((unsigned long long)mort << 6) | sub) << kMicroBits) | oct;
}
__host__ __device__ inline unsigned long long fine_of_key(unsigned long long k) { return k >> kMicroBits; }
__host__ __device__ inline unsigned long long block_of_key(unsigned long long k) { return k >> (6 + kMicroBits); }

}  // namespace b2
