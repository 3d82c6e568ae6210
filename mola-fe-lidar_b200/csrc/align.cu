// align.cu -- the registration loop of mp2p_icp::ICP::align as the reference
// drives it (LidarOdometry.cpp:869-871),
// resident on the device for the whole iteration loop.
//
// Per outer iteration, two launches over a table of independent jobs:
//   match_kernel  : per local point -- transform (A.2), radius-capped kNN on
//                   the grid index (A.3/A.4), plane fit + gates (A.5), and the
//                   point-to-plane MOMENTS of the pairing, reduced per CTA in a
//                   fixed-shape tree to partials[job][chunk][74] (f64).
//   solve_kernel  : fixed-order reduction of the partials, then the whole
//                   Gauss-Newton inner loop (A.6) on the 12x12 moment matrix,
//                   the SE(3) update, the convergence test (A.7) and the job's
//                   status flags -- no host round trip.
//
// Why moments: the point-to-plane residual r_i(T) = n_i.(R p_i + t - c_i) is
// LINEAR in theta = (R row-major | t) interleaved as 3 rows of (R_i0 R_i1 R_i2 t_i):
// r_i(T) = r_i(T0) + a_i.(theta - theta0), a_i = n_i (x) [p_i;1].  So
//   H = J^T (sum a a^T) J,  g = J^T (sum a r0 + (sum a a^T)(theta - theta0))
// for every inner GN iterate, with J = d theta / d eps (12x6) of the right
// perturbation T (+) exp(eps).  One pass over the pairings per OUTER iteration
// instead of one per inner iteration; the iterates equal the reference's
// per-pairing Gauss-Newton in exact arithmetic.
#include <cub/device/device_radix_sort.cuh>

#include "icp_math.cuh"
#include "knn_search.cuh"
#include "runtime.cuh"

#include <algorithm>
#include <cstring>
#include <map>

namespace b2
{
// index tables for the symmetric 3x3 (nn) and 4x4 (hh) products
__device__ __forceinline__ int sym3(int i, int k)
{
    if (i > k)
    {
        const int t = i;
        i = k, k = t;
    }
    return (i == 0) ? k : (i == 1 ? 2 + k : 5);  // 00 01 02 11 12 22
}
__device__ __forceinline__ int sym4(int j, int l)
{
    if (j > l)
    {
        const int t = j;
        j = l, l = t;
    }
    return (j == 0) ? l : (j == 1 ? 3 + l : (j == 2 ? 5 + l : 9));  // 00 01 02 03 11 12 13 22 23 33
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

struct MatchOut
{
    uint8_t*  paired;
    uint32_t* nn_idx;
    uint32_t* nn_cnt;
    double*   centroid;
    double*   normal;
};

// ------------------------------------------------------------------ matcher
template <int K, bool WRITE>
__global__ void __launch_bounds__(kChunk)
    match_kernel(const CloudView* __restrict__ clouds, const JobDev* __restrict__ jobs,
                 const uint32_t* __restrict__ qorder, const uint32_t* __restrict__ qoff,
                 double* __restrict__ partials, uint32_t max_chunks, IcpDevParams P, MatchOut out)
{
    const uint32_t job = blockIdx.y;
    const JobDev&  J = jobs[job];
    if (J.status != 0) return;
    const CloudView cvL = clouds[J.to_cloud];
    const CloudView cvG = clouds[J.from_cloud];
    const uint32_t  nchunks = (cvL.n + kChunk - 1) / kChunk;
    if (blockIdx.x >= nchunks) return;

    __shared__ GridDev sgrid;
    __shared__ double  sRt[12];
    __shared__ double  sred[kChunk / 32][kNumMoments];
    __shared__ uint32_t s_nvalid;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 12) sRt[tid] = (tid < 9) ? J.R[tid] : J.t[tid - 9];
    if (tid == 32) sgrid = *cvG.grid;
    if (tid == 64) s_nvalid = cvL.grid->n_valid;
    __syncthreads();

    const uint32_t it = J.iter;
    const bool     active = (P.run_from_iteration <= it) &&
                        (P.run_up_to_iteration == 0 || it <= P.run_up_to_iteration);
    // queries are taken in the order of their bin (global fine cell under a
    // recent pose): the lanes of a warp then walk the same shells / blocks
    const uint32_t slot = blockIdx.x * kChunk + tid;
    const uint32_t qi = (slot < cvL.n) ? __ldg(qorder + qoff[job] + slot) : kInvalid;
    const bool     valid = active && (qi < s_nvalid) && (sgrid.n_valid > 0);

    bool   paired = false;
    double nrm[3] = {0, 0, 0}, h[3] = {0, 0, 0}, r0 = 0;

    if (valid)
    {
        const float4 pl = __ldg(cvL.pts + qi);
        const double px = pl.x, py = pl.y, pz = pl.z;
        // A.2: q = fl32(R p + t), f64 accumulate in this fixed order
        const double gx = ((sRt[0] * px + sRt[1] * py) + sRt[2] * pz) + sRt[9];
        const double gy = ((sRt[3] * px + sRt[4] * py) + sRt[5] * pz) + sRt[10];
        const double gz = ((sRt[6] * px + sRt[7] * py) + sRt[8] * pz) + sRt[11];
        const float  qx = (float)gx, qy = (float)gy, qz = (float)gz;

        uint64_t key[K];
#pragma unroll
        for (int i = 0; i < K; i++) key[i] = sentinel_key(P.thr2);
        knn_search<K>(cvG, sgrid, qx, qy, qz, P.thr2, key);

        // neighbours kept after the distance cut; K may exceed the configured knn
        const uint64_t sent = sentinel_key(P.thr2);
        uint32_t       m = 0;
#pragma unroll
        for (int i = 0; i < K; i++)
            if ((uint32_t)i < P.knn && key[i] != sent) m++;

        const uint32_t orig = __float_as_uint(pl.w);
        if (WRITE)
        {
            if (out.nn_cnt) out.nn_cnt[orig] = m;
            if (out.nn_idx)
#pragma unroll
                for (int i = 0; i < K; i++)
                    if ((uint32_t)i < P.knn)
                        out.nn_idx[(size_t)orig * P.knn + i] =
                            ((uint32_t)i < m) ? key_idx(key[i]) : kInvalid;
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
                    const uint32_t pos = __ldg(cvG.rank + key_idx(key[i]));
                    const float4   pn = __ldg(cvG.pts + pos);
                    nx_[i] = (double)pn.x, ny_[i] = (double)pn.y, nz_[i] = (double)pn.z;
                    sx += nx_[i], sy += ny_[i], sz += nz_[i];
                }
            const double inv = 1.0 / (double)m;
            const double cx = sx * inv, cy = sy * inv, cz = sz * inv;
            double c00 = 0, c01 = 0, c02 = 0, c11 = 0, c12 = 0, c22 = 0;
#pragma unroll
            for (int i = 0; i < K; i++)
                if ((uint32_t)i < m)
                {
                    const double dx = nx_[i] - cx, dy = ny_[i] - cy, dz = nz_[i] - cz;
                    c00 += dx * dx, c01 += dx * dy, c02 += dx * dz;
                    c11 += dy * dy, c12 += dy * dz, c22 += dz * dz;
                }
            double C[9] = {c00 * inv, c01 * inv, c02 * inv, c01 * inv, c11 * inv,
                           c12 * inv, c02 * inv, c12 * inv, c22 * inv};
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
                    h[0] = px, h[1] = py, h[2] = pz;
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
    }

    // ---- moments, reduced warp (xor butterfly) -> CTA in a fixed tree ------
    const unsigned any = __ballot_sync(0xFFFFFFFFu, paired);
    if (any == 0)
    {
        for (int i = lane; i < kNumMoments; i += 32) sred[warp][i] = 0.0;
    }
    else
    {
        const double hh[10] = {h[0] * h[0], h[0] * h[1], h[0] * h[2], h[0],        h[1] * h[1],
                               h[1] * h[2], h[1],        h[2] * h[2], h[2],        paired ? 1.0 : 0.0};
        const double nn[6] = {nrm[0] * nrm[0], nrm[0] * nrm[1], nrm[0] * nrm[2],
                              nrm[1] * nrm[1], nrm[1] * nrm[2], nrm[2] * nrm[2]};
#pragma unroll
        for (int u = 0; u < 6; u++)
#pragma unroll
            for (int v = 0; v < 10; v++)
            {
                const double s = warp_sum(nn[u] * hh[v]);
                if (lane == 0) sred[warp][u * 10 + v] = s;
            }
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            const double rn = r0 * nrm[i];
#pragma unroll
            for (int j = 0; j < 4; j++)
            {
                const double s = warp_sum(j < 3 ? rn * h[j] : rn);
                if (lane == 0) sred[warp][60 + i * 4 + j] = s;
            }
        }
        const double se = warp_sum(r0 * r0);
        const double sc = warp_sum(paired ? 1.0 : 0.0);
        if (lane == 0) sred[warp][72] = se, sred[warp][73] = sc;
    }
    __syncthreads();
    if (tid < kNumMoments)
    {
        const double s = (sred[0][tid] + sred[1][tid]) + (sred[2][tid] + sred[3][tid]);
        partials[((size_t)job * max_chunks + blockIdx.x) * kNumMoments + tid] = s;
    }
}

// ------------------------------------------------------------------- solver
// One CTA per job. Warp 0 runs the Gauss-Newton inner loop on the moments.
constexpr int kSolveThreads = 640;
constexpr int kSolveGroups = 8;  // 8 x 74 threads share the partials reduction

__global__ void __launch_bounds__(kSolveThreads)
    solve_kernel(const CloudView* __restrict__ clouds, JobDev* __restrict__ jobs,
                 const double* __restrict__ partials, uint32_t max_chunks, IcpDevParams P,
                 uint32_t* __restrict__ n_active)
{
    const uint32_t job = blockIdx.x;
    JobDev&        J = jobs[job];
    if (J.status != 0) return;
    const int tid = threadIdx.x, lane = tid & 31;

    __shared__ double sM[kNumMoments];
    __shared__ double sPart[kSolveGroups][kNumMoments];
    __shared__ double sA[144];
    __shared__ double sT0[12];   // theta0 layout: [R_i0 R_i1 R_i2 t_i] x 3
    __shared__ double sR[9], st[3];
    __shared__ double sX[12], sG12[12], sJ[72], sB[72], sH[36], sg[6];
    __shared__ int    sStop;

    // fixed-order reduction of the per-CTA partials: group g takes the chunks
    // c = g (mod 8) with two interleaved accumulators, then the 8 group sums
    // are added as a balanced tree -- the order never depends on scheduling
    const uint32_t nchunks = (clouds[J.to_cloud].n + kChunk - 1) / kChunk;
    if (tid < kSolveGroups * kNumMoments)
    {
        const int     grp = tid / kNumMoments, comp = tid % kNumMoments;
        const double* p = partials + (size_t)job * max_chunks * kNumMoments + comp;
        double        a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        uint32_t      c = grp;
        // eight loads in flight per thread: the partials sit in L2, the loop
        // is latency bound
        for (; c + 7 * kSolveGroups < nchunks; c += 8 * kSolveGroups)
        {
            const double v0 = p[(size_t)c * kNumMoments];
            const double v1 = p[(size_t)(c + kSolveGroups) * kNumMoments];
            const double v2 = p[(size_t)(c + 2 * kSolveGroups) * kNumMoments];
            const double v3 = p[(size_t)(c + 3 * kSolveGroups) * kNumMoments];
            const double v4 = p[(size_t)(c + 4 * kSolveGroups) * kNumMoments];
            const double v5 = p[(size_t)(c + 5 * kSolveGroups) * kNumMoments];
            const double v6 = p[(size_t)(c + 6 * kSolveGroups) * kNumMoments];
            const double v7 = p[(size_t)(c + 7 * kSolveGroups) * kNumMoments];
            a0 += v0, a1 += v1, a2 += v2, a3 += v3;
            a0 += v4, a1 += v5, a2 += v6, a3 += v7;
        }
        for (; c < nchunks; c += kSolveGroups) a0 += p[(size_t)c * kNumMoments];
        sPart[grp][comp] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    if (tid < kNumMoments)
    {
        const double s = ((sPart[0][tid] + sPart[1][tid]) + (sPart[2][tid] + sPart[3][tid])) +
                         ((sPart[4][tid] + sPart[5][tid]) + (sPart[6][tid] + sPart[7][tid]));
        sM[tid] = s;
        J.M[tid] = s;
    }
    __syncthreads();
    const uint32_t npair = (uint32_t)(sM[73] + 0.5);
    if (npair == 0)
    {
        if (tid == 0)
        {
            J.n_pairings = 0;
            J.status = 1;
            J.term_reason = B200ICP_TERM_NO_PAIRINGS;
            atomicSub(n_active, 1u);
        }
        return;
    }
    for (int e = tid; e < 144; e += blockDim.x)
    {
        const int a = e / 12, b = e % 12;
        sA[e] = sM[sym3(a >> 2, b >> 2) * 10 + sym4(a & 3, b & 3)];
    }
    if (tid < 12)
    {
        const int    i = tid >> 2, j = tid & 3;
        const double v = (j < 3) ? J.R[i * 3 + j] : J.t[i];
        sT0[tid] = v;
        if (j < 3)
            sR[i * 3 + j] = v;
        else
            st[i] = v;
    }
    __syncthreads();
    if (tid >= 32) return;

    // ---- Gauss-Newton on the moments (A.6), warp 0 -------------------------
    uint32_t inner = 0;
    for (uint32_t iter = 0; iter < P.solver_max_iterations; iter++)
    {
        // x = theta(T) - theta0
        if (lane < 12)
        {
            const int i = lane >> 2, j = lane & 3;
            sX[lane] = ((j < 3) ? sR[i * 3 + j] : st[i]) - sT0[lane];
        }
        // J = d theta / d eps (12 x 6): columns v0..2 then w0..2
        for (int e = lane; e < 72; e += 32)
        {
            const int a = e / 6, c = e % 6, i = a >> 2, l = a & 3;
            double    v = 0.0;
            if (c < 3)
                v = (l == 3) ? sR[i * 3 + c] : 0.0;
            else if (l < 3)
            {
                const int j = c - 3;
                // (R [e_j]x)[i][l]
                if (j == 0)
                    v = (l == 1) ? sR[i * 3 + 2] : (l == 2 ? -sR[i * 3 + 1] : 0.0);
                else if (j == 1)
                    v = (l == 0) ? -sR[i * 3 + 2] : (l == 2 ? sR[i * 3 + 0] : 0.0);
                else
                    v = (l == 0) ? sR[i * 3 + 1] : (l == 1 ? -sR[i * 3 + 0] : 0.0);
            }
            sJ[e] = v;
        }
        __syncwarp();
        // g12 = s + A x
        if (lane < 12)
        {
            double acc = sM[60 + lane];
#pragma unroll
            for (int b = 0; b < 12; b++) acc += sA[lane * 12 + b] * sX[b];
            sG12[lane] = acc;
        }
        // B = A J
        for (int e = lane; e < 72; e += 32)
        {
            const int a = e / 6, c = e % 6;
            double    acc = 0;
#pragma unroll
            for (int b = 0; b < 12; b++) acc += sA[a * 12 + b] * sJ[b * 6 + c];
            sB[e] = acc;
        }
        __syncwarp();
        // H = J^T B, g = J^T g12
        for (int e = lane; e < 36; e += 32)
        {
            const int r = e / 6, c = e % 6;
            double    acc = 0;
#pragma unroll
            for (int a = 0; a < 12; a++) acc += sJ[a * 6 + r] * sB[a * 6 + c];
            sH[e] = acc;
        }
        if (lane < 6)
        {
            double acc = 0;
#pragma unroll
            for (int a = 0; a < 12; a++) acc += sJ[a * 6 + lane] * sG12[a];
            sg[lane] = acc;
        }
        __syncwarp();
        if (lane == 0)
        {
            double Hm[36], mg[6], delta[6];
            for (int i = 0; i < 36; i++) Hm[i] = 0.5 * (sH[i] + sH[(i % 6) * 6 + i / 6]);
            for (int i = 0; i < 6; i++) mg[i] = -sg[i];
            solve6_spd(Hm, mg, delta);
            Pose T, dT, Tn;
            for (int i = 0; i < 9; i++) T.R[i] = sR[i];
            for (int i = 0; i < 3; i++) T.t[i] = st[i];
            se3_exp(delta, dT);
            pose_compose(T, dT, Tn);
            for (int i = 0; i < 9; i++) sR[i] = Tn.R[i];
            for (int i = 0; i < 3; i++) st[i] = Tn.t[i];
            double nd = 0;
            for (int i = 0; i < 6; i++) nd += delta[i] * delta[i];
            sStop = (sqrt(nd) < P.gn_min_delta) ? 1 : 0;
        }
        __syncwarp();
        inner++;
        if (sStop) break;
    }
    if (lane == 0)
    {
        // convergence (A.7): delta = log(T0^-1 * Tnew), T0 = previous solution
        Pose T0, Tn, dT;
        for (int i = 0; i < 9; i++) T0.R[i] = J.R[i], Tn.R[i] = sR[i];
        for (int i = 0; i < 3; i++) T0.t[i] = J.t[i], Tn.t[i] = st[i];
        pose_inverse_compose(T0, Tn, dT);
        double d[6];
        se3_log(dT, d);
        const double dxyz = sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
        const double drot = sqrt((d[3] * d[3] + d[4] * d[4]) + d[5] * d[5]);
        for (int i = 0; i < 9; i++) J.Rprev[i] = T0.R[i], J.R[i] = Tn.R[i];
        for (int i = 0; i < 3; i++) J.tprev[i] = T0.t[i], J.t[i] = Tn.t[i];
        J.n_pairings = npair;
        J.inner_iters_total += inner;
        if (dxyz < P.min_abs_step_trans && drot < P.min_abs_step_rot)
        {
            J.status = 1;
            J.term_reason = B200ICP_TERM_STALLED;
            atomicSub(n_active, 1u);
        }
        else
        {
            const uint32_t it = J.iter + 1;
            J.iter = it;
            if (it >= P.max_iterations)
            {
                J.status = 1;
                J.term_reason = B200ICP_TERM_MAX_ITERATIONS;
                atomicSub(n_active, 1u);
            }
        }
    }
}

// ------------------------------------------------------------------ quality
// QualityEvaluator_PairedRatio (row O / A.8): 1-NN within thresholdDistance.
__global__ void __launch_bounds__(kChunk)
    quality_kernel(const CloudView* __restrict__ clouds, JobDev* __restrict__ jobs,
                   const uint32_t* __restrict__ qorder, const uint32_t* __restrict__ qoff,
                   IcpDevParams P)
{
    const uint32_t  job = blockIdx.y;
    JobDev&         J = jobs[job];
    const CloudView cvL = clouds[J.to_cloud];
    const CloudView cvG = clouds[J.from_cloud];
    const uint32_t  nchunks = (cvL.n + kChunk - 1) / kChunk;
    if (blockIdx.x >= nchunks) return;
    __shared__ GridDev sgrid;
    __shared__ double  sRt[12];
    __shared__ uint32_t s_nvalid;
    const int tid = threadIdx.x;
    if (tid < 12) sRt[tid] = (tid < 9) ? J.R[tid] : J.t[tid - 9];
    if (tid == 32) sgrid = *cvG.grid;
    if (tid == 64) s_nvalid = cvL.grid->n_valid;
    __syncthreads();
    const uint32_t slot = blockIdx.x * kChunk + tid;
    const uint32_t qi = (slot < cvL.n) ? __ldg(qorder + qoff[job] + slot) : kInvalid;
    bool           hit = false;
    if (qi < s_nvalid && sgrid.n_valid > 0)
    {
        const float4 pl = __ldg(cvL.pts + qi);
        const double px = pl.x, py = pl.y, pz = pl.z;
        const float  qx = (float)(((sRt[0] * px + sRt[1] * py) + sRt[2] * pz) + sRt[9]);
        const float  qy = (float)(((sRt[3] * px + sRt[4] * py) + sRt[5] * pz) + sRt[10]);
        const float  qz = (float)(((sRt[6] * px + sRt[7] * py) + sRt[8] * pz) + sRt[11]);
        uint64_t     key[1] = {sentinel_key(P.q_thr2)};
        knn_search<1>(cvG, sgrid, qx, qy, qz, P.q_thr2, key);
        hit = (key[0] != sentinel_key(P.q_thr2)) && (key_d2(key[0]) < P.q_thr2);  // strict
    }
    const unsigned b = __ballot_sync(0xFFFFFFFFu, hit);
    if ((tid & 31) == 0 && b) atomicAdd(&J.quality_count, (uint32_t)__popc(b));
}

// --------------------------------------------------------------- covariance
// mp2p_icp::covariance (row P / A.9): forward-difference Jacobian of the
// stacked residuals wrt (x,y,z,yaw,pitch,roll) at the solution, H = J^T J,
// cov = H^-1.  With linear-in-theta residuals J^T J = dTheta^T A dTheta.
__global__ void __launch_bounds__(64) covariance_kernel(JobDev* __restrict__ jobs, IcpDevParams P)
{
    JobDev&   J = jobs[blockIdx.x];
    const int tid = threadIdx.x;
    __shared__ double sA[144], sD[72], sH[36];
    if (J.n_pairings == 0)
    {
        if (tid < 36) J.cov[tid] = 0.0;
        if (tid == 0) J.cov_singular = 1;
        return;
    }
    for (int e = tid; e < 144; e += blockDim.x)
    {
        const int a = e / 12, b = e % 12;
        sA[e] = J.M[sym3(a >> 2, b >> 2) * 10 + sym4(a & 3, b & 3)];
    }
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

// ---------------------------------------------------------------- kNN query
template <int K>
__global__ void __launch_bounds__(kChunk)
    knn_kernel(CloudView cvG, CloudView cvL, const uint32_t* __restrict__ qorder, Pose T,
               uint32_t k, float cap_d2, uint32_t* __restrict__ idx_out,
               float* __restrict__ d2_out)
{
    __shared__ GridDev sgrid;
    __shared__ uint32_t s_nvalid;
    if (threadIdx.x == 0) sgrid = *cvG.grid;
    if (threadIdx.x == 32) s_nvalid = cvL.grid->n_valid;
    __syncthreads();
    const uint32_t slot = blockIdx.x * kChunk + threadIdx.x;
    if (slot >= cvL.n) return;
    const uint32_t qi = __ldg(qorder + slot);
    if (qi >= s_nvalid) return;
    const float4 pl = __ldg(cvL.pts + qi);
    const double px = pl.x, py = pl.y, pz = pl.z;
    const float  qx = (float)(((T.R[0] * px + T.R[1] * py) + T.R[2] * pz) + T.t[0]);
    const float  qy = (float)(((T.R[3] * px + T.R[4] * py) + T.R[5] * pz) + T.t[1]);
    const float  qz = (float)(((T.R[6] * px + T.R[7] * py) + T.R[8] * pz) + T.t[2]);
    uint64_t     key[K];
#pragma unroll
    for (int i = 0; i < K; i++) key[i] = sentinel_key(cap_d2);
    if (sgrid.n_valid > 0) knn_search<K>(cvG, sgrid, qx, qy, qz, cap_d2, key);
    const uint32_t orig = __float_as_uint(pl.w);
    const uint64_t sent = sentinel_key(cap_d2);
#pragma unroll
    for (int i = 0; i < K; i++)
        if ((uint32_t)i < k)
        {
            const bool ok = key[i] != sent;
            idx_out[(size_t)orig * k + i] = ok ? key_idx(key[i]) : kInvalid;
            d2_out[(size_t)orig * k + i] = ok ? key_d2(key[i]) : INFINITY;
        }
}

// ------------------------------------------------------------- query binning
// Key of every query of every job = (job, sort key of the GLOBAL fine cell its
// transformed position falls in). Sorting these pairs makes the lanes of a
// warp share home cells, hence shells, blocks and candidate ranges. The order
// only affects which lanes work together and the (fixed) summation order.
__global__ void __launch_bounds__(kChunk)
    bin_key_kernel(const CloudView* __restrict__ clouds, const JobDev* __restrict__ jobs,
                   const uint32_t* __restrict__ qoff, unsigned long long* __restrict__ keys,
                   uint32_t* __restrict__ vals)
{
    const uint32_t  job = blockIdx.y;
    const JobDev&   J = jobs[job];
    const CloudView cvL = clouds[J.to_cloud];
    const uint32_t  slot = blockIdx.x * kChunk + threadIdx.x;
    if (slot >= cvL.n) return;
    const GridDev* g = clouds[J.from_cloud].grid;
    unsigned long long key = kInvalidSortKey;
    if (slot < cvL.grid->n_valid)
    {
        const float4 pl = __ldg(cvL.pts + slot);
        const double px = pl.x, py = pl.y, pz = pl.z;
        const float  qx = (float)(((J.R[0] * px + J.R[1] * py) + J.R[2] * pz) + J.t[0]);
        const float  qy = (float)(((J.R[3] * px + J.R[4] * py) + J.R[5] * pz) + J.t[1]);
        const float  qz = (float)(((J.R[6] * px + J.R[7] * py) + J.R[8] * pz) + J.t[2]);
        const float  inv = g->inv_cell, hi = (float)kFineMax;
        const float  ux = fminf(fmaxf((qx - g->ox) * inv, 0.0f), hi);
        const float  uy = fminf(fmaxf((qy - g->oy) * inv, 0.0f), hi);
        const float  uz = fminf(fmaxf((qz - g->oz) * inv, 0.0f), hi);
        if (ux == ux && uy == uy && uz == uz)
            key = fine_sort_key((uint32_t)ux, (uint32_t)uy, (uint32_t)uz);
    }
    const size_t o = (size_t)qoff[job] + slot;
    keys[o] = ((unsigned long long)job << 37) | key;
    vals[o] = slot;
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

// =========================================================== host orchestration
static int wait_cloud(Workspace* ws, const b200icp_cloud* c)
{
    B2_CUDA_TRY(cudaStreamWaitEvent(ws->stream, c->ready, 0));
    return B200ICP_OK;
}

// Bins the queries of every job by the global fine cell of their transformed
// position (one key kernel + one radix sort over all jobs of the wave).
struct Binner
{
    unsigned long long *k0 = nullptr, *k1 = nullptr;
    uint32_t *          v0 = nullptr, *v1 = nullptr, *qoff = nullptr;
    void*               temp = nullptr;
    size_t              temp_bytes = 0, total = 0, njobs = 0;
    int                 end_bit = 37;
    const uint32_t*     order = nullptr;  // queries of job j: order[qoff[j] ...]

    int plan(size_t total_queries, size_t jobs, cudaStream_t s)
    {
        total = total_queries ? total_queries : 1;
        njobs = jobs;
        end_bit = 37;
        while (end_bit < 64 && (1ull << (end_bit - 37)) < jobs) end_bit++;
        cub::DoubleBuffer<unsigned long long> dk(nullptr, nullptr);
        cub::DoubleBuffer<uint32_t>           dv(nullptr, nullptr);
        B2_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, dk, dv, (int)total, 0, end_bit, s));
        return B200ICP_OK;
    }
    void layout(Carver& c)
    {
        k0 = c.take<unsigned long long>(total), k1 = c.take<unsigned long long>(total);
        v0 = c.take<uint32_t>(total), v1 = c.take<uint32_t>(total);
        qoff = c.take<uint32_t>(njobs + 1);
        temp = c.take<char>(temp_bytes);
    }
    int run(Workspace* ws, dim3 grid, const CloudView* d_clouds, const JobDev* d_jobs)
    {
        cudaStream_t s = ws->stream;
        bin_key_kernel<<<grid, kChunk, 0, s>>>(d_clouds, d_jobs, qoff, k0, v0);
        cub::DoubleBuffer<unsigned long long> keys(k0, k1);
        cub::DoubleBuffer<uint32_t>           vals(v0, v1);
        B2_CUDA_TRY(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, vals, (int)total, 0, end_bit, s));
        order = vals.Current();
        ws->launches += 1 + 1 + (end_bit + 7) / 8;
        return B200ICP_OK;
    }
};

template <bool WRITE>
static void launch_match(Workspace* ws, uint32_t knn, dim3 grid, const CloudView* d_clouds,
                         const JobDev* d_jobs, const Binner& bin, double* d_partials,
                         uint32_t max_chunks, const IcpDevParams& D, const MatchOut& mo)
{
    if (knn == 6)
        match_kernel<6, WRITE><<<grid, kChunk, 0, ws->stream>>>(d_clouds, d_jobs, bin.order, bin.qoff,
                                                                d_partials, max_chunks, D, mo);
    else if (knn <= 4)
        match_kernel<4, WRITE><<<grid, kChunk, 0, ws->stream>>>(d_clouds, d_jobs, bin.order, bin.qoff,
                                                                d_partials, max_chunks, D, mo);
    else
        match_kernel<8, WRITE><<<grid, kChunk, 0, ws->stream>>>(d_clouds, d_jobs, bin.order, bin.qoff,
                                                                d_partials, max_chunks, D, mo);
    ws->launches++;
}

static int check_supported(const ::b200icp* ctx)
{
    const auto& P = ctx->P;
    if (P.matcher_kind != B200ICP_MATCHER_POINT2PLANE || P.solver_kind != B200ICP_SOLVER_GAUSS_NEWTON)
    {
        set_error("this build runs Matcher_Point2Plane + Solver_GaussNewton on the device; "
                  "matcher_kind=%d solver_kind=%d is not available yet",
                  P.matcher_kind, P.solver_kind);
        return B200ICP_ERR_UNSUPPORTED;
    }
    if (P.knn < 1 || P.knn > B200ICP_MAX_KNN)
    {
        set_error("knn=%u outside [1,%d]", P.knn, B200ICP_MAX_KNN);
        return B200ICP_ERR_UNSUPPORTED;
    }
    if (P.use_robust_kernel)
    {
        set_error("use_robust_kernel=true is not available with Solver_GaussNewton");
        return B200ICP_ERR_UNSUPPORTED;
    }
    return B200ICP_OK;
}

// Jobs are processed in waves that bound the partials buffer.
static int run_wave(::b200icp* ctx, Workspace* ws, size_t n, const b200icp_cloud* const* from,
                    const b200icp_cloud* const* to, const double* guesses, b200icp_result_t* out)
{
    const IcpDevParams& D = ctx->D;
    cudaStream_t        s = ws->stream;
    // unique cloud table
    std::map<const b200icp_cloud*, uint32_t> cmap;
    std::vector<CloudView>                   views;
    uint32_t                                 max_chunks = 1;
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
        const uint32_t ch = (uint32_t)((to[j]->n + kChunk - 1) / kChunk);
        max_chunks = std::max(max_chunks, ch);
        total_queries += to[j]->n;
    }
    for (auto& kv : cmap)
        if (int r = wait_cloud(ws, kv.first)) return r;

    Binner bin;
    if (int r = bin.plan(total_queries, n, s)) return r;
    Carver sz(nullptr);
    auto layout = [&](Carver& k, CloudView*& dc, JobDev*& dj, double*& dp, uint32_t*& da) {
        dc = k.take<CloudView>(views.size());
        dj = k.take<JobDev>(n);
        dp = k.take<double>((size_t)n * max_chunks * kNumMoments);
        da = k.take<uint32_t>(4);
        bin.layout(k);
    };
    CloudView* d_clouds;
    JobDev*    d_jobs;
    double*    d_partials;
    uint32_t*  d_active;
    layout(sz, d_clouds, d_jobs, d_partials, d_active);
    if (int r = ws->reserve_device(sz.off)) return r;
    Carver real(ws->d_scratch);
    layout(real, d_clouds, d_jobs, d_partials, d_active);
    // pinned staging: views | jobs | query offsets | active flags (2 slots) | n_active init
    const size_t off_jobs = align_up(views.size() * sizeof(CloudView));
    const size_t off_qoff = off_jobs + align_up(n * sizeof(JobDev));
    const size_t off_flags = off_qoff + align_up((n + 1) * sizeof(uint32_t));
    if (int r = ws->reserve_pinned(off_flags + 256)) return r;
    char*      hp = (char*)ws->h_pinned;
    CloudView* h_views = (CloudView*)hp;
    JobDev*    h_jobs = (JobDev*)(hp + off_jobs);
    uint32_t*  h_qoff = (uint32_t*)(hp + off_qoff);
    uint32_t*  h_flags = (uint32_t*)(hp + off_flags);
    memcpy(h_views, views.data(), views.size() * sizeof(CloudView));
    memcpy(h_jobs, hjobs.data(), n * sizeof(JobDev));
    h_qoff[0] = 0;
    for (size_t j = 0; j < n; j++) h_qoff[j + 1] = h_qoff[j] + (uint32_t)to[j]->n;
    h_flags[0] = h_flags[1] = 0xFFFFFFFFu;
    h_flags[2] = (uint32_t)n;
    B2_CUDA_TRY(cudaMemcpyAsync(d_clouds, h_views, views.size() * sizeof(CloudView),
                                cudaMemcpyHostToDevice, s));
    B2_CUDA_TRY(cudaMemcpyAsync(d_jobs, h_jobs, n * sizeof(JobDev), cudaMemcpyHostToDevice, s));
    B2_CUDA_TRY(cudaMemcpyAsync(bin.qoff, h_qoff, (n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    B2_CUDA_TRY(cudaMemcpyAsync(d_active, h_flags + 2, sizeof(uint32_t), cudaMemcpyHostToDevice, s));

    const bool prof = ctx->profile_on;
    if (prof)
        if (int r = ws->reserve_prof_events(2 * (size_t)D.max_iterations + 2)) return r;

    const dim3     mgrid(max_chunks, (unsigned)n);
    const MatchOut no_out = {nullptr, nullptr, nullptr, nullptr, nullptr};
    const uint32_t kBatch = 4;
    uint32_t       enq = 0, batch = 0;
    bool           finished = false;
    while (enq < D.max_iterations && !finished)
    {
        const uint32_t todo = std::min(kBatch, D.max_iterations - enq);
        for (uint32_t i = 0; i < todo; i++, enq++)
        {
            // re-bin while the pose still moves by more than a fine cell (first
            // iterations), then every 8th iteration
            if (enq < 3 || (enq & 7u) == 0)
                if (int r = bin.run(ws, mgrid, d_clouds, d_jobs)) return r;
            if (prof) B2_CUDA_TRY(cudaEventRecord(ws->prof_ev[4 * enq + 0], s));
            launch_match<false>(ws, D.knn, mgrid, d_clouds, d_jobs, bin, d_partials, max_chunks, D, no_out);
            if (prof) B2_CUDA_TRY(cudaEventRecord(ws->prof_ev[4 * enq + 1], s));
            if (prof) B2_CUDA_TRY(cudaEventRecord(ws->prof_ev[4 * enq + 2], s));
            solve_kernel<<<(unsigned)n, kSolveThreads, 0, s>>>(d_clouds, d_jobs, d_partials, max_chunks, D,
                                                     d_active);
            ws->launches++;
            if (prof) B2_CUDA_TRY(cudaEventRecord(ws->prof_ev[4 * enq + 3], s));
        }
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
        batch++;
    }
    const dim3 qgrid(max_chunks, (unsigned)n);
    if (D.max_iterations == 0)
        if (int r = bin.run(ws, mgrid, d_clouds, d_jobs)) return r;
    quality_kernel<<<qgrid, kChunk, 0, s>>>(d_clouds, d_jobs, bin.order, bin.qoff, D);
    covariance_kernel<<<(unsigned)n, 64, 0, s>>>(d_jobs, D);
    ws->launches += 2;
    B2_CUDA_TRY(cudaMemcpyAsync(h_jobs, d_jobs, n * sizeof(JobDev), cudaMemcpyDeviceToHost, s));
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
    if (prof)
    {
        double mm = 0, sm = 0;
        for (uint32_t i = 0; i < std::min(max_runs, enq); i++)
        {
            float a = 0, b = 0;
            B2_CUDA_TRY(cudaEventElapsedTime(&a, ws->prof_ev[4 * i + 0], ws->prof_ev[4 * i + 1]));
            B2_CUDA_TRY(cudaEventElapsedTime(&b, ws->prof_ev[4 * i + 2], ws->prof_ev[4 * i + 3]));
            mm += a, sm += b;
        }
        std::lock_guard<std::mutex> lk(ctx->mtx);
        ctx->prof.match_launches += std::min(max_runs, enq);
        ctx->prof.match_ms += mm;
        // queries examined by those launches (single job: exact; batches: upper bound)
        ctx->prof.match_queries += (uint64_t)std::min(max_runs, enq) * total_queries;
        ctx->prof.solve_launches += std::min(max_runs, enq);
        ctx->prof.solve_ms += sm;
    }
    return B200ICP_OK;
}

int run_align_batch(::b200icp* ctx, size_t n, const b200icp_cloud* const* from,
                    const b200icp_cloud* const* to, const double* guesses, b200icp_result_t* out)
{
    if (int r = check_supported(ctx)) return r;
    if (n == 0) return B200ICP_OK;
    Lease L(ctx);
    if (!L.ws) return B200ICP_ERR_CUDA;
    // wave size: partials <= ~1.5 GB and gridDim.y <= 65535
    size_t i = 0;
    while (i < n)
    {
        size_t   cnt = 0;
        uint32_t mc = 1;
        while (i + cnt < n && cnt < 65535)
        {
            const uint32_t ch = (uint32_t)((to[i + cnt]->n + kChunk - 1) / kChunk);
            const uint32_t nmc = std::max(mc, ch);
            if (cnt > 0 && (size_t)(cnt + 1) * nmc * kNumMoments * sizeof(double) > (1536ull << 20)) break;
            mc = nmc;
            cnt++;
        }
        if (int r = run_wave(ctx, L.ws, cnt, from + i, to + i, guesses + 6 * i, out + i)) return r;
        i += cnt;
    }
    return B200ICP_OK;
}

// one (from, to, pose) job laid out, uploaded and binned; `extra` carves the
// caller's own arrays out of the same scratch allocation
struct SingleJob
{
    CloudView* d_clouds = nullptr;
    JobDev*    d_jobs = nullptr;
    Binner     bin;
    Pose       T;
};

template <typename F>
static int single_job_setup(Workspace* ws, const b200icp_cloud* from, const b200icp_cloud* to,
                            const double* pose6, uint32_t iter, SingleJob& sj, F&& extra)
{
    cudaStream_t s = ws->stream;
    if (int r = wait_cloud(ws, from)) return r;
    if (int r = wait_cloud(ws, to)) return r;
    if (int r = sj.bin.plan(to->n, 1, s)) return r;
    auto layout = [&](Carver& c) {
        sj.d_clouds = c.take<CloudView>(2);
        sj.d_jobs = c.take<JobDev>(1);
        sj.bin.layout(c);
        extra(c);
    };
    Carver sz(nullptr);
    layout(sz);
    if (int r = ws->reserve_device(sz.off)) return r;
    Carver real(ws->d_scratch);
    layout(real);
    const size_t off_job = align_up(2 * sizeof(CloudView));
    const size_t off_q = off_job + align_up(sizeof(JobDev));
    if (int r = ws->reserve_pinned(off_q + 64)) return r;
    CloudView* hv = (CloudView*)ws->h_pinned;
    JobDev*    hj = (JobDev*)((char*)ws->h_pinned + off_job);
    uint32_t*  hq = (uint32_t*)((char*)ws->h_pinned + off_q);
    hv[0] = from->view(), hv[1] = to->view();
    memset(hj, 0, sizeof(JobDev));
    const double ident[6] = {0, 0, 0, 0, 0, 0};
    pose_from_ypr(pose6 ? pose6 : ident, sj.T);
    memcpy(hj->R, sj.T.R, sizeof(sj.T.R)), memcpy(hj->t, sj.T.t, sizeof(sj.T.t));
    hj->from_cloud = 0, hj->to_cloud = 1;
    hj->iter = iter;
    hq[0] = 0, hq[1] = (uint32_t)to->n;
    B2_CUDA_TRY(cudaMemcpyAsync(sj.d_clouds, hv, 2 * sizeof(CloudView), cudaMemcpyHostToDevice, s));
    B2_CUDA_TRY(cudaMemcpyAsync(sj.d_jobs, hj, sizeof(JobDev), cudaMemcpyHostToDevice, s));
    B2_CUDA_TRY(cudaMemcpyAsync(sj.bin.qoff, hq, 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    const uint32_t chunks = (uint32_t)((to->n + kChunk - 1) / kChunk);
    return sj.bin.run(ws, dim3(chunks ? chunks : 1, 1), sj.d_clouds, sj.d_jobs);
}

int run_knn(::b200icp* ctx, const b200icp_cloud* ref, const b200icp_cloud* q, const double* pose6,
            uint32_t k, float max_dist, uint32_t* idx_out, float* d2_out)
{
    if (k < 1 || k > B200ICP_MAX_KNN)
    {
        set_error("k=%u outside [1,%d]", k, B200ICP_MAX_KNN);
        return B200ICP_ERR_BAD_ARG;
    }
    if (!(max_dist > 0) || !std::isfinite(max_dist))
    {
        set_error("max_dist must be positive and finite (radius-capped search)");
        return B200ICP_ERR_BAD_ARG;
    }
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
    const int             blocks = (int)((nq + kChunk - 1) / kChunk);
    const CloudView       vr = ref->view(), vq = q->view();
    const uint32_t* const ord = sj.bin.order;
    if (k == 1)
        knn_kernel<1><<<blocks, kChunk, 0, s>>>(vr, vq, ord, sj.T, k, cap_d2, d_idx, d_d2);
    else if (k <= 4)
        knn_kernel<4><<<blocks, kChunk, 0, s>>>(vr, vq, ord, sj.T, k, cap_d2, d_idx, d_d2);
    else if (k <= 6)
        knn_kernel<6><<<blocks, kChunk, 0, s>>>(vr, vq, ord, sj.T, k, cap_d2, d_idx, d_d2);
    else
        knn_kernel<8><<<blocks, kChunk, 0, s>>>(vr, vq, ord, sj.T, k, cap_d2, d_idx, d_d2);
    ws->launches++;
    if (prof) B2_CUDA_TRY(cudaEventRecord(ws->prof_ev[1], s));
    B2_CUDA_TRY(cudaGetLastError());
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

int run_match(::b200icp* ctx, const b200icp_cloud* from, const b200icp_cloud* to,
              const double* pose6, uint8_t* paired, uint32_t* nn_idx, uint32_t* nn_cnt,
              double* centroid, double* normal, uint32_t* n_pairings)
{
    if (int r = check_supported(ctx)) return r;
    Lease L(ctx);
    if (!L.ws) return B200ICP_ERR_CUDA;
    Workspace*          ws = L.ws;
    cudaStream_t        s = ws->stream;
    const IcpDevParams& D = ctx->D;
    const size_t        n = to->n, k = D.knn;
    if (n_pairings) *n_pairings = 0;
    if (n == 0) return B200ICP_OK;
    const uint32_t max_chunks = (uint32_t)((n + kChunk - 1) / kChunk);
    double*        d_partials = nullptr;
    MatchOut       mo;
    SingleJob      sj;
    // the matcher is active at iteration run_from_iteration
    if (int r = single_job_setup(ws, from, to, pose6, D.run_from_iteration, sj, [&](Carver& c) {
            d_partials = c.take<double>((size_t)max_chunks * kNumMoments);
            mo.paired = c.take<uint8_t>(n);
            mo.nn_idx = c.take<uint32_t>(n * k);
            mo.nn_cnt = c.take<uint32_t>(n);
            mo.centroid = c.take<double>(n * 3);
            mo.normal = c.take<double>(n * 3);
        }))
        return r;
    B2_CUDA_TRY(cudaMemsetAsync(mo.paired, 0, n, s));
    B2_CUDA_TRY(cudaMemsetAsync(mo.nn_cnt, 0, n * sizeof(uint32_t), s));
    B2_CUDA_TRY(cudaMemsetAsync(mo.nn_idx, 0xFF, n * k * sizeof(uint32_t), s));
    B2_CUDA_TRY(cudaMemsetAsync(mo.centroid, 0, n * 3 * sizeof(double), s));
    B2_CUDA_TRY(cudaMemsetAsync(mo.normal, 0, n * 3 * sizeof(double), s));
    launch_match<true>(ws, D.knn, dim3(max_chunks, 1), sj.d_clouds, sj.d_jobs, sj.bin, d_partials,
                       max_chunks, D, mo);
    B2_CUDA_TRY(cudaGetLastError());
    std::vector<double> part((size_t)max_chunks * kNumMoments);
    B2_CUDA_TRY(cudaMemcpyAsync(part.data(), d_partials, part.size() * sizeof(double),
                                cudaMemcpyDeviceToHost, s));
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
    {
        double cnt = 0;
        for (uint32_t c = 0; c < max_chunks; c++) cnt += part[(size_t)c * kNumMoments + 73];
        *n_pairings = (uint32_t)(cnt + 0.5);
    }
    return B200ICP_OK;
}

}  // namespace b2
