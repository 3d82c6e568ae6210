// sharded.inl -- a very large `from` cloud split by spatial cell over the GPUs of one box, and registrations
// against it (BASELINE config 5).  Included at the end of align.cu (same kernels).
//
// One process per GPU; the library owns the NCCL communicator (b200icp_comm_*), no Python in the data path.
// Every rank holds the search index of ITS cells only -- that is what costs time and memory to build and to
// walk -- and a plain copy of all map coordinates (16 B per point: 320 MB for 20 M points), because the plane
// fit of a query needs the coordinates of neighbours that other ranks found.
//
// One outer iteration of a registration against the sharded map:
//   1. every rank searches ALL local points against its shard            -> partial keys [sorted position][k]
//   2. reduce-scatter: the lists of the queries of slice s go to rank s  (grouped ncclSend / ncclRecv)
//   3. rank s merges the `world` lists of its slice (k smallest keys)    -> neighbour rows of its slice
//   4. plane fit + moments of its slice, summed per chunk and per group  (fit kernel, tail_mode 2)
//   5. all-gather of the group partials (1.5 KB each)                    (ncclAllGather)
//   6. every rank sums ALL group partials in group order and solves      (sharded_solve_kernel)
// Slices are whole groups of chunks, so the chunk partials, the group partials and the order of the final sum are
// exactly those of the single-GPU registration (align.cu, "chunk partials"): the pose, the iteration count, the
// covariance are BIT-IDENTICAL to b200icp_align against the unsharded map, on every rank.
// (<dlfcn.h> and <nccl.h> are included at the top of align.cu, outside the namespace)

// ---- NCCL, bound at run time: the process may already hold a libnccl (PyTorch ships its own) ---------------
struct NcclApi
{
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool        ok = false;
    std::string why;
};

static const NcclApi& nccl_api()
{
    static const NcclApi api = [] {
        NcclApi a;
        void*   h = nullptr;
        for (const char* name : {"libnccl.so.2", "libnccl.so"})
        {
            h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h)
        {
            a.why = std::string("libnccl not found: ") + (dlerror() ? dlerror() : "");
            return a;
        }
        auto sym = [&](const char* n) { return dlsym(h, n); };
        a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
        a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
        a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
        a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
        a.Send = (decltype(a.Send))sym("ncclSend");
        a.Recv = (decltype(a.Recv))sym("ncclRecv");
        a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
        a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
        a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.AllReduce && a.Send && a.Recv &&
               a.GroupStart && a.GroupEnd && a.GetErrorString;
        if (!a.ok) a.why = "libnccl lacks a required entry point";
        return a;
    }();
    return api;
}

#define B2_NCCL_TRY(expr)                                                                                  \
    do                                                                                                     \
    {                                                                                                      \
        ncclResult_t _r = (expr);                                                                          \
        if (_r != ncclSuccess)                                                                             \
        {                                                                                                  \
            set_error("%s failed: %s (%s:%d)", #expr, nccl_api().GetErrorString(_r), __FILE__, __LINE__);  \
            return B200ICP_ERR_CUDA;                                                                       \
        }                                                                                                  \
    } while (0)

}  // namespace b2

struct b200icp_comm
{
    ::b200icp* ctx = nullptr;
    ncclComm_t comm = nullptr;
    int        world = 1, rank = 0;
};

struct b200icp_sharded_map
{
    b200icp_comm*  comm = nullptr;
    b200icp_cloud* shard = nullptr;       // this rank's cells, indexed
    float4*        d_all = nullptr;       // every map point by global index (x, y, z, bitcast index)
    uint32_t*      d_index_map = nullptr; // shard-local index -> global index (increasing)
    b2::GridDev*   d_all_grid = nullptr;  // a grid record that only says "n_all points" (the fit's emptiness test)
    size_t         n_all = 0, n_shard = 0;
    float          radius = 0;
};

namespace b2
{
// ---- kernels of the sharded path ------------------------------------------------------------------------
__global__ void all_points_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z,
                                  size_t n, float4* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float4(x[i], y[i], z[i], __uint_as_float((uint32_t)i));
}

// merged keys of a slice -> neighbour rows as the fit stage reads them (global index = position in d_all)
__global__ void keys_to_rows_kernel(const uint64_t* __restrict__ keys, size_t n, uint32_t* __restrict__ rows)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = keys[i];
    rows[i] = (k == B200ICP_NO_KEY) ? kInvalid : (uint32_t)(k & 0xFFFFFFFFull);
}

// every rank, after the all-gather of the group partials: the job's moments in group order, then the solver
__global__ void __launch_bounds__(32)
    sharded_solve_kernel(const CloudView* __restrict__ clouds, JobDev* __restrict__ jobs, FitBuffers fb,
                         uint32_t chunk_items, IcpDevParams P, uint32_t* __restrict__ n_active)
{
    JobDev& J = jobs[0];
    if (J.status != 0) return;
    __shared__ SolveSmem ss;
    const int            lane = threadIdx.x;
    const CloudView      cvL = clouds[J.to_cloud];
    const bool           any_global = clouds[2].grid->n_valid > 0;  // as the fit kernel tests its global side
    const uint32_t       n_items = (matcher_active(P, J.iter) && any_global) ? cvL.grid->n_items : 0u;
    const uint32_t       n_chunks = (n_items + chunk_items - 1) / chunk_items;
    const uint32_t       n_groups = (n_chunks + kFitGroup - 1) / kFitGroup;
    int                  idx[6];
    frag_slots(lane, idx);
    double v[6] = {0, 0, 0, 0, 0, 0};
    if (n_groups) sum_records(fb.gpartials + (size_t)J.group_base * kNumMoments, n_groups, idx, v);
#pragma unroll
    for (int i = 0; i < 6; i++) J.Mprev[idx[i]] = J.M[idx[i]], ss.S[idx[i]] = v[i], J.M[idx[i]] = v[i];
    __syncwarp();
    solve_job_warp(J, ss, P, n_active);
}

// QualityEvaluator_PairedRatio on the merged nearest-neighbour keys (A.8)
__global__ void count_hits_kernel(const uint64_t* __restrict__ keys, size_t n, float thr2, JobDev* __restrict__ jobs)
{
    const size_t   i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool     hit = i < n && keys[i] != B200ICP_NO_KEY && key_d2(keys[i]) < thr2;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&jobs[0].quality_count, (uint32_t)__popc(m));
}

__global__ void mark_evaluating_kernel(JobDev* jobs) { jobs[0].quality_count = 0; }

// ---- host side ------------------------------------------------------------------------------------------
// owner rank of every map point: coarse (x, y) cells along a Morton curve, dealt round-robin ("interleaved":
// every rank holds 1/world of every neighbourhood) or cut into `world` runs of equal point count
static void partition_by_cell(const float* x, const float* y, size_t n, int world, float cell, bool interleaved,
                              std::vector<int>& owner)
{
    owner.assign(n, 0);
    if (n == 0 || world <= 1) return;
    float lox = INFINITY, loy = INFINITY;
    for (size_t i = 0; i < n; i++)
        if (std::isfinite(x[i]) && std::isfinite(y[i])) lox = std::min(lox, x[i]), loy = std::min(loy, y[i]);
    if (!std::isfinite(lox)) lox = loy = 0.f;
    auto spread16 = [](uint64_t v) {
        v &= 0xFFFF;
        v = (v | (v << 8)) & 0x00FF00FFull;
        v = (v | (v << 4)) & 0x0F0F0F0Full;
        v = (v | (v << 2)) & 0x33333333ull;
        v = (v | (v << 1)) & 0x55555555ull;
        return v;
    };
    std::vector<uint64_t> code(n);
    for (size_t i = 0; i < n; i++)
    {
        long cx = 0, cy = 0;
        if (std::isfinite(x[i]) && std::isfinite(y[i]))
        {
            cx = (long)std::floor((x[i] - lox) / cell), cy = (long)std::floor((y[i] - loy) / cell);
            cx = std::min(std::max(cx, 0l), 0xFFFFl), cy = std::min(std::max(cy, 0l), 0xFFFFl);
        }
        code[i] = spread16((uint64_t)cx) | (spread16((uint64_t)cy) << 1);
    }
    std::map<uint64_t, size_t> cells;  // code -> points in the cell, ordered along the curve
    for (size_t i = 0; i < n; i++) cells[code[i]]++;
    std::map<uint64_t, int> owner_of;
    size_t                  ord = 0, before = 0;
    for (auto& kv : cells)
    {
        owner_of[kv.first] = interleaved ? (int)(ord % (size_t)world)
                                         : (int)std::min<size_t>((size_t)world - 1, before * (size_t)world / n);
        ord++, before += kv.second;
    }
    for (size_t i = 0; i < n; i++) owner[i] = owner_of[code[i]];
}

int run_comm_unique_id(unsigned char* id_out)
{
    const NcclApi& N = nccl_api();
    if (!N.ok)
    {
        set_error("%s", N.why.c_str());
        return B200ICP_ERR_UNSUPPORTED;
    }
    ncclUniqueId id;
    B2_NCCL_TRY(N.GetUniqueId(&id));
    static_assert(sizeof(id) == B200ICP_COMM_ID_BYTES, "unique id size");
    memcpy(id_out, &id, sizeof(id));
    return B200ICP_OK;
}

int run_comm_create(::b200icp* ctx, const unsigned char* id_bytes, int world, int rank, b200icp_comm** out)
{
    const NcclApi& N = nccl_api();
    if (!N.ok)
    {
        set_error("%s", N.why.c_str());
        return B200ICP_ERR_UNSUPPORTED;
    }
    B2_CUDA_TRY(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    auto* c = new b200icp_comm();
    c->ctx = ctx, c->world = world, c->rank = rank;
    ncclResult_t r = N.CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess)
    {
        set_error("ncclCommInitRank failed: %s", N.GetErrorString(r));
        delete c;
        return B200ICP_ERR_CUDA;
    }
    *out = c;
    return B200ICP_OK;
}

void run_comm_destroy(b200icp_comm* c)
{
    if (!c) return;
    if (c->comm) nccl_api().CommDestroy(c->comm);
    delete c;
}

size_t run_sharded_map_local_size(const b200icp_sharded_map* m) { return m ? m->n_shard : 0; }

void run_sharded_map_destroy(b200icp_sharded_map* m)
{
    if (!m) return;
    cudaSetDevice(m->comm->ctx->device);
    if (m->shard) b200icp_cloud_free(m->shard);
    if (m->d_all) cudaFree(m->d_all);
    if (m->d_index_map) cudaFree(m->d_index_map);
    if (m->d_all_grid) cudaFree(m->d_all_grid);
    delete m;
}

int run_sharded_map_create(b200icp_comm* comm, const float* x, const float* y, const float* z, size_t n, float cell,
                           int interleaved, float search_radius, b200icp_sharded_map** out)
{
    ::b200icp* ctx = comm->ctx;
    if (n >= 0xFFFFFFFEull)
    {
        set_error("map too large for 32-bit global indices: %zu points", n);
        return B200ICP_ERR_BAD_ARG;
    }
    B2_CUDA_TRY(cudaSetDevice(ctx->device));
    std::vector<int> owner;
    partition_by_cell(x, y, n, comm->world, cell > 0 ? cell : 4.0f, interleaved != 0, owner);
    std::vector<uint32_t> mine;
    for (size_t i = 0; i < n; i++)
        if (owner[i] == comm->rank) mine.push_back((uint32_t)i);
    std::vector<float> sx(mine.size()), sy(mine.size()), sz(mine.size());
    for (size_t j = 0; j < mine.size(); j++) sx[j] = x[mine[j]], sy[j] = y[mine[j]], sz[j] = z[mine[j]];

    auto* m = new b200icp_sharded_map();
    m->comm = comm, m->n_all = n, m->n_shard = mine.size();
    m->radius = search_radius > 0 ? search_radius : (float)ctx->P.distance_threshold;
    auto fail = [&](int rc) {
        run_sharded_map_destroy(m);
        return rc;
    };
    if (int r = b200icp_cloud_upload(ctx, sx.data(), sy.data(), sz.data(), mine.size(), m->radius, &m->shard)) return fail(r);
    // all coordinates by global index
    Lease L(ctx, true);
    if (!L.ws) return fail(B200ICP_ERR_CUDA);
    cudaStream_t s = L.ws->stream;
    const size_t nn = n ? n : 1;
    float *      dx = nullptr, *dy = nullptr, *dz = nullptr;
    cudaError_t  e = cudaMalloc(&m->d_all, nn * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc(&m->d_index_map, (mine.size() ? mine.size() : 1) * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&m->d_all_grid, sizeof(GridDev));
    if (e == cudaSuccess) e = cudaMalloc(&dx, 3 * nn * sizeof(float));
    if (e != cudaSuccess)
    {
        set_error("device allocation for the sharded map failed: %s", cudaGetErrorString(e));
        if (dx) cudaFree(dx);
        return fail(B200ICP_ERR_NOMEM);
    }
    dy = dx + nn, dz = dy + nn;
    GridDev g;
    memset(&g, 0, sizeof(g));
    g.n_valid = (uint32_t)std::min<size_t>(n, 0xFFFFFFFFull);
    if (n)
    {
        e = cudaMemcpyAsync(dx, x, n * sizeof(float), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dy, y, n * sizeof(float), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dz, z, n * sizeof(float), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess)
            all_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dx, dy, dz, n, m->d_all);
        L.ws->launches++;
    }
    if (e == cudaSuccess && !mine.empty())
        e = cudaMemcpyAsync(m->d_index_map, mine.data(), mine.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(m->d_all_grid, &g, sizeof(g), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(dx);
    if (e != cudaSuccess)
    {
        set_error("upload of the sharded map failed: %s", cudaGetErrorString(e));
        return fail(B200ICP_ERR_CUDA);
    }
    *out = m;
    return B200ICP_OK;
}

template <class Epi>
static void launch_search_any_k(const ::b200icp* ctx, Workspace* ws, uint32_t k, size_t nq, const CloudView* d_clouds,
                                JobDev* d_jobs, const IcpDevParams& D, float cap_d2, int gate, const Epi& w)
{
    if (k == 1)
        launch_search_k<1>(ctx, ws, nq, 1, d_clouds, d_jobs, D, cap_d2, gate, w);
    else if (k <= 4)
        launch_search_k<4>(ctx, ws, nq, 1, d_clouds, d_jobs, D, cap_d2, gate, w);
    else if (k <= 6)
        launch_search_k<6>(ctx, ws, nq, 1, d_clouds, d_jobs, D, cap_d2, gate, w);
    else
        launch_search_k<8>(ctx, ws, nq, 1, d_clouds, d_jobs, D, cap_d2, gate, w);
}

static void launch_merge_keys(cudaStream_t s, uint32_t k, const uint64_t* parts, uint32_t nparts, size_t stride,
                              size_t nq, uint64_t* out)
{
    if (nq == 0) return;
    const int blocks = (int)((nq + 255) / 256);
    if (k == 1)
        merge_keys_kernel<1><<<blocks, 256, 0, s>>>(parts, nparts, stride, nq, k, out);
    else if (k <= 4)
        merge_keys_kernel<4><<<blocks, 256, 0, s>>>(parts, nparts, stride, nq, k, out);
    else if (k <= 6)
        merge_keys_kernel<6><<<blocks, 256, 0, s>>>(parts, nparts, stride, nq, k, out);
    else
        merge_keys_kernel<8><<<blocks, 256, 0, s>>>(parts, nparts, stride, nq, k, out);
}

// the k nearest map points of every query, merged over the ranks: the same keys on every rank
int run_sharded_knn_keys(b200icp_sharded_map* m, const b200icp_cloud* q, const double* pose6, uint32_t k,
                         float max_dist, uint64_t* d_keys_out)
{
    ::b200icp*     ctx = m->comm->ctx;
    const NcclApi& N = nccl_api();
    if (k < 1 || k > B200ICP_MAX_KNN || !(max_dist > 0) || !std::isfinite(max_dist))
    {
        set_error("k=%u outside [1,%d] or max_dist not positive and finite", k, B200ICP_MAX_KNN);
        return B200ICP_ERR_BAD_ARG;
    }
    const size_t nq = q->n;
    if (nq == 0) return B200ICP_OK;
    const int world = m->comm->world;
    Lease     L(ctx);
    if (!L.ws) return B200ICP_ERR_CUDA;
    Workspace*   ws = L.ws;
    cudaStream_t s = ws->stream;
    uint64_t*    d_parts = nullptr;
    SingleJob    sj;
    if (int r = single_job_setup(ws, m->shard, q, pose6, 0, sj, [&](Carver& c) {
            if (k > 1 && world > 1) d_parts = c.take<uint64_t>((size_t)world * nq * k);
        }))
        return r;
    fill_u64_kernel<<<(int)((nq * k + 255) / 256), 256, 0, s>>>(d_keys_out, nq * k, B200ICP_NO_KEY);
    ws->launches++;
    const KeyWriter w = {d_keys_out, m->d_index_map, k, 0u};
    launch_search_any_k(ctx, ws, k, nq, sj.d_clouds, sj.d_jobs, ctx->D, max_dist * max_dist, 0, w);
    B2_CUDA_TRY(cudaGetLastError());
    if (world > 1)
    {
        if (k == 1)  // the integer order of a key is the tie rule: a plain minimum
            B2_NCCL_TRY(N.AllReduce(d_keys_out, d_keys_out, nq, ncclUint64, ncclMin, m->comm->comm, s));
        else
        {
            B2_NCCL_TRY(N.AllGather(d_keys_out, d_parts, nq * k, ncclUint64, m->comm->comm, s));
            launch_merge_keys(s, k, d_parts, (uint32_t)world, nq * k, nq, d_keys_out);
            ws->launches++;
        }
    }
    B2_CUDA_TRY(cudaStreamSynchronize(s));
    B2_CUDA_TRY(cudaGetLastError());
    return B200ICP_OK;
}

int run_sharded_align(b200icp_sharded_map* m, const b200icp_cloud* to, const double* guess6,
                      const b200icp_call_params_t* call, b200icp_result_t* out)
{
    ::b200icp*         ctx = m->comm->ctx;
    const NcclApi&     N = nccl_api();
    const IcpDevParams D = merged_params(ctx, call);
    if (int r = check_supported(D)) return r;
    if (D.solver_kind != B200ICP_SOLVER_GAUSS_NEWTON)
    {
        set_error("the sharded registration runs the Gauss-Newton solver only");
        return B200ICP_ERR_UNSUPPORTED;
    }
    const int    world = m->comm->world, rank = m->comm->rank;
    ncclComm_t   comm = m->comm->comm;
    const int    K = matcher_k(D);
    const size_t nq = to->n;
    Lease        L(ctx);
    if (!L.ws) return B200ICP_ERR_CUDA;
    Workspace*   ws = L.ws;
    cudaStream_t s = ws->stream;
    if (int r = wait_cloud(ws, m->shard)) return r;
    if (int r = wait_cloud(ws, to)) return r;

    // the reduction tree of the single-GPU registration (fit_chunk_items(n, 1) == 2); slices = whole groups
    const uint32_t chunk_items = fit_chunk_items(nq, 1);
    const size_t   items = (nq + kItem - 1) / kItem;
    const size_t   chunks = (items + chunk_items - 1) / chunk_items;
    const size_t   groups = (chunks + kFitGroup - 1) / kFitGroup;
    const size_t   gpr = (groups + world - 1) / world;                 // groups per rank
    const size_t   group_pts = (size_t)kFitGroup * chunk_items * kItem;  // sorted positions per group
    const size_t   slice_cap = gpr * group_pts;                        // positions per slice (upper bound)
    auto slice_lo = [&](int r) { return std::min(nq, (size_t)r * slice_cap); };
    auto slice_cnt = [&](int r) { return std::min(nq, (size_t)(r + 1) * slice_cap) - slice_lo(r); };
    const size_t my_lo = slice_lo(rank), my_cnt = slice_cnt(rank);

    CloudView* d_clouds = nullptr;
    JobDev*    d_jobs = nullptr;
    uint32_t * d_active = nullptr, *d_nn = nullptr;
    uint64_t * d_keys = nullptr, *d_gather = nullptr, *d_merged = nullptr;
    FitBuffers fb = {nullptr, nullptr, nullptr};
    auto layout = [&](Carver& c) {
        d_clouds = c.take<CloudView>(4);
        d_jobs = c.take<JobDev>(1);
        d_active = c.take<uint32_t>(4);
        fb.tickets = c.take<uint32_t>(world * gpr ? world * gpr : 1);
        fb.gpartials = c.take<double>((world * gpr ? world * gpr : 1) * (size_t)kNumMoments);
        fb.partials = c.take<double>((chunks ? chunks : 1) * (size_t)kNumMoments);
        d_nn = c.take<uint32_t>((nq ? nq : 1) * (size_t)K);
        d_keys = c.take<uint64_t>((nq ? nq : 1) * (size_t)K);
        d_gather = c.take<uint64_t>((size_t)world * (slice_cap ? slice_cap : 1) * K);
        d_merged = c.take<uint64_t>((slice_cap ? slice_cap : 1) * (size_t)K);
    };
    Carver sz(nullptr);
    layout(sz);
    if (int r = ws->reserve_device(sz.off)) return r;
    Carver real(ws->d_scratch);
    layout(real);
    const size_t off_job = align_up(4 * sizeof(CloudView));
    if (int r = ws->reserve_pinned(off_job + align_up(sizeof(JobDev)) + 64)) return r;
    CloudView* hv = (CloudView*)ws->h_pinned;
    JobDev*    hj = (JobDev*)((char*)ws->h_pinned + off_job);
    uint32_t*  hflag = (uint32_t*)((char*)hj + align_up(sizeof(JobDev)));
    // cloud table: the job names its clouds 0 (from) and 1 (to).  The search reads the table at its start:
    // 0 = this rank's shard, 1 = the local cloud; the fit reads it from entry 2 on: 0 = EVERY map point by global
    // index (the neighbour rows of the merged lists hold global indices), 1 = the local cloud again
    hv[0] = m->shard->view(), hv[1] = to->view(), hv[3] = to->view();
    memset(&hv[2], 0, sizeof(CloudView));
    hv[2].pts = m->d_all, hv[2].grid = m->d_all_grid, hv[2].n = (uint32_t)m->n_all;
    memset(hj, 0, sizeof(JobDev));
    Pose T;
    pose_from_ypr(guess6, T);
    memcpy(hj->R, T.R, sizeof(T.R)), memcpy(hj->t, T.t, sizeof(T.t));
    memcpy(hj->Rprev, T.R, sizeof(T.R)), memcpy(hj->tprev, T.t, sizeof(T.t));
    hj->from_cloud = 0, hj->to_cloud = 1;
    *hflag = 1;
    B2_CUDA_TRY(cudaMemcpyAsync(d_clouds, hv, 4 * sizeof(CloudView), cudaMemcpyHostToDevice, s));
    B2_CUDA_TRY(cudaMemcpyAsync(d_jobs, hj, sizeof(JobDev), cudaMemcpyHostToDevice, s));
    B2_CUDA_TRY(cudaMemcpyAsync(d_active, hflag, sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    B2_CUDA_TRY(cudaMemsetAsync(fb.tickets, 0, (world * gpr ? world * gpr : 1) * sizeof(uint32_t), s));

    const uint32_t G = fit_ctas_per_job(ctx, std::max<size_t>(my_cnt, 1), 1, chunk_items);
    const MatchOut no_out = {nullptr, nullptr, nullptr, nullptr, nullptr};
    const uint32_t chunk_begin = (uint32_t)((size_t)rank * gpr * kFitGroup);
    const uint32_t chunk_end = (uint32_t)std::min<size_t>(chunks, (size_t)(rank + 1) * gpr * kFitGroup);
    const size_t   nkeys = (nq ? nq : 1) * (size_t)K;

    bool finished = false;
    for (uint32_t it = 0; it < D.max_iterations && !finished; it++)
    {
        // 1. this rank's shard, all queries; rows by sorted position, global indices
        fill_u64_kernel<<<(int)((nkeys + 255) / 256), 256, 0, s>>>(d_keys, nkeys, B200ICP_NO_KEY);
        ws->launches++;
        const KeyWriter w = {d_keys, m->d_index_map, (uint32_t)K, 1u};
        launch_search_any_k(ctx, ws, (uint32_t)K, std::max<size_t>(nq, 1), d_clouds, d_jobs, D, D.thr2, 1, w);
        // 2. reduce-scatter: the rows of slice r go to rank r
        if (world > 1)
        {
            B2_NCCL_TRY(N.GroupStart());
            for (int r = 0; r < world; r++)
            {
                if (slice_cnt(r))
                    B2_NCCL_TRY(N.Send(d_keys + slice_lo(r) * K, slice_cnt(r) * K, ncclUint64, r, comm, s));
                if (my_cnt) B2_NCCL_TRY(N.Recv(d_gather + (size_t)r * slice_cap * K, my_cnt * K, ncclUint64, r, comm, s));
            }
            B2_NCCL_TRY(N.GroupEnd());
            // 3. k smallest of the `world` lists of every query of the slice
            launch_merge_keys(s, (uint32_t)K, d_gather, (uint32_t)world, slice_cap * K, my_cnt, d_merged);
            ws->launches++;
        }
        const uint64_t* merged = (world > 1) ? d_merged : d_keys + my_lo * K;
        if (my_cnt)
        {
            keys_to_rows_kernel<<<(int)((my_cnt * K + 255) / 256), 256, 0, s>>>(merged, my_cnt * K, d_nn + my_lo * K);
            ws->launches++;
        }
        // 4. plane fit + moments of the slice: chunk partials, group partials (no job-level sum: tail_mode 2)
        launch_fit<false>(ws, D, dim3(G, 1), d_clouds + 2, d_jobs, d_nn, fb, chunk_items, 2, no_out, nullptr, d_active,
                          chunk_begin, chunk_end);
        // 5. every rank's group partials to every rank, in global group order
        if (world > 1)
            B2_NCCL_TRY(N.AllGather(fb.gpartials + (size_t)rank * gpr * kNumMoments, fb.gpartials, gpr * kNumMoments,
                                    ncclDouble, comm, s));
        // 6. the same sum and the same solve on every rank
        sharded_solve_kernel<<<1, 32, 0, s>>>(d_clouds, d_jobs, fb, chunk_items, D, d_active);
        ws->launches++;
        B2_CUDA_TRY(cudaMemcpyAsync(hflag, d_active, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        B2_CUDA_TRY(cudaStreamSynchronize(s));
        B2_CUDA_TRY(cudaGetLastError());
        finished = (*hflag == 0);
    }
    // quality: nearest map point of every local point under the final pose, merged with a minimum
    {
        const size_t n1 = nq ? nq : 1;
        fill_u64_kernel<<<(int)((n1 + 255) / 256), 256, 0, s>>>(d_keys, n1, B200ICP_NO_KEY);
        const KeyWriter w = {d_keys, m->d_index_map, 1u, 1u};
        launch_search_any_k(ctx, ws, 1u, n1, d_clouds, d_jobs, D, D.q_thr2, 2, w);
        if (world > 1 && nq) B2_NCCL_TRY(N.AllReduce(d_keys, d_keys, nq, ncclUint64, ncclMin, comm, s));
        mark_evaluating_kernel<<<1, 1, 0, s>>>(d_jobs);
        if (nq) count_hits_kernel<<<(int)((nq + 255) / 256), 256, 0, s>>>(d_keys, nq, D.q_thr2, d_jobs);
        covariance_kernel<<<1, 64, 0, s>>>(d_jobs, D);
        ws->launches += 5;
    }
    B2_CUDA_TRY(cudaMemcpyAsync(hj, d_jobs, sizeof(JobDev), cudaMemcpyDeviceToHost, s));
    B2_CUDA_TRY(cudaStreamSynchronize(s));
    B2_CUDA_TRY(cudaGetLastError());
    const JobDev& J = *hj;
    memset(out, 0, sizeof(*out));
    Pose Tf;
    memcpy(Tf.R, J.R, sizeof(Tf.R)), memcpy(Tf.t, J.t, sizeof(Tf.t));
    pose_to_ypr(Tf, out->pose);
    memcpy(out->R, J.R, sizeof(out->R)), memcpy(out->t, J.t, sizeof(out->t));
    memcpy(out->cov, J.cov, sizeof(out->cov));
    out->quality = nq ? (double)J.quality_count / (double)nq : 0.0;
    out->n_iterations = J.iter;
    out->termination_reason = J.term_reason;
    out->n_pairings = J.n_pairings;
    out->cov_singular = J.cov_singular;
    return B200ICP_OK;
}
