// runtime.cuh -- host-side runtime of the library: context (the ICP object),
// per-call workspaces (stream + scratch, pooled for re-entrancy, see the
// threading note in include/b200icp.h), error reporting and profiling.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200icp.h"
#include "device_types.cuh"

namespace b2
{
void        set_error(const char* fmt, ...);
const char* get_error();

#define B2_CUDA_TRY(expr)                                                                    \
    do                                                                                       \
    {                                                                                        \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
        {                                                                                    \
            ::b2::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                                 __FILE__, __LINE__);                                        \
            return B200ICP_ERR_CUDA;                                                         \
        }                                                                                    \
    } while (0)

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// carve typed arrays out of one scratch allocation: first pass with base=null
// to size it, second pass with the real base
struct Carver
{
    char*  base;
    size_t off = 0;
    explicit Carver(void* b) : base((char*)b) {}
    template <typename T>
    T* take(size_t count)
    {
        T* p = base ? (T*)(base + off) : nullptr;
        off += align_up(count * sizeof(T));
        return p;
    }
};

// what is baked into a cached CUDA graph of a single registration's first batch
struct AlignGraphKey
{
    void*    d_scratch;
    void*    h_pinned;
    uint64_t max_points, total_queries;
    uint32_t n_views, fit_ctas, first, k;
    uint64_t params_hash;
    bool operator==(const AlignGraphKey& o) const
    {
        return d_scratch == o.d_scratch && h_pinned == o.h_pinned && max_points == o.max_points &&
               total_queries == o.total_queries && n_views == o.n_views && fit_ctas == o.fit_ctas &&
               first == o.first && k == o.k && params_hash == o.params_hash;
    }
};

struct Workspace
{
    int          device = 0;
    cudaStream_t stream = nullptr;
    void*        d_scratch = nullptr;
    size_t       d_bytes = 0;
    void*        h_pinned = nullptr;
    size_t       h_bytes = 0;
    cudaEvent_t  ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> prof_ev;  // pairs
    uint64_t     launches = 0;
    uint32_t*    d_flag = nullptr;  // 256 B of device memory / pinned memory for small status words
    uint32_t*    h_flag = nullptr;
    bool         upload = false;    // belongs to the upload pool

    struct AlignGraph
    {
        AlignGraphKey   key;
        cudaGraphExec_t exec;
        uint64_t        launches;  // kernels one replay launches
    };
    std::vector<AlignGraph> align_graphs;
    bool                    graph_failed = false;
    cudaGraphExec_t find_align_graph(const AlignGraphKey& k) const
    {
        for (const auto& g : align_graphs)
            if (g.key == k) return g.exec;
        return nullptr;
    }
    uint64_t align_graph_launches(const AlignGraphKey& k) const
    {
        for (const auto& g : align_graphs)
            if (g.key == k) return g.launches;
        return 0;
    }
    void store_align_graph(const AlignGraphKey& k, cudaGraphExec_t e, uint64_t launches)
    {
        if (align_graphs.size() >= 8)
        {  // oldest out
            cudaGraphExecDestroy(align_graphs.front().exec);
            align_graphs.erase(align_graphs.begin());
        }
        align_graphs.push_back({k, e, launches});
    }
    void drop_align_graphs()
    {
        for (auto& g : align_graphs) cudaGraphExecDestroy(g.exec);
        align_graphs.clear();
    }

    int  init(int dev);
    void destroy();
    int  reserve_device(size_t bytes);
    int  reserve_pinned(size_t bytes);
    int  reserve_prof_events(size_t pairs);
};

}  // namespace b2

struct b200icp
{
    int                                device = 0;
    b200icp_params_t                   P;
    b2::IcpDevParams              D;
    std::mutex                         mtx;
    std::vector<b2::Workspace*>   free_ws;
    std::vector<b2::Workspace*>   free_upload_ws;  // their streams carry cloud uploads / index builds only, so that
                                                   // the next scan is indexed while a registration is running
    std::vector<b2::Workspace*>   all_ws;
    // Cloud slabs and their "ready" events are recycled: a scan's cloud lives for two registrations, and a
    // device allocation per scan costs more host time than the upload itself
    struct Slab
    {
        void*  p;
        size_t bytes;
    };
    std::vector<Slab>        slab_cache;
    size_t                   slab_cache_bytes = 0;
    std::vector<cudaEvent_t> ready_events;
    void*                    take_slab(size_t need, size_t* got, cudaStream_t stream);  // null on failure
    void                     give_slab(void* p, size_t bytes, cudaStream_t stream);
    cudaEvent_t              take_ready_event();
    void                     give_ready_event(cudaEvent_t e);
    bool                               profile_on = false;
    b200icp_profile_t                  prof;
    int                                sm_count = 148;
    std::atomic<int>                   expected_runs{0};  // matcher runs of the last single registration
    // index-build timings are resolved lazily (profile_get / reset) so that
    // profiling adds no host synchronisation to the upload path
    struct PendingIndexTime
    {
        cudaEvent_t e0, e1;
        uint64_t    points;
    };
    std::vector<PendingIndexTime> pending_index;
    std::vector<cudaEvent_t>      event_pool;
    void                          drain_pending();  // call with mtx held

    b2::Workspace* acquire(bool upload = false);
    void                release(b2::Workspace* ws);
};

struct b200icp_cloud
{
    b200icp* ctx = nullptr;
    size_t   n = 0;
    void*    slab = nullptr;  // one allocation holding everything below
    size_t   slab_bytes = 0;
    float *  dx = nullptr, *dy = nullptr, *dz = nullptr;
    float4*  pts = nullptr;
    float4*  gbox = nullptr;
    uint32_t* rank = nullptr;
    uint32_t* hkeys = nullptr;
    uint4*    hrecs = nullptr;
    uint2*    hrange = nullptr;
    uint32_t* fine_start = nullptr;
    uint32_t* item_first = nullptr;
    b2::GridDev* grid = nullptr;
    uint32_t* bbox_enc = nullptr;  // 6 order-encoded floats
    uint32_t  hcap = 0, hshift = 0;
    float     cell_req = 0;
    bool      indexed = true;  // false: coordinates only (b200icp_cloud_upload_raw), not usable as a search side
    cudaEvent_t ready = nullptr;

    b2::CloudView view() const
    {
        b2::CloudView v;
        v.pts = pts, v.gbox = gbox, v.rank = rank, v.grid = grid, v.hkeys = hkeys, v.hrecs = hrecs, v.hrange = hrange;
        v.fine_start = fine_start;
        v.item_first = item_first;
        v.hshift = hshift, v.hmask = hcap - 1, v.n = (uint32_t)n;
        return v;
    }
};

namespace b2
{
// RAII lease of a workspace
struct Lease
{
    ::b200icp* ctx;
    Workspace* ws;
    explicit Lease(::b200icp* c, bool upload = false) : ctx(c), ws(c->acquire(upload)) {}
    ~Lease()
    {
        if (ws) ctx->release(ws);
    }
};

int cloud_alloc(::b200icp* ctx, Workspace* ws, size_t n, float search_radius, b200icp_cloud** out,
                float min_cell = 0.f, bool coords_only = false);
int cloud_build_index(::b200icp* ctx, Workspace* ws, b200icp_cloud* c);

int run_align_batch(::b200icp* ctx, size_t n, const b200icp_cloud* const* from,
                    const b200icp_cloud* const* to, const double* guesses, const b200icp_call_params_t* call,
                    b200icp_result_t* out);
int run_knn(::b200icp* ctx, const b200icp_cloud* ref, const b200icp_cloud* q, const double* pose6,
            uint32_t k, float max_dist, uint32_t* idx_out, float* d2_out);
int run_knn_keys(::b200icp* ctx, const b200icp_cloud* ref, const b200icp_cloud* q, const double* pose6,
                 uint32_t k, float max_dist, const uint32_t* d_index_map, uint64_t* d_keys_out);
int run_knn_keys_scatter(::b200icp* ctx, const b200icp_cloud* ref, const b200icp_cloud* q, const double* pose6,
                         uint32_t k, float max_dist, const uint32_t* d_index_map, uint64_t* const* d_gather,
                         uint32_t world, uint32_t rank, int atomic_min);
int run_knn_exchange(::b200icp* ctx, const b200icp_cloud* ref, const b200icp_cloud* q, const double* pose6,
                     uint32_t k, float max_dist, const uint32_t* d_index_map, uint64_t* const* d_bases,
                     uint32_t world, uint32_t rank, uint64_t* epoch_io, uint64_t* d_out);
int run_peer_barrier(::b200icp* ctx, uint64_t* const* d_flags, uint32_t world, uint32_t rank, uint64_t epoch);
int run_fill_no_key(::b200icp* ctx, uint64_t* d_keys, size_t n);
int run_merge_keys(::b200icp* ctx, const uint64_t* d_parts, uint32_t parts, size_t part_stride, size_t nq,
                   uint32_t k, uint64_t* d_out);
int run_match(::b200icp* ctx, const b200icp_cloud* from, const b200icp_cloud* to,
              const double* pose6, uint8_t* paired, uint32_t* nn_idx, uint32_t* nn_cnt,
              double* centroid, double* normal, uint32_t* n_pairings);
int run_voxel(::b200icp* ctx, const b200icp_cloud* in, float resolution, int use_average,
              float search_radius, b200icp_cloud** out, uint32_t* keep_idx);

int run_edges_planes(::b200icp* ctx, const b200icp_cloud* in, const b200icp_edges_planes_params_t* prm,
                     float search_radius, b200icp_cloud** out3, uint8_t* layer_out, uint32_t* n_voxels_out);

// sharded maps over the GPUs of one box (sharded.inl); the library owns the NCCL communicator
int  run_comm_unique_id(unsigned char* id_out);
int  run_comm_create(::b200icp* ctx, const unsigned char* id_bytes, int world, int rank, b200icp_comm** out);
void run_comm_destroy(b200icp_comm* c);
int  run_sharded_map_create(b200icp_comm* comm, const float* x, const float* y, const float* z, size_t n, float cell,
                            int interleaved, float search_radius, b200icp_sharded_map** out);
void run_sharded_map_destroy(b200icp_sharded_map* m);
size_t run_sharded_map_local_size(const b200icp_sharded_map* m);
int  run_sharded_knn_keys(b200icp_sharded_map* m, const b200icp_cloud* q, const double* pose6, uint32_t k,
                          float max_dist, uint64_t* d_keys_out);
int  run_sharded_align(b200icp_sharded_map* m, const b200icp_cloud* to, const double* guess6,
                       const b200icp_call_params_t* call, b200icp_result_t* out);

void make_dev_params(const b200icp_params_t& P, IcpDevParams& D);
}  // namespace b2
