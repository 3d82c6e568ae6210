// capi.cu -- extern "C" entry points declared in include/b200icp.h.
#include <cmath>
#include <algorithm>
#include <thread>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "runtime.cuh"

using namespace b2;

extern "C" const char* b200icp_last_error(void) { return get_error(); }

extern "C" int b200icp_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int b200icp_create(const b200icp_params_t* params, int device, b200icp_t** out)
{
    if (!params || !out)
    {
        set_error("null argument");
        return B200ICP_ERR_BAD_ARG;
    }
    *out = nullptr;
    int ndev = 0;
    B2_CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev)
    {
        set_error("CUDA device %d not available (%d devices): this library has no CPU path", device, ndev);
        return B200ICP_ERR_CUDA;
    }
    B2_CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    B2_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
    {
        set_error("device %d is sm_%d%d; this library carries sm_100a kernels only", device, prop.major,
                  prop.minor);
        return B200ICP_ERR_CUDA;
    }
    // keep freed stream-ordered allocations cached in the pool
    cudaMemPool_t pool;
    B2_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thr = UINT64_MAX;
    B2_CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    {
        // Give the pool a reserve once per device: growing it later (keyframe
        // clouds stay alive in the world model, LidarOdometry.cpp:384-388) maps
        // fresh memory inside cudaMallocAsync, a multi-millisecond stall on the
        // per-scan path.  4 GiB of 180 (clouds of the key-frames, plus the scratch of every workspace, which also
        // comes from the pool); returned to the pool at once.
        static std::mutex       pm;
        static std::vector<int> primed;
        std::lock_guard<std::mutex> lk(pm);
        if (std::find(primed.begin(), primed.end(), device) == primed.end())
        {
            void* p = nullptr;
            if (cudaMallocAsync(&p, (size_t)4 << 30, 0) == cudaSuccess)
            {
                cudaFreeAsync(p, 0);
                cudaStreamSynchronize(0);
            }
            else
                cudaGetLastError();  // a small or busy device: not an error
            primed.push_back(device);
        }
    }
    auto* ctx = new (std::nothrow) b200icp();
    if (!ctx) return B200ICP_ERR_NOMEM;
    ctx->device = device;
    ctx->P = *params;
    ctx->sm_count = prop.multiProcessorCount;
    make_dev_params(ctx->P, ctx->D);
    memset(&ctx->prof, 0, sizeof(ctx->prof));
    // the first workspaces now: a broken device fails here, and the threads that later call into this object
    // at the same time (the reference's pools: one per-scan thread + hw/2 threads for nearby key-frames,
    // LidarOdometry.cpp:94-96) find a stream, events and pinned staging ready instead of creating them -- with
    // device-wide synchronisations -- in the middle of someone else's registration
    // (as many as that pool can have threads: hardware_concurrency / 2 + the per-scan thread + one spare)
    const unsigned hw = std::thread::hardware_concurrency();
    const size_t   n_first = std::min<size_t>(16, std::max<size_t>(4, hw / 2 + 2));
    std::vector<Workspace*> first;
    for (size_t i = 0; i < n_first; i++)
    {
        Workspace* w = ctx->acquire();
        if (!w) break;
        first.push_back(w);
    }
    if (first.empty())
    {
        delete ctx;
        return B200ICP_ERR_CUDA;
    }
    if (Workspace* up = ctx->acquire(true)) ctx->release(up);  // and one for uploads / index builds
    for (auto* w : first) ctx->release(w);
    *out = ctx;
    return B200ICP_OK;
}

extern "C" int b200icp_create_from_yaml(const char* yaml_text, int device, b200icp_t** out)
{
    b200icp_params_t p;
    if (int r = b200icp_params_from_yaml(yaml_text, &p)) return r;
    return b200icp_create(&p, device, out);
}

extern "C" void b200icp_destroy(b200icp_t* icp)
{
    if (!icp) return;
    for (auto* w : icp->all_ws)
    {
        w->destroy();
        delete w;
    }
    for (auto& sl : icp->slab_cache) cudaFree(sl.p);
    for (auto e : icp->ready_events) cudaEventDestroy(e);
    {
        std::lock_guard<std::mutex> lk(icp->mtx);
        icp->drain_pending();
        for (auto e : icp->event_pool) cudaEventDestroy(e);
        icp->event_pool.clear();
    }
    delete icp;
}

extern "C" int b200icp_get_params(const b200icp_t* icp, b200icp_params_t* out)
{
    if (!icp || !out) return B200ICP_ERR_BAD_ARG;
    *out = icp->P;
    return B200ICP_OK;
}

extern "C" int b200icp_device(const b200icp_t* icp) { return icp ? icp->device : -1; }

static int upload_common(b200icp_t* icp, const float* x, const float* y, const float* z, size_t n,
                         float search_radius, cudaMemcpyKind kind, b200icp_cloud_t** out, bool coords_only = false)
{
    if (!icp || !out || (n && (!x || !y || !z)))
    {
        set_error("null argument");
        return B200ICP_ERR_BAD_ARG;
    }
    *out = nullptr;
    Lease L(icp, /*upload=*/true);  // its own stream: overlaps the registration that is running
    if (!L.ws) return B200ICP_ERR_CUDA;
    b200icp_cloud* c = nullptr;
    if (int r = cloud_alloc(icp, L.ws, n, search_radius, &c, 0.f, coords_only)) return r;
    cudaStream_t s = L.ws->stream;
    if (n)
    {
        cudaError_t e = cudaMemcpyAsync(c->dx, x, n * sizeof(float), kind, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->dy, y, n * sizeof(float), kind, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->dz, z, n * sizeof(float), kind, s);
        // the caller's buffers may be reused as soon as we return
        if (e == cudaSuccess) e = cudaEventRecord(L.ws->ev[2], s);
        if (e != cudaSuccess)
        {
            set_error("cloud copy failed: %s", cudaGetErrorString(e));
            b200icp_cloud_free(c);
            return B200ICP_ERR_CUDA;
        }
    }
    if (int r = cloud_build_index(icp, L.ws, c))
    {
        b200icp_cloud_free(c);
        return r;
    }
    if (n)
    {
        cudaError_t e = cudaEventSynchronize(L.ws->ev[2]);
        if (e != cudaSuccess)
        {
            set_error("cloud copy failed: %s", cudaGetErrorString(e));
            b200icp_cloud_free(c);
            return B200ICP_ERR_CUDA;
        }
    }
    *out = c;
    return B200ICP_OK;
}

extern "C" int b200icp_cloud_upload(b200icp_t* icp, const float* x, const float* y, const float* z,
                                    size_t n, float search_radius, b200icp_cloud_t** out)
{
    return upload_common(icp, x, y, z, n, search_radius, cudaMemcpyHostToDevice, out);
}

extern "C" int b200icp_cloud_upload_raw(b200icp_t* icp, const float* x, const float* y, const float* z,
                                        size_t n, b200icp_cloud_t** out)
{
    return upload_common(icp, x, y, z, n, 0.f, cudaMemcpyHostToDevice, out, /*coords_only=*/true);
}

extern "C" int b200icp_cloud_from_device(b200icp_t* icp, const float* dx, const float* dy,
                                         const float* dz, size_t n, float search_radius,
                                         b200icp_cloud_t** out)
{
    return upload_common(icp, dx, dy, dz, n, search_radius, cudaMemcpyDeviceToDevice, out);
}

extern "C" void b200icp_cloud_free(b200icp_cloud_t* c)
{
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    // the index build is the last device work that touches the slab on the library's side; the registrations
    // that read it have returned (the caller frees a cloud after its calls)
    if (c->ready) cudaEventSynchronize(c->ready);
    if (c->slab)
    {
        Workspace* ws = c->ctx->acquire(/*upload=*/true);
        c->ctx->give_slab(c->slab, c->slab_bytes, ws ? ws->stream : nullptr);
        if (ws) c->ctx->release(ws);
    }
    if (c->ready) c->ctx->give_ready_event(c->ready);
    delete c;
}

extern "C" size_t b200icp_cloud_size(const b200icp_cloud_t* c) { return c ? c->n : 0; }
extern "C" size_t b200icp_cloud_device_bytes(const b200icp_cloud_t* c) { return c ? c->slab_bytes : 0; }

extern "C" int b200icp_cloud_download(const b200icp_cloud_t* c, float* x, float* y, float* z)
{
    if (!c || !x || !y || !z)
    {
        set_error("null argument");
        return B200ICP_ERR_BAD_ARG;
    }
    if (!c->n) return B200ICP_OK;
    Lease L(c->ctx);
    if (!L.ws) return B200ICP_ERR_CUDA;
    cudaStream_t s = L.ws->stream;
    B2_CUDA_TRY(cudaStreamWaitEvent(s, c->ready, 0));
    B2_CUDA_TRY(cudaMemcpyAsync(x, c->dx, c->n * sizeof(float), cudaMemcpyDeviceToHost, s));
    B2_CUDA_TRY(cudaMemcpyAsync(y, c->dy, c->n * sizeof(float), cudaMemcpyDeviceToHost, s));
    B2_CUDA_TRY(cudaMemcpyAsync(z, c->dz, c->n * sizeof(float), cudaMemcpyDeviceToHost, s));
    B2_CUDA_TRY(cudaStreamSynchronize(s));
    return B200ICP_OK;
}

extern "C" int b200icp_voxel_decimate(b200icp_t* icp, const b200icp_cloud_t* in, float resolution,
                                      int use_average, float search_radius, b200icp_cloud_t** out,
                                      uint32_t* keep_idx)
{
    if (!icp || !in || !out)
    {
        set_error("null argument");
        return B200ICP_ERR_BAD_ARG;
    }
    *out = nullptr;
    return run_voxel(icp, in, resolution, use_average, search_radius, out, keep_idx);
}

extern "C" void b200icp_edges_planes_defaults(b200icp_edges_planes_params_t* p)
{
    if (!p) return;
    p->voxel_filter_resolution = 1.0f;  // params/kitti-default.yaml:25-32
    p->full_pointcloud_decimation = 10;
    p->voxel_filter_decimation = 10;
    p->voxel_filter_max_e2_e0 = 30.f, p->voxel_filter_max_e1_e0 = 30.f;
    p->voxel_filter_min_e2_e0 = 80.f, p->voxel_filter_min_e1_e0 = 80.f;
    p->min_points_per_voxel = 5;
}

extern "C" int b200icp_filter_edges_planes(b200icp_t* icp, const b200icp_cloud_t* in,
                                           const b200icp_edges_planes_params_t* params, float search_radius,
                                           b200icp_cloud_t* layers_out[3], uint8_t* layer_flags_out,
                                           uint32_t* n_classified_voxels_out)
{
    if (!icp || !in || !params || !layers_out)
    {
        set_error("null argument");
        return B200ICP_ERR_BAD_ARG;
    }
    layers_out[0] = layers_out[1] = layers_out[2] = nullptr;
    return run_edges_planes(icp, in, params, search_radius, layers_out, layer_flags_out, n_classified_voxels_out);
}

extern "C" int b200icp_knn(b200icp_t* icp, const b200icp_cloud_t* ref, const b200icp_cloud_t* queries,
                           const double* pose6, uint32_t k, float max_dist, uint32_t* idx_out,
                           float* d2_out)
{
    if (!icp || !ref || !queries)
    {
        set_error("null argument");
        return B200ICP_ERR_BAD_ARG;
    }
    return run_knn(icp, ref, queries, pose6, k, max_dist, idx_out, d2_out);
}

extern "C" int b200icp_knn_keys_device(b200icp_t* icp, const b200icp_cloud_t* ref,
                                       const b200icp_cloud_t* queries, const double* pose6, uint32_t k,
                                       float max_dist, const uint32_t* d_index_map, uint64_t* d_keys_out)
{
    if (!icp || !ref || !queries || !d_keys_out)
    {
        set_error("null argument");
        return B200ICP_ERR_BAD_ARG;
    }
    return run_knn_keys(icp, ref, queries, pose6, k, max_dist, d_index_map, d_keys_out);
}

extern "C" int b200icp_knn_keys_scatter(b200icp_t* icp, const b200icp_cloud_t* ref,
                                        const b200icp_cloud_t* queries, const double* pose6, uint32_t k,
                                        float max_dist, const uint32_t* d_index_map,
                                        uint64_t* const* d_gather, uint32_t world, uint32_t rank, int atomic_min)
{
    if (!icp || !ref || !queries || !d_gather || world == 0 || world > 8 || rank >= world)
    {
        set_error("null argument or world outside [1,8]");
        return B200ICP_ERR_BAD_ARG;
    }
    for (uint32_t r = 0; r < world; r++)
        if (!d_gather[r])
        {
            set_error("gather buffer of rank %u is null", r);
            return B200ICP_ERR_BAD_ARG;
        }
    return run_knn_keys_scatter(icp, ref, queries, pose6, k, max_dist, d_index_map, d_gather, world, rank,
                                atomic_min);
}

extern "C" int b200icp_peer_alloc(b200icp_t* icp, size_t bytes, void** d_ptr, unsigned char handle_out[64])
{
    if (!icp || !d_ptr || !handle_out || bytes == 0)
    {
        set_error("null argument");
        return B200ICP_ERR_BAD_ARG;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    B2_CUDA_TRY(cudaSetDevice(icp->device));
    void* p = nullptr;
    B2_CUDA_TRY(cudaMalloc(&p, bytes));  // IPC needs a plain allocation, not the stream-ordered pool
    B2_CUDA_TRY(cudaMemset(p, 0, bytes));  // barrier flags start at zero
    B2_CUDA_TRY(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    cudaError_t        e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess)
    {
        set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        cudaFree(p);
        return B200ICP_ERR_CUDA;
    }
    memcpy(handle_out, &h, 64);
    *d_ptr = p;
    return B200ICP_OK;
}

extern "C" int b200icp_peer_free(b200icp_t* icp, void* d_ptr)
{
    if (!icp) return B200ICP_ERR_BAD_ARG;
    B2_CUDA_TRY(cudaSetDevice(icp->device));
    if (d_ptr) B2_CUDA_TRY(cudaFree(d_ptr));
    return B200ICP_OK;
}

extern "C" int b200icp_peer_open(b200icp_t* icp, const unsigned char handle[64], void** d_peer_ptr)
{
    if (!icp || !handle || !d_peer_ptr)
    {
        set_error("null argument");
        return B200ICP_ERR_BAD_ARG;
    }
    B2_CUDA_TRY(cudaSetDevice(icp->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* p = nullptr;
    B2_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *d_peer_ptr = p;
    return B200ICP_OK;
}

extern "C" int b200icp_peer_close(b200icp_t* icp, void* d_peer_ptr)
{
    if (!icp) return B200ICP_ERR_BAD_ARG;
    B2_CUDA_TRY(cudaSetDevice(icp->device));
    if (d_peer_ptr) B2_CUDA_TRY(cudaIpcCloseMemHandle(d_peer_ptr));
    return B200ICP_OK;
}

extern "C" int b200icp_peer_barrier(b200icp_t* icp, uint64_t* const* d_flags, uint32_t world, uint32_t rank,
                                    uint64_t epoch)
{
    if (!icp || !d_flags || world == 0 || world > 8 || rank >= world)
    {
        set_error("null argument or world outside [1,8]");
        return B200ICP_ERR_BAD_ARG;
    }
    return run_peer_barrier(icp, d_flags, world, rank, epoch);
}

extern "C" int b200icp_knn_keys_exchange(b200icp_t* icp, const b200icp_cloud_t* ref,
                                         const b200icp_cloud_t* queries, const double* pose6, uint32_t k,
                                         float max_dist, const uint32_t* d_index_map,
                                         uint64_t* const* d_bases, uint32_t world, uint32_t rank,
                                         uint64_t* epoch_io, uint64_t* d_out)
{
    if (!icp || !ref || !queries || !d_bases || !epoch_io || !d_out || world == 0 || world > 8 || rank >= world)
    {
        set_error("null argument or world outside [1,8]");
        return B200ICP_ERR_BAD_ARG;
    }
    for (uint32_t r = 0; r < world; r++)
        if (!d_bases[r])
        {
            set_error("exchange buffer of rank %u is null", r);
            return B200ICP_ERR_BAD_ARG;
        }
    return run_knn_exchange(icp, ref, queries, pose6, k, max_dist, d_index_map, d_bases, world, rank, epoch_io,
                            d_out);
}

extern "C" int b200icp_fill_no_key(b200icp_t* icp, uint64_t* d_keys, size_t n)
{
    if (!icp || (n && !d_keys))
    {
        set_error("null argument");
        return B200ICP_ERR_BAD_ARG;
    }
    return run_fill_no_key(icp, d_keys, n);
}

extern "C" int b200icp_merge_keys_device(b200icp_t* icp, const uint64_t* d_parts, uint32_t parts,
                                         size_t part_stride, size_t nq, uint32_t k, uint64_t* d_out)
{
    if (!icp || !d_parts || !d_out || parts == 0)
    {
        set_error("null argument");
        return B200ICP_ERR_BAD_ARG;
    }
    return run_merge_keys(icp, d_parts, parts, part_stride, nq, k, d_out);
}

extern "C" int b200icp_match(b200icp_t* icp, const b200icp_cloud_t* from_global,
                             const b200icp_cloud_t* to_local, const double* pose6, uint8_t* paired,
                             uint32_t* nn_idx, uint32_t* nn_cnt, double* centroid, double* normal,
                             uint32_t* n_pairings)
{
    if (!icp || !from_global || !to_local)
    {
        set_error("null argument");
        return B200ICP_ERR_BAD_ARG;
    }
    return run_match(icp, from_global, to_local, pose6, paired, nn_idx, nn_cnt, centroid, normal,
                     n_pairings);
}

extern "C" int b200icp_align(b200icp_t* icp, const b200icp_cloud_t* from_global,
                             const b200icp_cloud_t* to_local, const double guess6[6],
                             b200icp_result_t* out)
{
    if (!icp || !from_global || !to_local || !guess6 || !out)
    {
        set_error("null argument");  // ASSERT_(in.from_pc); ASSERT_(in.to_pc) cpp:860-861
        return B200ICP_ERR_BAD_ARG;
    }
    const b200icp_cloud_t* f[1] = {from_global};
    const b200icp_cloud_t* t[1] = {to_local};
    return run_align_batch(icp, 1, f, t, guess6, nullptr, out);
}

extern "C" void b200icp_call_params_of(const b200icp_params_t* p, b200icp_call_params_t* out)
{
    if (!p || !out) return;
    memset(out, 0, sizeof(*out));
    out->max_iterations = p->max_iterations;
    out->min_abs_step_trans = p->min_abs_step_trans;
    out->min_abs_step_rot = p->min_abs_step_rot;
    out->use_scale_outlier_detector = p->use_scale_outlier_detector;
    out->scale_outlier_threshold = p->scale_outlier_threshold;
    out->use_robust_kernel = p->use_robust_kernel;
    out->robust_kernel_param = p->robust_kernel_param;
    out->robust_kernel_scale = p->robust_kernel_scale;
}

extern "C" int b200icp_align_with(b200icp_t* icp, const b200icp_cloud_t* from_global,
                                  const b200icp_cloud_t* to_local, const double guess6[6],
                                  const b200icp_call_params_t* call, b200icp_result_t* out)
{
    if (!icp || !from_global || !to_local || !guess6 || !out)
    {
        set_error("null argument");
        return B200ICP_ERR_BAD_ARG;
    }
    const b200icp_cloud_t* f[1] = {from_global};
    const b200icp_cloud_t* t[1] = {to_local};
    return run_align_batch(icp, 1, f, t, guess6, call, out);
}

extern "C" int b200icp_align_batch(b200icp_t* icp, size_t n, const b200icp_cloud_t* const* from_global,
                                   const b200icp_cloud_t* const* to_local, const double* guesses6,
                                   b200icp_result_t* out)
{
    if (!icp || (n && (!from_global || !to_local || !guesses6 || !out)))
    {
        set_error("null argument");
        return B200ICP_ERR_BAD_ARG;
    }
    for (size_t i = 0; i < n; i++)
        if (!from_global[i] || !to_local[i])
        {
            set_error("null cloud in job %zu", i);
            return B200ICP_ERR_BAD_ARG;
        }
    return run_align_batch(icp, n, from_global, to_local, guesses6, nullptr, out);
}

// ---- sharded maps (sharded.inl) ---------------------------------------------------------------------------
extern "C" int b200icp_comm_unique_id(unsigned char id_out[B200ICP_COMM_ID_BYTES])
{
    if (!id_out)
    {
        set_error("null argument");
        return B200ICP_ERR_BAD_ARG;
    }
    return run_comm_unique_id(id_out);
}
extern "C" int b200icp_comm_create(b200icp_t* icp, const unsigned char id[B200ICP_COMM_ID_BYTES], int world, int rank,
                                   b200icp_comm_t** out)
{
    if (!icp || !id || !out || world < 1 || rank < 0 || rank >= world)
    {
        set_error("null argument or rank outside [0, world)");
        return B200ICP_ERR_BAD_ARG;
    }
    *out = nullptr;
    return run_comm_create(icp, id, world, rank, out);
}
extern "C" void b200icp_comm_destroy(b200icp_comm_t* comm) { run_comm_destroy(comm); }
extern "C" int b200icp_sharded_map_create(b200icp_comm_t* comm, const float* x, const float* y, const float* z, size_t n,
                                          float cell, int interleaved, float search_radius,
                                          b200icp_sharded_map_t** out)
{
    if (!comm || !out || (n && (!x || !y || !z)))
    {
        set_error("null argument");
        return B200ICP_ERR_BAD_ARG;
    }
    *out = nullptr;
    return run_sharded_map_create(comm, x, y, z, n, cell, interleaved, search_radius, out);
}
extern "C" void   b200icp_sharded_map_destroy(b200icp_sharded_map_t* map) { run_sharded_map_destroy(map); }
extern "C" size_t b200icp_sharded_map_local_size(const b200icp_sharded_map_t* map) { return run_sharded_map_local_size(map); }
extern "C" int b200icp_sharded_knn_keys(b200icp_sharded_map_t* map, const b200icp_cloud_t* queries, const double* pose6,
                                        uint32_t k, float max_dist, uint64_t* d_keys_out)
{
    if (!map || !queries || !d_keys_out)
    {
        set_error("null argument");
        return B200ICP_ERR_BAD_ARG;
    }
    return run_sharded_knn_keys(map, queries, pose6, k, max_dist, d_keys_out);
}
extern "C" int b200icp_sharded_align(b200icp_sharded_map_t* map, const b200icp_cloud_t* to_local, const double guess6[6],
                                     const b200icp_call_params_t* call, b200icp_result_t* out)
{
    if (!map || !to_local || !guess6 || !out)
    {
        set_error("null argument");
        return B200ICP_ERR_BAD_ARG;
    }
    return run_sharded_align(map, to_local, guess6, call, out);
}

extern "C" void b200icp_profile_enable(b200icp_t* icp, int enable)
{
    if (icp) icp->profile_on = enable != 0;
}
extern "C" void b200icp_profile_reset(b200icp_t* icp)
{
    if (!icp) return;
    std::lock_guard<std::mutex> lk(icp->mtx);
    icp->drain_pending();
    memset(&icp->prof, 0, sizeof(icp->prof));
}
extern "C" void b200icp_profile_get(b200icp_t* icp, b200icp_profile_t* out)
{
    if (!icp || !out) return;
    std::lock_guard<std::mutex> lk(icp->mtx);
    icp->drain_pending();
    *out = icp->prof;
}
extern "C" int b200icp_synchronize(b200icp_t* icp)
{
    if (!icp) return B200ICP_ERR_BAD_ARG;
    B2_CUDA_TRY(cudaSetDevice(icp->device));
    B2_CUDA_TRY(cudaDeviceSynchronize());
    return B200ICP_OK;
}
