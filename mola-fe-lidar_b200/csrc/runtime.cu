// runtime.cu -- context / workspace pool / error string (host side).
#include "runtime.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace b2
{
static thread_local std::string g_err;

void set_error(const char* fmt, ...)
{
    char    buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}
const char* get_error() { return g_err.c_str(); }

int Workspace::init(int dev)
{
    device = dev;
    B2_CUDA_TRY(cudaSetDevice(dev));
    // a BLOCKING stream: it orders against the legacy default stream, so a
    // harness can bracket library work with events recorded on stream 0
    // (bench.py does); streams of different workspaces still run concurrently
    B2_CUDA_TRY(cudaStreamCreate(&stream));
    for (auto& e : ev) B2_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    B2_CUDA_TRY(cudaMalloc(&d_flag, 256));
    B2_CUDA_TRY(cudaMallocHost(&h_flag, 256));
    // pinned staging sized once for what a call moves (job records, results, counters): growing it later is a
    // cudaFreeHost + cudaMallocHost, both device-wide synchronisations
    if (int r = reserve_pinned((size_t)256 << 10)) return r;
    return B200ICP_OK;
}

void Workspace::destroy()
{
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    drop_align_graphs();
    if (d_scratch) cudaFree(d_scratch);
    if (h_pinned) cudaFreeHost(h_pinned);
    if (d_flag) cudaFree(d_flag);
    if (h_flag) cudaFreeHost(h_flag);
    d_flag = nullptr, h_flag = nullptr;
    for (auto& e : ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : prof_ev) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
    d_scratch = nullptr, h_pinned = nullptr, stream = nullptr;
}

int Workspace::reserve_device(size_t bytes)
{
    if (bytes <= d_bytes) return B200ICP_OK;
    // Growth is STREAM-ORDERED (the device's pool is kept warm, capi.cu): cudaFree / cudaMalloc would wait for every
    // stream of the device -- the registrations other threads are running -- each time a workspace meets its first
    // large cloud.  Work already enqueued on this stream keeps the old block until it has run.
    drop_align_graphs();  // they hold pointers into the old allocation
    if (d_scratch) B2_CUDA_TRY(cudaFreeAsync(d_scratch, stream));
    d_scratch = nullptr, d_bytes = 0;
    const size_t want = align_up(bytes + bytes / 4, 1 << 20);
    B2_CUDA_TRY(cudaMallocAsync(&d_scratch, want, stream));
    d_bytes = want;
    return B200ICP_OK;
}

int Workspace::reserve_pinned(size_t bytes)
{
    if (bytes <= h_bytes) return B200ICP_OK;
    B2_CUDA_TRY(cudaStreamSynchronize(stream));
    drop_align_graphs();
    if (h_pinned) B2_CUDA_TRY(cudaFreeHost(h_pinned));
    h_pinned = nullptr, h_bytes = 0;
    const size_t want = align_up(bytes + bytes / 4, 1 << 16);
    B2_CUDA_TRY(cudaMallocHost(&h_pinned, want));
    h_bytes = want;
    return B200ICP_OK;
}

int Workspace::reserve_prof_events(size_t pairs)
{
    while (prof_ev.size() < 2 * pairs)
    {
        cudaEvent_t e;
        B2_CUDA_TRY(cudaEventCreate(&e));
        prof_ev.push_back(e);
    }
    return B200ICP_OK;
}

void make_dev_params(const b200icp_params_t& P, IcpDevParams& D)
{
    memset(&D, 0, sizeof(D));
    D.max_iterations = P.max_iterations;
    D.min_abs_step_trans = P.min_abs_step_trans;
    D.min_abs_step_rot = P.min_abs_step_rot;
    D.solver_max_iterations = P.solver_max_iterations;
    D.gn_min_delta = P.gn_min_delta;
    D.matcher_kind = P.matcher_kind;
    D.thr = (float)P.distance_threshold;
    D.thr2 = D.thr * D.thr;  // float product (Appendix A.5)
    D.distance_threshold = P.distance_threshold;
    D.plane_eigen_threshold = P.plane_eigen_threshold;
    D.knn = P.knn;
    D.min_plane_points = P.min_plane_points;
    D.run_from_iteration = P.run_from_iteration;
    D.run_up_to_iteration = P.run_up_to_iteration;
    D.q_thr = (float)P.quality_threshold_distance;
    D.q_thr2 = D.q_thr * D.q_thr;
    D.cov_fd_step = P.cov_fd_step;
    D.solver_kind = P.solver_kind;
    D.use_scale_outlier_detector = P.use_scale_outlier_detector;
    D.scale_outlier_threshold = P.scale_outlier_threshold;
    D.use_robust_kernel = P.use_robust_kernel;
    D.robust_kernel_param = P.robust_kernel_param;
    D.robust_kernel_scale = P.robust_kernel_scale;
    const char* cyc = getenv("B200ICP_CYCLE");
    D.detect_cycles = cyc ? (uint32_t)std::max(0, std::min(2, atoi(cyc))) : 2u;
}
}  // namespace b2

b2::Workspace* b200icp::acquire(bool upload)
{
    {
        std::lock_guard<std::mutex> lk(mtx);
        auto&                       pool = upload ? free_upload_ws : free_ws;
        if (!pool.empty())
        {
            auto* w = pool.back();
            pool.pop_back();
            cudaSetDevice(device);
            return w;
        }
    }
    auto* w = new b2::Workspace();
    if (w->init(device) != B200ICP_OK)
    {
        delete w;
        return nullptr;
    }
    w->upload = upload;
    std::lock_guard<std::mutex> lk(mtx);
    all_ws.push_back(w);
    return w;
}

// smallest cached slab that fits and is not more than twice too large; else a fresh allocation
void* b200icp::take_slab(size_t need, size_t* got, cudaStream_t stream)
{
    {
        std::lock_guard<std::mutex> lk(mtx);
        int                         best = -1;
        for (int i = 0; i < (int)slab_cache.size(); i++)
            if (slab_cache[i].bytes >= need && slab_cache[i].bytes <= 2 * need &&
                (best < 0 || slab_cache[i].bytes < slab_cache[best].bytes))
                best = i;
        if (best >= 0)
        {
            const Slab sl = slab_cache[best];
            slab_cache.erase(slab_cache.begin() + best);
            slab_cache_bytes -= sl.bytes;
            *got = sl.bytes;
            return sl.p;
        }
    }
    // a miss goes to the device's stream-ordered pool (kept warm: release threshold = max, primed at create):
    // a few microseconds, where cudaMalloc costs hundreds and serialises with running work
    const size_t want = b2::align_up(need + need / 8, (size_t)1 << 20);  // room for the next, slightly larger scan
    void*        p = nullptr;
    if (cudaMallocAsync(&p, want, stream) != cudaSuccess)
    {
        cudaGetLastError();
        std::vector<Slab> drop;
        {
            std::lock_guard<std::mutex> lk(mtx);
            drop.swap(slab_cache);
            slab_cache_bytes = 0;
        }
        for (auto& d : drop) cudaFreeAsync(d.p, stream);
        if (cudaMallocAsync(&p, need, stream) != cudaSuccess)
        {
            cudaGetLastError();
            return nullptr;
        }
        *got = need;
        return p;
    }
    *got = want;
    return p;
}

void b200icp::give_slab(void* p, size_t bytes, cudaStream_t stream)
{
    constexpr size_t kMaxCached = 16;
    constexpr size_t kMaxBytes = (size_t)2 << 30;
    std::vector<Slab> drop;
    {
        std::lock_guard<std::mutex> lk(mtx);
        slab_cache.push_back({p, bytes});
        slab_cache_bytes += bytes;
        while (slab_cache.size() > kMaxCached || slab_cache_bytes > kMaxBytes)
        {   // oldest out
            drop.push_back(slab_cache.front());
            slab_cache_bytes -= slab_cache.front().bytes;
            slab_cache.erase(slab_cache.begin());
        }
    }
    for (auto& d : drop) cudaFreeAsync(d.p, stream);
}

cudaEvent_t b200icp::take_ready_event()
{
    {
        std::lock_guard<std::mutex> lk(mtx);
        if (!ready_events.empty())
        {
            cudaEvent_t e = ready_events.back();
            ready_events.pop_back();
            return e;
        }
    }
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    return e;
}

void b200icp::give_ready_event(cudaEvent_t e)
{
    std::lock_guard<std::mutex> lk(mtx);
    if (ready_events.size() < 64)
        ready_events.push_back(e);
    else
        cudaEventDestroy(e);
}

void b200icp::drain_pending()
{
    for (auto& p : pending_index)
    {
        float ms = 0;
        if (cudaEventSynchronize(p.e1) == cudaSuccess && cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess)
        {
            prof.index_builds++;
            prof.index_ms += ms;
            prof.index_points += p.points;
        }
        event_pool.push_back(p.e0);
        event_pool.push_back(p.e1);
    }
    pending_index.clear();
}

void b200icp::release(b2::Workspace* ws)
{
    std::lock_guard<std::mutex> lk(mtx);
    prof.total_kernel_launches += ws->launches;
    ws->launches = 0;
    (ws->upload ? free_upload_ws : free_ws).push_back(ws);
}
