// runtime.cu -- context / workspace pool / error string (host side).
#include "runtime.cuh"

#include <cmath>
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
    B2_CUDA_TRY(cudaStreamSynchronize(stream));
    drop_align_graphs();  // they hold pointers into the old allocation
    if (d_scratch) B2_CUDA_TRY(cudaFree(d_scratch));
    d_scratch = nullptr, d_bytes = 0;
    const size_t want = align_up(bytes + bytes / 4, 1 << 20);
    B2_CUDA_TRY(cudaMalloc(&d_scratch, want));
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
}
}  // namespace b2

b2::Workspace* b200icp::acquire()
{
    {
        std::lock_guard<std::mutex> lk(mtx);
        if (!free_ws.empty())
        {
            auto* w = free_ws.back();
            free_ws.pop_back();
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
    std::lock_guard<std::mutex> lk(mtx);
    all_ws.push_back(w);
    return w;
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
    free_ws.push_back(ws);
}
