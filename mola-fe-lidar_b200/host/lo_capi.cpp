// lo_capi.cpp -- extern "C" facade of include/b200_lidar_odometry.h.
#include <dlfcn.h>

#include <cstring>
#include <sstream>
#include <string>

#include "../../include/b200_lidar_odometry.h"
#include "LidarOdometry.h"

using namespace mola;

struct b200lo
{
    std::shared_ptr<WorldModel>    wm;
    std::shared_ptr<SimpleBackEnd> backend;
    std::unique_ptr<LidarOdometry> lo;
};

static thread_local std::string g_lo_err;
extern "C" const char* b200lo_last_error(void) { return g_lo_err.c_str(); }

// directory of the package = parent of the directory holding this library
static std::string package_dir()
{
    Dl_info info;
    if (dladdr((void*)&package_dir, &info) && info.dli_fname)
    {
        std::string  p = info.dli_fname;
        const size_t a = p.find_last_of('/');
        if (a != std::string::npos)
        {
            p = p.substr(0, a);  // .../lib
            const size_t b = p.find_last_of('/');
            if (b != std::string::npos) return p.substr(0, b);
        }
    }
    return ".";
}

extern "C" int b200lo_create(const char* yaml_path, const char* yaml_text, const char* mola_dir, b200lo_t** out)
{
    if (!out || (!yaml_path && !yaml_text))
    {
        g_lo_err = "null argument";
        return -1;
    }
    *out = nullptr;
    try
    {
        yaml_lite::Options opt;
        opt.module_dirs["mola-fe-lidar"] = mola_dir ? std::string(mola_dir) : package_dir();
        const Yaml root = yaml_path ? yaml_lite::parse_file(yaml_path, opt) : yaml_lite::parse(yaml_text, opt);
        auto       h = std::make_unique<b200lo>();
        h->wm = std::make_shared<WorldModel>();
        h->backend = std::make_shared<SimpleBackEnd>(h->wm);
        h->lo = std::make_unique<LidarOdometry>();
        h->lo->slam_backend_ = h->backend;
        h->lo->setWorldModel(h->wm);
        h->lo->initialize_common(root);
        h->lo->initialize(root);
        *out = h.release();
        return 0;
    }
    catch (const std::exception& e)
    {
        g_lo_err = e.what();
        return -3;
    }
}

extern "C" void b200lo_destroy(b200lo_t* lo)
{
    if (!lo) return;
    lo->lo.reset();  // frees clouds and ICP objects before the rest
    delete lo;
}

extern "C" void b200lo_reset(b200lo_t* lo)
{
    if (!lo) return;
    lo->lo->waitIdle();
    lo->lo->reset();
}

static CObservation::Ptr make_obs(const char* label, double t, const float* x, const float* y, const float* z,
                                  size_t n, bool copy)
{
    auto o = std::make_shared<CObservation>();
    o->sensorLabel = label ? label : "";
    o->timestamp = t;
    if (copy)
    {
        o->x.assign(x, x + n), o->y.assign(y, y + n), o->z.assign(z, z + n);
    }
    else
        o->px = x, o->py = y, o->pz = z, o->n = n;
    return o;
}

extern "C" int b200lo_on_new_observation(b200lo_t* lo, const char* label, double t, const float* x,
                                         const float* y, const float* z, size_t n)
{
    if (!lo || (n && (!x || !y || !z))) return -1;
    try
    {
        auto o = make_obs(label, t, x, y, z, n, true);
        lo->lo->onNewObservation(o);
        return 0;
    }
    catch (const std::exception& e)
    {
        g_lo_err = e.what();
        return -2;
    }
}

extern "C" int b200lo_enqueue_observation(b200lo_t* lo, const char* label, double t, const float* x,
                                          const float* y, const float* z, size_t n)
{
    if (!lo || (n && (!x || !y || !z))) return -1;
    try
    {
        auto o = make_obs(label, t, x, y, z, n, false);  // the caller keeps the buffers alive until processed
        lo->lo->onNewObservation(o);
        return 0;
    }
    catch (const std::exception& e)
    {
        g_lo_err = e.what();
        return -2;
    }
}

extern "C" size_t b200lo_queue_length(b200lo_t* lo) { return lo ? lo->lo->queueLength() : 0; }

extern "C" int b200lo_process_observation(b200lo_t* lo, const char* label, double t, const float* x,
                                          const float* y, const float* z, size_t n)
{
    if (!lo || (n && (!x || !y || !z))) return -1;
    try
    {
        auto o = make_obs(label, t, x, y, z, n, false);
        if (o->sensorLabel != lo->lo->raw_sensor_label_) return 0;
        lo->lo->doProcessNewObservation(o);
        return 0;
    }
    catch (const std::exception& e)
    {
        g_lo_err = e.what();
        return -2;
    }
}

extern "C" void b200lo_spin_once(b200lo_t* lo)
{
    if (lo) lo->lo->spinOnce();
}
extern "C" void b200lo_wait_idle(b200lo_t* lo)
{
    if (lo) lo->lo->waitIdle();
}

static void pose6(const CPose3D& p, double* o) { b2::pose_to_ypr(p, o); }

extern "C" int b200lo_get_state(b200lo_t* lo, b200lo_state_t* out)
{
    if (!lo || !out) return -1;
    const auto s = lo->lo->stateCopy();
    memset(out, 0, sizeof(*out));
    out->last_obs_tim = s.last_obs_tim;
    pose6(s.accum_since_last_kf, out->accum_since_last_kf);
    const double tw[6] = {s.last_iter_twist.vx, s.last_iter_twist.vy, s.last_iter_twist.vz,
                          s.last_iter_twist.wx, s.last_iter_twist.wy, s.last_iter_twist.wz};
    memcpy(out->last_twist, tw, sizeof(tw));
    out->last_iter_twist_is_good = s.last_iter_twist_is_good ? 1 : 0;
    out->last_kf = s.last_kf;
    {
        std::lock_guard<std::mutex> lk(lo->backend->mtx_);
        out->n_keyframes = lo->backend->kf_stamps.size();
        out->n_factors = lo->backend->factors.size();
        out->n_localizations = lo->backend->localizations.size();
    }
    out->n_processed = s.n_processed, out->n_dropped = s.n_dropped, out->n_icp = s.n_icp;
    out->last_icp_goodness = s.last_icp_out.goodness;
    pose6(s.last_icp_out.found_pose_to_wrt_from.mean, out->last_icp_pose);
    out->last_icp_iterations = s.last_icp_out.n_iterations;
    out->last_icp_termination = s.last_icp_out.termination_reason;
    out->last_points_size = s.last_points ? s.last_points->size() : 0;
    out->n_graph_edges = s.local_pose_graph.graph.edges.size();
    out->n_checked_pairs = s.local_pose_graph.checked_KF_pairs.size();
    out->n_kf_spills = DeviceCloud::spills().load();
    out->n_kf_reloads = DeviceCloud::reloads().load();
    return 0;
}

extern "C" size_t b200lo_last_montecarlo(b200lo_t* lo, uint64_t* from_kf, uint64_t* to_kf, double* guesses6,
                                         double* goodness, size_t cap, double* best_goodness, double* best_pose6)
{
    if (!lo) return 0;
    const auto   r = lo->lo->lastMonteCarlo();
    const size_t n = r.goodness.size();
    if (from_kf) *from_kf = r.from_id;
    if (to_kf) *to_kf = r.to_id;
    for (size_t i = 0; i < n && i < cap; i++)
    {
        if (guesses6) memcpy(guesses6 + 6 * i, r.guesses.data() + 6 * i, 6 * sizeof(double));
        if (goodness) goodness[i] = r.goodness[i];
    }
    if (best_goodness) *best_goodness = r.best_goodness;
    if (best_pose6) memcpy(best_pose6, r.best_pose, sizeof(r.best_pose));
    return n;
}

extern "C" size_t b200lo_get_factors(b200lo_t* lo, b200lo_factor_t* out, size_t cap)
{
    if (!lo) return 0;
    std::lock_guard<std::mutex> lk(lo->backend->mtx_);
    const auto&                 f = lo->backend->factors;
    for (size_t i = 0; i < f.size() && i < cap && out; i++)
    {
        out[i].from_kf = f[i].from_kf, out[i].to_kf = f[i].to_kf;
        const double p[6] = {f[i].rel_pose.x,   f[i].rel_pose.y,     f[i].rel_pose.z,
                             f[i].rel_pose.yaw, f[i].rel_pose.pitch, f[i].rel_pose.roll};
        memcpy(out[i].rel_pose, p, sizeof(p));
    }
    return f.size();
}

static size_t copy_out(const std::string& s, char* buf, size_t cap)
{
    if (buf && cap)
    {
        const size_t n = std::min(cap - 1, s.size());
        memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return s.size();
}

extern "C" size_t b200lo_dump_params(b200lo_t* lo, char* buf, size_t cap)
{
    if (!lo) return 0;
    const auto&        p = lo->lo->params_;
    std::ostringstream o;
    o.precision(17);
    o << "min_time_between_scans=" << p.min_time_between_scans << "\n"
      << "min_dist_xyz_between_keyframes=" << p.min_dist_xyz_between_keyframes << "\n"
      << "min_rotation_between_keyframes=" << p.min_rotation_between_keyframes << "\n"
      << "min_icp_goodness=" << p.min_icp_goodness << "\n"
      << "min_icp_goodness_lc=" << p.min_icp_goodness_lc << "\n"
      << "min_dist_to_matching=" << p.min_dist_to_matching << "\n"
      << "max_dist_to_matching=" << p.max_dist_to_matching << "\n"
      << "max_dist_to_loop_closure=" << p.max_dist_to_loop_closure << "\n"
      << "loop_closure_montecarlo_samples=" << p.loop_closure_montecarlo_samples << "\n"
      << "max_nearby_align_checks=" << p.max_nearby_align_checks << "\n"
      << "min_topo_dist_to_consider_loopclosure=" << p.min_topo_dist_to_consider_loopclosure << "\n"
      << "max_KFs_local_graph=" << p.max_KFs_local_graph << "\n"
      << "viz_decor_decimation=" << p.viz_decor_decimation << "\n"
      << "viz_decor_pointsize=" << p.viz_decor_pointsize << "\n"
      << "voxel_decimation_resolution=" << p.voxel_decimation_resolution << "\n"
      << "voxel_use_average=" << (p.voxel_use_average ? 1 : 0) << "\n"
      << "raw_sensor_label=" << lo->lo->raw_sensor_label_ << "\n";
    for (const auto& kv : p.icp)
    {
        const auto& q = kv.second.icpParameters;
        o << "icp[" << (int)kv.first << "].maxIterations=" << q.max_iterations << "\n"
          << "icp[" << (int)kv.first << "].minAbsStep_trans=" << q.min_abs_step_trans << "\n"
          << "icp[" << (int)kv.first << "].minAbsStep_rot=" << q.min_abs_step_rot << "\n"
          << "icp[" << (int)kv.first << "].solver_kind=" << q.solver_kind << "\n"
          << "icp[" << (int)kv.first << "].solver_maxIterations=" << q.solver_max_iterations << "\n"
          << "icp[" << (int)kv.first << "].matcher_kind=" << q.matcher_kind << "\n"
          << "icp[" << (int)kv.first << "].distanceThreshold=" << q.distance_threshold << "\n"
          << "icp[" << (int)kv.first << "].planeEigenThreshold=" << q.plane_eigen_threshold << "\n"
          << "icp[" << (int)kv.first << "].knn=" << q.knn << "\n"
          << "icp[" << (int)kv.first << "].quality_thresholdDistance=" << q.quality_threshold_distance << "\n";
    }
    return copy_out(o.str(), buf, cap);
}

extern "C" size_t b200lo_dump_profile(b200lo_t* lo, char* buf, size_t cap)
{
    if (!lo) return 0;
    std::ostringstream o;
    o.precision(9);
    for (const auto& kv : lo->lo->profiler_.stats())
        o << kv.first << "," << kv.second.n << "," << kv.second.total << "," << kv.second.max << "\n";
    return copy_out(o.str(), buf, cap);
}

extern "C" void* b200lo_icp_handle(b200lo_t* lo, int align_kind)
{
    if (!lo) return nullptr;
    auto it = lo->lo->params_.icp.find((LidarOdometry::AlignKind)align_kind);
    return it == lo->lo->params_.icp.end() ? nullptr : it->second.icp;
}
