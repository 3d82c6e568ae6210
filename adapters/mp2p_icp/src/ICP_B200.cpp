#include <mola_b200/ICP_B200.h>
#include <mp2p_icp/Matcher_Point2Plane.h>
#include <mp2p_icp/QualityEvaluator.h>
#include <mp2p_icp/Solver.h>

#include <cstring>
#include <stdexcept>
#include <string>

IMPLEMENTS_MRPT_OBJECT(ICP_B200, mp2p_icp::ICP, mola)

namespace mola
{
namespace
{
[[noreturn]] void fail(const std::string& what) { throw std::runtime_error("mola::ICP_B200: " + what); }

// the part of the block that belongs to the OBJECT (solvers, matchers, quality evaluators); the rest is
// mp2p_icp::Parameters and travels with each call (b200icp_align_with)
bool same_object_params(const b200icp_params_t& a, const b200icp_params_t& b)
{
    return a.solver_kind == b.solver_kind && a.solver_max_iterations == b.solver_max_iterations &&
           a.gn_min_delta == b.gn_min_delta && a.matcher_kind == b.matcher_kind &&
           a.distance_threshold == b.distance_threshold && a.plane_eigen_threshold == b.plane_eigen_threshold &&
           a.knn == b.knn && a.min_plane_points == b.min_plane_points &&
           a.run_from_iteration == b.run_from_iteration && a.run_up_to_iteration == b.run_up_to_iteration &&
           a.quality_threshold_distance == b.quality_threshold_distance && a.cov_fd_step == b.cov_fd_step;
}
}  // namespace

ICP_B200::~ICP_B200()
{
    clouds_.clear();  // clouds before their contexts
    for (auto& c : ctxs_) b200icp_destroy(c.h);
}

b200icp_params_t ICP_B200::translate(const mp2p_icp::Parameters& p) const
{
    b200icp_params_t q;
    b200icp_default_params(&q);
    // mp2p_icp::Parameters (loaded at cpp:78 from `params:`)
    q.max_iterations             = p.maxIterations;
    q.min_abs_step_trans         = p.minAbsStep_trans;
    q.min_abs_step_rot           = p.minAbsStep_rot;
    const auto& w                = p.pairingsWeightParameters;
    q.use_scale_outlier_detector = w.use_scale_outlier_detector ? 1 : 0;
    q.scale_outlier_threshold    = w.scale_outlier_threshold;
    q.use_robust_kernel          = w.use_robust_kernel ? 1 : 0;
    q.robust_kernel_param        = w.robust_kernel_param;
    q.robust_kernel_scale        = w.robust_kernel_scale;

    // solvers (cpp:80-81)
    if (solvers().size() != 1) fail("exactly one solver is supported, got " + std::to_string(solvers().size()));
    if (auto gn = std::dynamic_pointer_cast<mp2p_icp::Solver_GaussNewton>(solvers()[0]))
    {
        q.solver_kind           = B200ICP_SOLVER_GAUSS_NEWTON;
        q.solver_max_iterations = gn->maxIterations;
    }
    else if (std::dynamic_pointer_cast<mp2p_icp::Solver_Horn>(solvers()[0]))
        q.solver_kind = B200ICP_SOLVER_HORN;
    else
        fail(std::string("solver class not available on the device: ") + solvers()[0]->className());

    // matchers (cpp:83-84)
    if (matchers().size() != 1) fail("exactly one matcher is supported, got " + std::to_string(matchers().size()));
    const auto& m0 = matchers()[0];
    if (auto m = std::dynamic_pointer_cast<mp2p_icp::Matcher_Point2Plane>(m0))
    {
        q.matcher_kind          = B200ICP_MATCHER_POINT2PLANE;
        q.distance_threshold    = m->distanceThreshold;
        q.plane_eigen_threshold = m->planeEigenThreshold;
        q.knn                   = m->knn;
        q.min_plane_points      = m->minimumPlanePoints;
    }
    else if (auto m = std::dynamic_pointer_cast<mp2p_icp::Matcher_Points_DistanceThreshold>(m0))
    {
        q.matcher_kind       = B200ICP_MATCHER_POINTS_DISTANCE;
        q.distance_threshold = m->threshold;
    }
    else
        fail(std::string("matcher class not available on the device: ") + m0->className());
    q.run_from_iteration  = m0->runFromIteration;
    q.run_up_to_iteration = m0->runUpToIteration;

    // quality evaluators (cpp:86-87)
    if (quality_evaluators().size() != 1)
        fail("exactly one quality evaluator is supported, got " + std::to_string(quality_evaluators().size()));
    if (auto e = std::dynamic_pointer_cast<mp2p_icp::QualityEvaluator_PairedRatio>(quality_evaluators()[0].obj))
        q.quality_threshold_distance = e->thresholdDistance;
    else
        fail(std::string("quality evaluator class not available on the device: ") +
             quality_evaluators()[0].obj->className());
    return q;
}

b200icp_t* ICP_B200::context_for(const b200icp_params_t& q)
{
    std::lock_guard<std::mutex> lk(mtx_);
    for (auto& c : ctxs_)
        if (same_object_params(c.q, q)) return c.h;
    b200icp_t* h = nullptr;
    if (b200icp_create(&q, device_, &h) != B200ICP_OK) fail(std::string("b200icp_create: ") + b200icp_last_error());
    ctxs_.push_back({q, h});
    return h;
}

void ICP_B200::align(const mp2p_icp::metric_map_t& pc1, const mp2p_icp::metric_map_t& pc2,
                     const mrpt::math::TPose3D& guess, const mp2p_icp::Parameters& p, mp2p_icp::Results& result)
{
    // the reference passes icp_params per call (cpp:871; chosen at cpp:287-290), the lists live in the object:
    // one device context per object configuration, the call's Parameters go with the call
    const b200icp_params_t q = translate(p);
    b200icp_t*             h = context_for(q);
    b200icp_call_params_t  call;
    b200icp_call_params_of(&q, &call);

    const auto global = pc1.point_layer(layer_);
    const auto local  = pc2.point_layer(layer_);
    // same search radius for both sides: either cloud may be the reference of a later call
    const auto g = clouds_.get(h, *global, 0.f);
    const auto l = clouds_.get(h, *local, 0.f);

    const double     g6[6] = {guess.x, guess.y, guess.z, guess.yaw, guess.pitch, guess.roll};
    b200icp_result_t r;
    if (b200icp_align_with(h, g->cloud, l->cloud, g6, &call, &r) != B200ICP_OK)
        fail(std::string("b200icp_align_with: ") + b200icp_last_error());

    result = mp2p_icp::Results();
    result.optimal_tf.mean = mrpt::poses::CPose3D(r.pose[0], r.pose[1], r.pose[2], r.pose[3], r.pose[4], r.pose[5]);
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) result.optimal_tf.cov(i, j) = r.cov[i * 6 + j];
    result.quality           = r.quality;
    result.nIterations       = r.n_iterations;
    result.terminationReason = static_cast<mp2p_icp::IterTermReason>(r.termination_reason);
}
}  // namespace mola
