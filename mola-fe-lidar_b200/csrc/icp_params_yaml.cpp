// icp_params_yaml.cpp -- one ICP settings block -> b200icp_params_t.
//
// Restates what load_icp_set_of_params() does with the block
// (LidarOdometry.cpp:57-88): `icp_class` must name a known mp2p_icp::ICP class
// (cpp:62-75), `params` feeds mp2p_icp::Parameters::load_from (cpp:77-78),
// `solvers` / `matchers` / `quality` are (class, params) lists resolved by
// class name (cpp:80-87).  Keys and defaults: params/icp-settings-regular.yaml.
#include <cmath>
#include <cstring>

#include "../../include/b200icp.h"
#include "../host/yaml_lite.h"
#include "runtime.cuh"

using yaml_lite::Node;

namespace
{
const double kDeg2Rad = 3.14159265358979323846 / 180.0;

const Node& single_entry(const Node& root, const char* key, std::string& cls)
{
    const Node& list = root.at(key);  // ENSURE_YAML_ENTRY_EXISTS
    if (!list.isSeq() || list.seq.empty())
        throw std::runtime_error(std::string("`") + key + "` must be a non-empty sequence of {class, params}");
    if (list.seq.size() != 1)
        throw std::runtime_error(std::string("`") + key + "`: this build runs exactly one entry, got " +
                                 std::to_string(list.seq.size()));
    const Node& e = list.seq[0];
    cls = e.at("class").as_string();
    return e["params"];
}
}  // namespace

extern "C" void b200icp_default_params(b200icp_params_t* p)
{
    memset(p, 0, sizeof(*p));
    p->max_iterations = 100;  // mp2p_icp::Parameters defaults overwritten by the YAML
    p->min_abs_step_trans = 5e-5;
    p->min_abs_step_rot = 1e-5;
    p->use_scale_outlier_detector = 1;
    p->scale_outlier_threshold = 1.1;
    p->use_robust_kernel = 0;
    p->robust_kernel_param = 0.1 * kDeg2Rad;
    p->robust_kernel_scale = 400.0;
    p->solver_kind = B200ICP_SOLVER_GAUSS_NEWTON;
    p->solver_max_iterations = 20;
    p->gn_min_delta = 1e-10;  // SURVEY Appendix A.6 (normative)
    p->matcher_kind = B200ICP_MATCHER_POINT2PLANE;
    p->distance_threshold = 0.70;
    p->plane_eigen_threshold = 0.07;
    p->knn = 6;
    p->min_plane_points = 3;  // A.5
    p->run_from_iteration = 0;
    p->run_up_to_iteration = 0;
    p->quality_threshold_distance = 0.10;
    p->cov_fd_step = 1e-7;  // A.9
}

extern "C" int b200icp_params_from_yaml(const char* yaml_text, b200icp_params_t* out)
{
    if (!yaml_text || !out)
    {
        b2::set_error("null argument");
        return B200ICP_ERR_BAD_ARG;
    }
    try
    {
        b200icp_default_params(out);
        const Node root = yaml_lite::parse(yaml_text);
        if (!root.isMap()) throw std::runtime_error("ICP settings block must be a map");

        std::string icp_class;
        root.load_req("icp_class", icp_class);  // YAML_LOAD_REQ, cpp:63
        if (icp_class != "mp2p_icp::ICP")
            throw std::runtime_error("icp_class=`" + icp_class +
                                     "` is a non-registered or incompatible class. Known classes: "
                                     "mp2p_icp::ICP");

        const Node& prm = root.at("params");
        unsigned int u;
        u = out->max_iterations, prm.load_opt("maxIterations", u), out->max_iterations = u;
        prm.load_opt("minAbsStep_trans", out->min_abs_step_trans);
        prm.load_opt("minAbsStep_rot", out->min_abs_step_rot);
        const Node& w = prm["pairingsWeightParameters"];
        if (w.isMap())
        {
            bool b;
            b = out->use_scale_outlier_detector != 0, w.load_opt("use_scale_outlier_detector", b),
            out->use_scale_outlier_detector = b;
            w.load_opt("scale_outlier_threshold", out->scale_outlier_threshold);
            b = out->use_robust_kernel != 0, w.load_opt("use_robust_kernel", b), out->use_robust_kernel = b;
            if (w.has("robust_kernel_param"))
                out->robust_kernel_param = w["robust_kernel_param"].as_double() * kDeg2Rad;  // [degrees]
            w.load_opt("robust_kernel_scale", out->robust_kernel_scale);
        }

        std::string cls;
        {
            const Node& sp = single_entry(root, "solvers", cls);
            if (cls == "mp2p_icp::Solver_GaussNewton")
                out->solver_kind = B200ICP_SOLVER_GAUSS_NEWTON;
            else if (cls == "mp2p_icp::Solver_Horn")
                out->solver_kind = B200ICP_SOLVER_HORN;
            else
                throw std::runtime_error("solver class=`" + cls +
                                         "` is a non-registered or incompatible class. Known classes: "
                                         "mp2p_icp::Solver_GaussNewton, mp2p_icp::Solver_Horn");
            if (sp.isMap())
            {
                u = out->solver_max_iterations, sp.load_opt("maxIterations", u), out->solver_max_iterations = u;
                sp.load_opt("minDelta", out->gn_min_delta);
            }
        }
        {
            const Node& mp = single_entry(root, "matchers", cls);
            if (cls == "mp2p_icp::Matcher_Point2Plane")
                out->matcher_kind = B200ICP_MATCHER_POINT2PLANE;
            else if (cls == "mp2p_icp::Matcher_Points_DistanceThreshold")
                out->matcher_kind = B200ICP_MATCHER_POINTS_DISTANCE;
            else
                throw std::runtime_error("matcher class=`" + cls +
                                         "` is a non-registered or incompatible class. Known classes: "
                                         "mp2p_icp::Matcher_Point2Plane, "
                                         "mp2p_icp::Matcher_Points_DistanceThreshold");
            if (mp.isMap())
            {
                if (mp.has("distanceThreshold"))
                    out->distance_threshold = mp["distanceThreshold"].as_double();
                else if (mp.has("threshold"))
                    out->distance_threshold = mp["threshold"].as_double();
                mp.load_opt("planeEigenThreshold", out->plane_eigen_threshold);
                u = out->knn, mp.load_opt("knn", u), out->knn = u;
                u = out->min_plane_points, mp.load_opt("minimumPlanePoints", u), out->min_plane_points = u;
                u = out->run_from_iteration, mp.load_opt("runFromIteration", u), out->run_from_iteration = u;
                u = out->run_up_to_iteration, mp.load_opt("runUpToIteration", u), out->run_up_to_iteration = u;
            }
        }
        {
            const Node& qp = single_entry(root, "quality", cls);
            if (cls != "mp2p_icp::QualityEvaluator_PairedRatio")
                throw std::runtime_error("quality class=`" + cls +
                                         "` is a non-registered or incompatible class. Known classes: "
                                         "mp2p_icp::QualityEvaluator_PairedRatio");
            if (qp.isMap()) qp.load_opt("thresholdDistance", out->quality_threshold_distance);
        }
        if (!(out->distance_threshold > 0) || !std::isfinite(out->distance_threshold))
            throw std::runtime_error("matcher distanceThreshold must be positive");
        if (out->knn < 1) throw std::runtime_error("matcher knn must be >= 1");
    }
    catch (const std::exception& e)
    {
        b2::set_error("%s", e.what());
        return B200ICP_ERR_YAML;
    }
    return B200ICP_OK;
}
