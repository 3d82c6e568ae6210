#pragma once
// mp2p_icp::ICP as LidarOdometry.cpp:62-88, 869-871 uses it
#include <mp2p_icp/Matcher.h>
#include <mp2p_icp/Parameters.h>
#include <mp2p_icp/QualityEvaluator.h>
#include <mp2p_icp/Results.h>
#include <mp2p_icp/Solver.h>
#include <mp2p_icp/metric_map.h>
#include <mrpt/math/TPose3D.h>
namespace mp2p_icp
{
class ICP : public mrpt::rtti::CObject
{
    DEFINE_MRPT_OBJECT(ICP, mp2p_icp)
   public:
    // pc1 = reference ("from", global), pc2 = the cloud being moved ("to", local); the guess and
    // the result are the pose of pc2 with respect to pc1 (LidarOdometry.cpp:869-871, h:128-131)
    virtual void align(const metric_map_t& pc1, const metric_map_t& pc2,
                       const mrpt::math::TPose3D& init_guess_m2_wrt_m1, const Parameters& p, Results& result)
    {
        (void)pc1, (void)pc2, (void)init_guess_m2_wrt_m1, (void)p;
        result = Results();  // the stock CPU implementation is not part of the mock
    }
    // sequences of {class: <name>, params: {...}}; objects come from the class registry
    void initialize_solvers(const mrpt::containers::yaml& y)
    {
        solvers_.clear();
        for (const auto& e : y.asSequence())
        {
            auto o = mrpt::ptr_cast<Solver>::from(mrpt::rtti::classFactory(e["class"].as<std::string>()));
            if (!o) throw std::runtime_error("unknown solver class " + e["class"].as<std::string>());
            if (e.has("params")) o->initialize(e["params"]);
            solvers_.push_back(o);
        }
    }
    void initialize_matchers(const mrpt::containers::yaml& y)
    {
        matchers_.clear();
        for (const auto& e : y.asSequence())
        {
            auto o = mrpt::ptr_cast<Matcher>::from(mrpt::rtti::classFactory(e["class"].as<std::string>()));
            if (!o) throw std::runtime_error("unknown matcher class " + e["class"].as<std::string>());
            if (e.has("params")) o->initialize(e["params"]);
            matchers_.push_back(o);
        }
    }
    void initialize_quality_evaluators(const mrpt::containers::yaml& y)
    {
        quality_evaluators_.clear();
        for (const auto& e : y.asSequence())
        {
            auto o = mrpt::ptr_cast<QualityEvaluator>::from(mrpt::rtti::classFactory(e["class"].as<std::string>()));
            if (!o) throw std::runtime_error("unknown quality class " + e["class"].as<std::string>());
            if (e.has("params")) o->initialize(e["params"]);
            quality_evaluators_.push_back({o, e.getOrDefault<double>("weight", 1.0)});
        }
    }
    const solver_list_t&       solvers() const { return solvers_; }
    const matcher_list_t&      matchers() const { return matchers_; }
    const quality_eval_list_t& quality_evaluators() const { return quality_evaluators_; }
   protected:
    solver_list_t       solvers_;
    matcher_list_t      matchers_;
    quality_eval_list_t quality_evaluators_;
};
}  // namespace mp2p_icp
