#pragma once
#include <mp2p_icp/Pairings.h>
#include <mp2p_icp/metric_map.h>
#include <mrpt/containers/yaml.h>
#include <mrpt/poses/CPose3D.h>
#include <mrpt/rtti/CObject.h>
#include <vector>
namespace mp2p_icp
{
struct MatchContext { uint32_t icpIteration = 0; };
struct MatchState {};
class Matcher : public mrpt::rtti::CObject
{
   public:
    using Ptr = std::shared_ptr<Matcher>;
    virtual void initialize(const mrpt::containers::yaml& params)
    {
        runFromIteration = params.getOrDefault<uint32_t>("runFromIteration", 0);
        runUpToIteration = params.getOrDefault<uint32_t>("runUpToIteration", 0);
    }
    // public entry: honours runFrom/UpToIteration, then impl_match()
    virtual bool match(const metric_map_t& pcGlobal, const metric_map_t& pcLocal, const mrpt::poses::CPose3D& localPose,
                       const MatchContext& mc, MatchState& ms, Pairings& out) const
    {
        if (mc.icpIteration < runFromIteration) return false;
        if (runUpToIteration > 0 && mc.icpIteration > runUpToIteration) return false;
        return impl_match(pcGlobal, pcLocal, localPose, mc, ms, out);
    }
    uint32_t runFromIteration = 0, runUpToIteration = 0;
   protected:
    virtual bool impl_match(const metric_map_t& pcGlobal, const metric_map_t& pcLocal,
                            const mrpt::poses::CPose3D& localPose, const MatchContext& mc, MatchState& ms,
                            Pairings& out) const = 0;
};
using matcher_list_t = std::vector<Matcher::Ptr>;
}  // namespace mp2p_icp
