#pragma once
// the shape of mp2p_icp_filters::FilterBase (what apply_filter_pipeline calls, /root/reference/src/LidarOdometry.cpp:223-224)
#include <mp2p_icp/metric_map.h>
#include <mrpt/containers/yaml.h>
#include <mrpt/rtti/CObject.h>
#include <vector>
namespace mp2p_icp_filters
{
class FilterBase : public mrpt::rtti::CObject
{
   public:
    using Ptr = std::shared_ptr<FilterBase>;
    virtual void initialize(const mrpt::containers::yaml& cfg) = 0;
    // reads the input layer(s) of `inOut` and adds / replaces its output layers
    virtual void filter(mp2p_icp::metric_map_t& inOut) const = 0;
};
using FilterPipeline = std::vector<FilterBase::Ptr>;
inline void apply_filter_pipeline(const FilterPipeline& filters, mp2p_icp::metric_map_t& inOut)
{
    for (const auto& f : filters) f->filter(inOut);
}
}  // namespace mp2p_icp_filters
