#pragma once
#include <mrpt/maps/CPointsMap.h>
#include <map>
#include <stdexcept>
#include <string>
namespace mp2p_icp
{
using layer_name_t = std::string;
struct metric_map_t
{
    constexpr static const char* PT_LAYER_RAW = "raw";
    std::map<layer_name_t, mrpt::maps::CMetricMap::Ptr> layers;
    // throws when the layer is missing or is not a point map
    mrpt::maps::CPointsMap::Ptr point_layer(const layer_name_t& name) const
    {
        auto it = layers.find(name);
        if (it == layers.end()) throw std::runtime_error("metric_map_t: no layer " + name);
        auto p = std::dynamic_pointer_cast<mrpt::maps::CPointsMap>(it->second);
        if (!p) throw std::runtime_error("metric_map_t: layer is not a point map: " + name);
        return p;
    }
    bool empty() const { return layers.empty(); }
};
}  // namespace mp2p_icp
