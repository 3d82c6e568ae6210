// FilterEdgesPlanes_B200.h -- mp2p_icp_filters::FilterBase whose voxel classification runs on a B200.
//
// The filter stage of the reference (`apply_filter_pipeline(state_.pc_filter, *this_obs_points)`,
// /root/reference/src/LidarOdometry.cpp:223-224) with the class its parameter files name
// (params/kitti-default.yaml:21-32; defaults include/mola-fe-lidar/LidarOdometry.h:76-80).  Selected by class name in
// the `pointcloud_filter:` sequence:
//     pointcloud_filter:
//       - class_name: mola::FilterEdgesPlanes_B200
//         params: { voxel_filter_resolution: 1.0, voxel_filter_decimation: 10, full_pointcloud_decimation: 10,
//                   voxel_filter_max_e2_e0: 30, voxel_filter_max_e1_e0: 30, voxel_filter_min_e2_e0: 80,
//                   voxel_filter_min_e1_e0: 80 }
// Reads the layer `input_pointcloud_layer` (default "raw") and inserts the layers "edges", "planes" and
// "full_decim", each in ascending original index -- b200icp_filter_edges_planes (include/b200icp.h).
#pragma once
#include <b200icp.h>
#include <mp2p_icp_filters/FilterBase.h>

#include <mutex>
#include <string>

namespace mola
{
class FilterEdgesPlanes_B200 : public mp2p_icp_filters::FilterBase
{
    DEFINE_MRPT_OBJECT(FilterEdgesPlanes_B200, mola)

   public:
    FilterEdgesPlanes_B200();
    ~FilterEdgesPlanes_B200() override;

    void initialize(const mrpt::containers::yaml& cfg) override;
    void filter(mp2p_icp::metric_map_t& inOut) const override;

    b200icp_edges_planes_params_t params;
    std::string                   input_pointcloud_layer = mp2p_icp::metric_map_t::PT_LAYER_RAW;
    int                           device                 = 0;

   private:
    b200icp_t*         context() const;
    mutable std::mutex mtx_;
    mutable b200icp_t* h_ = nullptr;
};
}  // namespace mola
