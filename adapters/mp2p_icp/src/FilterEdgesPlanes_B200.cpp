#include <mola_b200/FilterEdgesPlanes_B200.h>
#include <mrpt/maps/CPointsMap.h>

#include <stdexcept>
#include <vector>

IMPLEMENTS_MRPT_OBJECT(FilterEdgesPlanes_B200, mp2p_icp_filters::FilterBase, mola)

namespace mola
{
FilterEdgesPlanes_B200::FilterEdgesPlanes_B200() { b200icp_edges_planes_defaults(&params); }

FilterEdgesPlanes_B200::~FilterEdgesPlanes_B200()
{
    if (h_) b200icp_destroy(h_);
}

void FilterEdgesPlanes_B200::initialize(const mrpt::containers::yaml& c)
{
    params.voxel_filter_resolution    = c.getOrDefault<float>("voxel_filter_resolution", params.voxel_filter_resolution);
    params.full_pointcloud_decimation = c.getOrDefault<uint32_t>("full_pointcloud_decimation", params.full_pointcloud_decimation);
    params.voxel_filter_decimation    = c.getOrDefault<uint32_t>("voxel_filter_decimation", params.voxel_filter_decimation);
    params.voxel_filter_max_e2_e0     = c.getOrDefault<float>("voxel_filter_max_e2_e0", params.voxel_filter_max_e2_e0);
    params.voxel_filter_max_e1_e0     = c.getOrDefault<float>("voxel_filter_max_e1_e0", params.voxel_filter_max_e1_e0);
    params.voxel_filter_min_e2_e0     = c.getOrDefault<float>("voxel_filter_min_e2_e0", params.voxel_filter_min_e2_e0);
    params.voxel_filter_min_e1_e0     = c.getOrDefault<float>("voxel_filter_min_e1_e0", params.voxel_filter_min_e1_e0);
    params.min_points_per_voxel       = c.getOrDefault<uint32_t>("b200_min_points_per_voxel", params.min_points_per_voxel);
    input_pointcloud_layer            = c.getOrDefault<std::string>("input_pointcloud_layer", input_pointcloud_layer);
    device                            = c.getOrDefault<int>("b200_device", device);
}

b200icp_t* FilterEdgesPlanes_B200::context() const
{
    std::lock_guard<std::mutex> lk(mtx_);
    if (h_) return h_;
    b200icp_params_t q;
    b200icp_default_params(&q);
    if (b200icp_create(&q, device, &h_) != B200ICP_OK)
        throw std::runtime_error(std::string("mola::FilterEdgesPlanes_B200: b200icp_create: ") + b200icp_last_error());
    return h_;
}

void FilterEdgesPlanes_B200::filter(mp2p_icp::metric_map_t& inOut) const
{
    b200icp_t* h  = context();
    const auto in = inOut.point_layer(input_pointcloud_layer);
    static const char* names[3] = {"edges", "planes", "full_decim"};

    b200icp_cloud_t* raw = nullptr;
    if (b200icp_cloud_upload_raw(h, in->getPointsBufferRef_x().data(), in->getPointsBufferRef_y().data(),
                                 in->getPointsBufferRef_z().data(), in->size(), &raw) != B200ICP_OK)
        throw std::runtime_error(std::string("mola::FilterEdgesPlanes_B200: upload: ") + b200icp_last_error());
    b200icp_cloud_t* layers[3] = {nullptr, nullptr, nullptr};
    const int        rc        = b200icp_filter_edges_planes(h, raw, &params, 0.f, layers, nullptr, nullptr);
    b200icp_cloud_free(raw);
    if (rc != B200ICP_OK)
        throw std::runtime_error(std::string("mola::FilterEdgesPlanes_B200: filter: ") + b200icp_last_error());
    for (int l = 0; l < 3; l++)
    {
        const std::size_t  m = b200icp_cloud_size(layers[l]);
        std::vector<float> x(m), y(m), z(m);
        if (m && b200icp_cloud_download(layers[l], x.data(), y.data(), z.data()) != B200ICP_OK)
        {
            for (auto* c : layers) b200icp_cloud_free(c);
            throw std::runtime_error(std::string("mola::FilterEdgesPlanes_B200: download: ") + b200icp_last_error());
        }
        auto pm = mrpt::maps::CSimplePointsMap::Create();
        for (std::size_t i = 0; i < m; i++) pm->insertPointFast(x[i], y[i], z[i]);
        inOut.layers[names[l]] = pm;
    }
    for (auto* c : layers) b200icp_cloud_free(c);
}
}  // namespace mola
