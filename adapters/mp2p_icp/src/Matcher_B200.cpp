#include <mola_b200/Matcher_B200.h>

#include <stdexcept>
#include <string>
#include <vector>

IMPLEMENTS_MRPT_OBJECT(Matcher_B200, mp2p_icp::Matcher, mola)

namespace mola
{
Matcher_B200::~Matcher_B200()
{
    clouds_.clear();
    if (h_) b200icp_destroy(h_);
}

void Matcher_B200::initialize(const mrpt::containers::yaml& params)
{
    mp2p_icp::Matcher::initialize(params);  // runFromIteration / runUpToIteration
    distanceThreshold   = params.getOrDefault<double>("distanceThreshold", distanceThreshold);
    planeEigenThreshold = params.getOrDefault<double>("planeEigenThreshold", planeEigenThreshold);
    knn                 = params.getOrDefault<uint32_t>("knn", knn);
    minimumPlanePoints  = params.getOrDefault<uint32_t>("minimumPlanePoints", minimumPlanePoints);
    device              = params.getOrDefault<int>("b200_device", device);
}

b200icp_t* Matcher_B200::context() const
{
    std::lock_guard<std::mutex> lk(mtx_);
    if (h_) return h_;
    b200icp_params_t q;
    b200icp_default_params(&q);
    q.matcher_kind          = B200ICP_MATCHER_POINT2PLANE;
    q.distance_threshold    = distanceThreshold;
    q.plane_eigen_threshold = planeEigenThreshold;
    q.knn                   = knn;
    q.min_plane_points      = minimumPlanePoints;
    q.run_from_iteration = q.run_up_to_iteration = 0;  // gating is done by Matcher::match() before impl_match
    if (b200icp_create(&q, device, &h_) != B200ICP_OK)
        throw std::runtime_error(std::string("mola::Matcher_B200: b200icp_create: ") + b200icp_last_error());
    return h_;
}

bool Matcher_B200::impl_match(const mp2p_icp::metric_map_t& pcGlobal, const mp2p_icp::metric_map_t& pcLocal,
                              const mrpt::poses::CPose3D& localPose, const mp2p_icp::MatchContext&,
                              mp2p_icp::MatchState&, mp2p_icp::Pairings& out) const
{
    b200icp_t* h      = context();
    const auto global = pcGlobal.point_layer(layer);
    const auto local  = pcLocal.point_layer(layer);
    if (global->empty() || local->empty()) return false;
    const auto g = clouds_.get(h, *global, 0.f);
    const auto l = clouds_.get(h, *local, 0.f);

    const std::size_t    n = local->size();
    std::vector<uint8_t> paired(n);
    std::vector<double>  centroid(3 * n), normal(3 * n);
    uint32_t             n_pairings = 0;
    const double         pose6[6]   = {localPose.x(),   localPose.y(),     localPose.z(),
                                       localPose.yaw(), localPose.pitch(), localPose.roll()};
    if (b200icp_match(h, g->cloud, l->cloud, pose6, paired.data(), nullptr, nullptr, centroid.data(), normal.data(),
                      &n_pairings) != B200ICP_OK)
        throw std::runtime_error(std::string("mola::Matcher_B200: b200icp_match: ") + b200icp_last_error());

    const auto& lx = local->getPointsBufferRef_x();
    const auto& ly = local->getPointsBufferRef_y();
    const auto& lz = local->getPointsBufferRef_z();
    out.paired_pt2pl.reserve(out.paired_pt2pl.size() + n_pairings);
    for (std::size_t i = 0; i < n; i++)
    {
        if (!paired[i]) continue;
        mp2p_icp::point_plane_pair_t p;
        p.pl_global.centroid = mrpt::math::TPoint3D(centroid[3 * i], centroid[3 * i + 1], centroid[3 * i + 2]);
        p.pl_global.plane    = mrpt::math::TPlane(
            p.pl_global.centroid, mrpt::math::TVector3D(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]));
        p.pt_local = mrpt::math::TPoint3Df(lx[i], ly[i], lz[i]);  // the ORIGINAL local point (A.5)
        out.paired_pt2pl.push_back(p);
    }
    return n_pairings != 0;
}
}  // namespace mola
