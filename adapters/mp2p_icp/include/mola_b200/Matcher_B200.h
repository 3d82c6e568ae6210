// Matcher_B200.h -- mp2p_icp::Matcher whose pairing search runs on a B200.
//
// For installations that keep the stock mp2p_icp::ICP loop and solvers on the CPU and move only the dominant
// stage to the device: select it by class name in the `matchers:` sequence the reference hands to
// initialize_matchers (/root/reference/src/LidarOdometry.cpp:83-84; params/icp-settings-regular.yaml:32-39):
//     matchers:
//       - class: mola::Matcher_B200
//         params: { distanceThreshold: 0.70, planeEigenThreshold: 0.07, knn: 6 }
// Same parameters and the same pairings as mp2p_icp::Matcher_Point2Plane: per local point the k nearest global
// points within distanceThreshold, plane fit, eigenvalue-ratio and point-to-plane distance gates.
#pragma once
#include <b200icp.h>
#include <mola_b200/DeviceCloudCache.h>
#include <mp2p_icp/Matcher.h>

#include <mutex>

namespace mola
{
class Matcher_B200 : public mp2p_icp::Matcher
{
    DEFINE_MRPT_OBJECT(Matcher_B200, mola)

   public:
    Matcher_B200() = default;
    ~Matcher_B200() override;

    void initialize(const mrpt::containers::yaml& params) override;

    double      distanceThreshold   = 0.50;
    double      planeEigenThreshold = 0.01;
    uint32_t    knn                 = 5;
    uint32_t    minimumPlanePoints  = 3;
    int         device              = 0;
    std::string layer               = mp2p_icp::metric_map_t::PT_LAYER_RAW;

   protected:
    bool impl_match(const mp2p_icp::metric_map_t& pcGlobal, const mp2p_icp::metric_map_t& pcLocal,
                    const mrpt::poses::CPose3D& localPose, const mp2p_icp::MatchContext& mc, mp2p_icp::MatchState& ms,
                    mp2p_icp::Pairings& out) const override;

   private:
    b200icp_t* context() const;
    mutable std::mutex                  mtx_;
    mutable b200icp_t*                  h_ = nullptr;
    mutable mola_b200::DeviceCloudCache clouds_;
};
}  // namespace mola
