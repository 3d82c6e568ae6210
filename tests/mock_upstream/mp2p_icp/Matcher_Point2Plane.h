#pragma once
#include <mp2p_icp/Matcher.h>
namespace mp2p_icp
{
class Matcher_Point2Plane : public Matcher
{
    DEFINE_MRPT_OBJECT(Matcher_Point2Plane, mp2p_icp)
   public:
    void initialize(const mrpt::containers::yaml& params) override
    {
        Matcher::initialize(params);
        distanceThreshold   = params.getOrDefault<double>("distanceThreshold", distanceThreshold);
        planeEigenThreshold = params.getOrDefault<double>("planeEigenThreshold", planeEigenThreshold);
        knn                 = params.getOrDefault<uint32_t>("knn", knn);
        minimumPlanePoints  = params.getOrDefault<uint32_t>("minimumPlanePoints", minimumPlanePoints);
    }
    double   distanceThreshold   = 0.50;
    double   planeEigenThreshold = 0.01;
    uint32_t knn                 = 5;
    uint32_t minimumPlanePoints  = 3;
   protected:
    bool impl_match(const metric_map_t&, const metric_map_t&, const mrpt::poses::CPose3D&, const MatchContext&,
                    MatchState&, Pairings&) const override { return false; }
};
class Matcher_Points_DistanceThreshold : public Matcher
{
    DEFINE_MRPT_OBJECT(Matcher_Points_DistanceThreshold, mp2p_icp)
   public:
    void initialize(const mrpt::containers::yaml& params) override
    {
        Matcher::initialize(params);
        threshold = params.getOrDefault<double>("threshold", threshold);
    }
    double threshold = 0.50;
   protected:
    bool impl_match(const metric_map_t&, const metric_map_t&, const mrpt::poses::CPose3D&, const MatchContext&,
                    MatchState&, Pairings&) const override { return false; }
};
}  // namespace mp2p_icp
