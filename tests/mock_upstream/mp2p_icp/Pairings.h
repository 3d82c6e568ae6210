#pragma once
#include <mrpt/math/TPoint3D.h>
#include <vector>
namespace mp2p_icp
{
struct plane_patch_t
{
    mrpt::math::TPlane   plane;
    mrpt::math::TPoint3D centroid;
};
struct point_plane_pair_t
{
    plane_patch_t          pl_global;
    mrpt::math::TPoint3Df  pt_local;
};
struct Pairings
{
    std::vector<point_plane_pair_t> paired_pt2pl;
    bool   empty() const { return paired_pt2pl.empty(); }
    size_t size() const { return paired_pt2pl.size(); }
};
}  // namespace mp2p_icp
