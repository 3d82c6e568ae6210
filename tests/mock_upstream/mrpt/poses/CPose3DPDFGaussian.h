#pragma once
#include <mrpt/math/CMatrixFixed.h>
#include <mrpt/poses/CPose3D.h>
namespace mrpt::poses
{
class CPose3DPDFGaussian
{
   public:
    CPose3D                      mean;
    mrpt::math::CMatrixDouble66  cov;
    const CPose3D& getMeanVal() const { return mean; }
};
}  // namespace mrpt::poses
