#pragma once
#include <mrpt/math/TPose3D.h>
namespace mrpt::poses
{
class CPose3D
{
   public:
    CPose3D() = default;
    CPose3D(double x, double y, double z, double yaw = 0, double pitch = 0, double roll = 0) : p_(x, y, z, yaw, pitch, roll) {}
    explicit CPose3D(const mrpt::math::TPose3D& p) : p_(p) {}
    mrpt::math::TPose3D asTPose() const { return p_; }
    double x() const { return p_.x; }
    double y() const { return p_.y; }
    double z() const { return p_.z; }
    double yaw() const { return p_.yaw; }
    double pitch() const { return p_.pitch; }
    double roll() const { return p_.roll; }
   private:
    mrpt::math::TPose3D p_;
};
}  // namespace mrpt::poses
