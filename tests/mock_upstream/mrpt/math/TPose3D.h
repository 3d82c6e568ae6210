#pragma once
namespace mrpt::math
{
struct TPose3D
{
    double x = 0, y = 0, z = 0, yaw = 0, pitch = 0, roll = 0;
    TPose3D() = default;
    TPose3D(double x_, double y_, double z_, double yaw_, double pitch_, double roll_)
        : x(x_), y(y_), z(z_), yaw(yaw_), pitch(pitch_), roll(roll_) {}
};
}  // namespace mrpt::math
