#pragma once
namespace mrpt::math
{
struct TPoint3D  { double x = 0, y = 0, z = 0; TPoint3D() = default; TPoint3D(double a, double b, double c) : x(a), y(b), z(c) {} };
struct TPoint3Df { float  x = 0, y = 0, z = 0; TPoint3Df() = default; TPoint3Df(float a, float b, float c) : x(a), y(b), z(c) {} };
struct TVector3D : TPoint3D { using TPoint3D::TPoint3D; };
// a x + b y + c z + d = 0
struct TPlane
{
    double coefs[4] = {0, 0, 1, 0};
    TPlane() = default;
    TPlane(const TPoint3D& p, const TVector3D& n)
    {
        coefs[0] = n.x, coefs[1] = n.y, coefs[2] = n.z;
        coefs[3] = -(n.x * p.x + n.y * p.y + n.z * p.z);
    }
};
}  // namespace mrpt::math
