#pragma once
#include <cstddef>
namespace mrpt::math
{
template <typename T, std::size_t R, std::size_t C>
struct CMatrixFixed
{
    T        m[R * C] = {};
    T&       operator()(std::size_t r, std::size_t c) { return m[r * C + c]; }
    const T& operator()(std::size_t r, std::size_t c) const { return m[r * C + c]; }
    void     setZero() { for (auto& v : m) v = T(0); }
};
using CMatrixDouble66 = CMatrixFixed<double, 6, 6>;
}  // namespace mrpt::math
