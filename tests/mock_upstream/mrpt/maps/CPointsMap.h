#pragma once
#include <cstddef>
#include <memory>
#include <vector>
namespace mrpt::maps
{
class CMetricMap
{
   public:
    using Ptr = std::shared_ptr<CMetricMap>;
    virtual ~CMetricMap() = default;
};
// SoA float storage (the buffers the device upload reads directly)
class CPointsMap : public CMetricMap
{
   public:
    using Ptr = std::shared_ptr<CPointsMap>;
    std::size_t size() const { return x_.size(); }
    bool        empty() const { return x_.empty(); }
    const std::vector<float>& getPointsBufferRef_x() const { return x_; }
    const std::vector<float>& getPointsBufferRef_y() const { return y_; }
    const std::vector<float>& getPointsBufferRef_z() const { return z_; }
    void insertPointFast(float x, float y, float z) { x_.push_back(x), y_.push_back(y), z_.push_back(z); }
   private:
    std::vector<float> x_, y_, z_;
};
class CSimplePointsMap : public CPointsMap
{
   public:
    static std::shared_ptr<CSimplePointsMap> Create() { return std::make_shared<CSimplePointsMap>(); }
};
}  // namespace mrpt::maps
