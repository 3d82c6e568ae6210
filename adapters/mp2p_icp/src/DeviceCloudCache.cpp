#include <mola_b200/DeviceCloudCache.h>

#include <cstring>
#include <stdexcept>
#include <string>

namespace mola_b200
{
std::uint64_t DeviceCloudCache::fingerprint(const mrpt::maps::CPointsMap& m)
{
    const auto&       x = m.getPointsBufferRef_x();
    const auto&       y = m.getPointsBufferRef_y();
    const auto&       z = m.getPointsBufferRef_z();
    const std::size_t n = x.size();
    std::uint64_t     h = 1469598103934665603ull ^ n;  // FNV-1a over 64 strided samples
    const std::size_t step = n > 64 ? n / 64 : 1;
    for (std::size_t i = 0; i < n; i += step)
    {
        std::uint32_t w[3];
        std::memcpy(&w[0], &x[i], 4), std::memcpy(&w[1], &y[i], 4), std::memcpy(&w[2], &z[i], 4);
        for (std::uint32_t v : w) h = (h ^ v) * 1099511628211ull;
    }
    if (n)
    {  // and the last point: appends change it
        std::uint32_t w[3];
        std::memcpy(&w[0], &x[n - 1], 4), std::memcpy(&w[1], &y[n - 1], 4), std::memcpy(&w[2], &z[n - 1], 4);
        for (std::uint32_t v : w) h = (h ^ v) * 1099511628211ull;
    }
    return h;
}

std::shared_ptr<DeviceCloud> DeviceCloudCache::get(b200icp_t* ctx, const mrpt::maps::CPointsMap& m, float search_radius)
{
    const Key         key{&m, ctx};
    const auto&       x = m.getPointsBufferRef_x();
    const auto&       y = m.getPointsBufferRef_y();
    const auto&       z = m.getPointsBufferRef_z();
    const std::size_t n = m.size();
    const auto        fp = fingerprint(m);
    {
        std::lock_guard<std::mutex> lk(mtx_);
        auto                        it = map_.find(key);
        if (it != map_.end())
        {
            Entry& e = it->second;
            if (e.n == n && e.px == x.data() && e.py == y.data() && e.pz == z.data() && e.fingerprint == fp &&
                e.radius == search_radius)
            {
                order_.splice(order_.begin(), order_, e.lru);
                return e.dev;
            }
            order_.erase(e.lru);  // same address, other contents
            map_.erase(it);
        }
    }
    // upload outside the lock: other threads keep hitting the cache meanwhile
    auto dev = std::make_shared<DeviceCloud>();
    dev->ctx = ctx;
    if (b200icp_cloud_upload(ctx, x.data(), y.data(), z.data(), n, search_radius, &dev->cloud) != B200ICP_OK)
        throw std::runtime_error(std::string("b200icp_cloud_upload: ") + b200icp_last_error());
    std::lock_guard<std::mutex> lk(mtx_);
    uploads_++;
    auto it = map_.find(key);
    if (it != map_.end())
    {  // another thread uploaded the same map meanwhile: keep the first
        order_.splice(order_.begin(), order_, it->second.lru);
        return it->second.dev;
    }
    order_.push_front(key);
    Entry e;
    e.dev = dev, e.n = n, e.px = x.data(), e.py = y.data(), e.pz = z.data(), e.fingerprint = fp;
    e.radius = search_radius, e.lru = order_.begin();
    map_.emplace(key, std::move(e));
    while (map_.size() > max_entries_)
    {
        map_.erase(order_.back());
        order_.pop_back();
    }
    return dev;
}

void DeviceCloudCache::clear()
{
    std::lock_guard<std::mutex> lk(mtx_);
    map_.clear();
    order_.clear();
}

std::size_t DeviceCloudCache::size() const
{
    std::lock_guard<std::mutex> lk(mtx_);
    return map_.size();
}
}  // namespace mola_b200
