// DeviceCloudCache.h -- device copies of mrpt point maps, kept by map identity.
//
// The reference hands the SAME shared_ptr<metric_map_t> to many align() calls:
// state_.last_points (LidarOdometry.cpp:231-234), the key-frame clouds kept in
// the world model (cpp:384-388) and re-fetched for every extra-edge / loop-
// closure job (cpp:658-666), ten Monte-Carlo runs per loop closure (cpp:775-787).
// MRPT builds each map's kd-tree once, lazily; the device counterpart is the
// indexed cloud of b200icp_cloud_upload, cached here instead of uploaded per call.
#pragma once
#include <b200icp.h>
#include <mrpt/maps/CPointsMap.h>

#include <cstdint>
#include <list>
#include <memory>
#include <mutex>
#include <unordered_map>

namespace mola_b200
{
/** One cached upload.  Shared: an entry evicted while an align() still uses it stays alive until released. */
struct DeviceCloud
{
    b200icp_t*       ctx   = nullptr;
    b200icp_cloud_t* cloud = nullptr;
    ~DeviceCloud()
    {
        if (cloud) b200icp_cloud_free(cloud);
    }
};

class DeviceCloudCache
{
   public:
    explicit DeviceCloudCache(std::size_t max_entries = 64) : max_entries_(max_entries) {}

    /** The device cloud of `m` for context `ctx`, uploading on a miss.  A map is recognised by its address, its
     *  size, the addresses of its three coordinate buffers and a hash of 64 sampled points, so that a map that was
     *  modified in place, or a new map allocated where a dead one used to be, is uploaded again.  Throws
     *  std::runtime_error with b200icp_last_error() on failure.  Thread-safe. */
    std::shared_ptr<DeviceCloud> get(b200icp_t* ctx, const mrpt::maps::CPointsMap& m, float search_radius);

    void        clear();
    std::size_t size() const;
    std::size_t uploads() const { return uploads_; }

   private:
    struct Key
    {
        const void* map;
        b200icp_t*  ctx;
        bool        operator==(const Key& o) const { return map == o.map && ctx == o.ctx; }
    };
    struct KeyHash
    {
        std::size_t operator()(const Key& k) const
        {
            return std::hash<const void*>()(k.map) ^ (std::hash<const void*>()(k.ctx) * 1000003u);
        }
    };
    struct Entry
    {
        std::shared_ptr<DeviceCloud> dev;
        std::size_t                  n = 0;
        const float *                px = nullptr, *py = nullptr, *pz = nullptr;
        std::uint64_t                fingerprint = 0;
        float                        radius = 0;
        std::list<Key>::iterator     lru;
    };
    static std::uint64_t fingerprint(const mrpt::maps::CPointsMap& m);

    mutable std::mutex                        mtx_;
    std::unordered_map<Key, Entry, KeyHash>   map_;
    std::list<Key>                            order_;  // most recently used first
    std::size_t                               max_entries_;
    std::size_t                               uploads_ = 0;
};
}  // namespace mola_b200
