// mola_stubs.h -- the slice of mola-kernel / MRPT that mola::LidarOdometry
// touches, as small standalone types, so that the front-end restatement builds
// and runs without the (absent) MOLA stack.  Shapes follow the reference's
// call sites: FrontEndBase (LidarOdometry.h:29-43; cpp:169 raw_sensor_label_,
// cpp:359 slam_backend_), BackEndBase::addKeyFrame / addFactor /
// advertiseUpdatedLocalization with std::future results (cpp:346-364, 440-455,
// 484-490), WorldModel annotations + neighbours + RW locks (cpp:377-428,
// 614-669).  A real MOLA build replaces this header with <mola-kernel/...>.
#pragma once
#include <atomic>
#include <cstdint>
#include <functional>
#include <future>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <set>
#include <shared_mutex>
#include <string>
#include <vector>

#include "../../include/b200icp.h"
#include "yaml_lite.h"

namespace mola
{
using id_t = std::uint64_t;
using fid_t = std::uint64_t;
constexpr id_t  INVALID_ID = static_cast<id_t>(-1);
constexpr fid_t INVALID_FID = static_cast<fid_t>(-1);
using Yaml = yaml_lite::Node;

/** mrpt::math::TPose3D: x y z yaw pitch roll */
struct TPose3D
{
    double x = 0, y = 0, z = 0, yaw = 0, pitch = 0, roll = 0;
};

/** mrpt::math::TTwist3D */
struct TTwist3D
{
    double vx = 0, vy = 0, vz = 0, wx = 0, wy = 0, wz = 0;
};

/** mrpt::obs::CObservationPointCloud reduced to what the path reads: label,
 *  timestamp and the SoA float buffers of its CPointsMap. */
struct DeviceCloud;
struct CObservation
{
    using Ptr = std::shared_ptr<CObservation>;
    std::string        sensorLabel;
    double             timestamp = 0;  // seconds
    std::vector<float> x, y, z;
    /** set by LidarOdometry::onNewObservation when uploads are prefetched: the scan's device cloud, being
     *  uploaded and indexed on its own stream while the previous scan is still being registered */
    std::shared_future<std::shared_ptr<DeviceCloud>> prefetched;
    /** non-null => coordinates are already in pinned / device-visible host memory */
    const float *px = nullptr, *py = nullptr, *pz = nullptr;
    size_t       n = 0;
    size_t       size() const { return px ? n : x.size(); }
    const float* xs() const { return px ? px : x.data(); }
    const float* ys() const { return py ? py : y.data(); }
    const float* zs() const { return pz ? pz : z.data(); }
};

/** One point layer of mp2p_icp::metric_map_t, resident on the device.
 *
 *  Key-frame clouds outlive the scan that produced them (they sit in the world
 *  model, LidarOdometry.cpp:384-388, and are fetched again for extra edges,
 *  cpp:658-666).  So that a long session does not run out of HBM, a cloud can be
 *  SPILLED: its coordinates are downloaded to host memory and the device copy
 *  (points + search index) is freed; pin() brings it back -- same points, same
 *  index, same registration results -- and keeps it resident until unpin(). */
struct DeviceCloud
{
    using Ptr = std::shared_ptr<DeviceCloud>;
    b200icp_cloud_t* h = nullptr;  // null while spilled
    explicit DeviceCloud(b200icp_cloud_t* c, b200icp_t* owner = nullptr, float search_radius = 0.f)
        : h(c), owner_(owner), radius_(search_radius), n_(b200icp_cloud_size(c))
    {
        last_use_ = ++clock();  // a new cloud is the most recently used one
    }
    DeviceCloud(const DeviceCloud&) = delete;
    DeviceCloud& operator=(const DeviceCloud&) = delete;
    ~DeviceCloud() { b200icp_cloud_free(h); }
    size_t size() const { return n_; }
    bool   empty() const { return n_ == 0; }

    /** HBM held right now (0 while spilled) */
    size_t device_bytes() const
    {
        std::lock_guard<std::mutex> lk(m_);
        return h ? b200icp_cloud_device_bytes(h) : 0;
    }
    bool resident() const
    {
        std::lock_guard<std::mutex> lk(m_);
        return h != nullptr;
    }
    uint64_t last_use() const { return last_use_.load(); }
    /** handle for a registration: re-uploads a spilled cloud; nullptr on failure */
    b200icp_cloud_t* pin()
    {
        std::lock_guard<std::mutex> lk(m_);
        if (!h && owner_)
        {
            b200icp_cloud_t* c = nullptr;
            if (b200icp_cloud_upload(owner_, sx_.data(), sy_.data(), sz_.data(), n_, radius_, &c) != B200ICP_OK)
                return nullptr;
            h = c;
            sx_ = std::vector<float>(), sy_ = std::vector<float>(), sz_ = std::vector<float>();
            reloads()++;
        }
        pins_++;
        last_use_ = ++clock();
        return h;
    }
    void unpin()
    {
        std::lock_guard<std::mutex> lk(m_);
        if (pins_ > 0) pins_--;
    }
    /** frees the device copy if nobody is using it; false when pinned, spilled already or not spillable */
    bool spill()
    {
        std::lock_guard<std::mutex> lk(m_);
        if (!h || pins_ > 0 || !owner_) return false;
        std::vector<float> x(n_), y(n_), z(n_);
        if (n_ && b200icp_cloud_download(h, x.data(), y.data(), z.data()) != B200ICP_OK) return false;
        b200icp_cloud_free(h);
        h = nullptr;
        sx_.swap(x), sy_.swap(y), sz_.swap(z);
        spills()++;
        return true;
    }
    static std::atomic<uint64_t>& spills()
    {
        static std::atomic<uint64_t> v{0};
        return v;
    }
    static std::atomic<uint64_t>& reloads()
    {
        static std::atomic<uint64_t> v{0};
        return v;
    }

   private:
    static std::atomic<uint64_t>& clock()
    {
        static std::atomic<uint64_t> v{0};
        return v;
    }
    mutable std::mutex    m_;
    b200icp_t*            owner_ = nullptr;
    float                 radius_ = 0.f;
    size_t                n_ = 0;
    int                   pins_ = 0;
    std::atomic<uint64_t> last_use_{0};
    std::vector<float>    sx_, sy_, sz_;  // coordinates while spilled
};

/** RAII use of a cloud by one registration */
struct CloudPin
{
    DeviceCloud::Ptr c;
    b200icp_cloud_t* h = nullptr;
    explicit CloudPin(DeviceCloud::Ptr cloud) : c(std::move(cloud)), h(c ? c->pin() : nullptr) {}
    CloudPin(const CloudPin&) = delete;
    CloudPin& operator=(const CloudPin&) = delete;
    ~CloudPin()
    {
        if (c) c->unpin();
    }
};

struct FactorRelativePose3
{
    id_t    from_kf = INVALID_ID, to_kf = INVALID_ID;
    TPose3D rel_pose;
    double  noise_model_diag_xyz_ = 0.10, noise_model_diag_rot_ = 0.0174532925199;
};
using Factor = FactorRelativePose3;

class BackEndBase
{
   public:
    virtual ~BackEndBase() = default;
    struct ProposeKF_Input
    {
        double timestamp = 0;
    };
    struct ProposeKF_Output
    {
        bool                success = false;
        std::optional<id_t> new_kf_id;
    };
    struct AddFactor_Output
    {
        bool                 success = false;
        std::optional<fid_t> new_factor_id;
    };
    struct AdvertiseUpdatedLocalization_Input
    {
        double  timestamp = 0;
        id_t    reference_kf = INVALID_ID;
        TPose3D pose;
    };
    virtual std::future<ProposeKF_Output> addKeyFrame(const ProposeKF_Input& i) = 0;
    virtual std::future<AddFactor_Output> addFactor(Factor& f) = 0;
    virtual std::future<void> advertiseUpdatedLocalization(const AdvertiseUpdatedLocalization_Input& l) = 0;
};

/** Entities (keyframes) with annotations and the factor adjacency. */
class WorldModel
{
   public:
    using Ptr = std::shared_ptr<WorldModel>;
    void entities_lock_for_write() { ent_mtx_.lock(); }
    void entities_unlock_for_write() { ent_mtx_.unlock(); }
    void entities_lock_for_read() { ent_mtx_.lock_shared(); }
    void entities_unlock_for_read() { ent_mtx_.unlock_shared(); }
    void factors_lock_for_read() { fac_mtx_.lock_shared(); }
    void factors_unlock_for_read() { fac_mtx_.unlock_shared(); }

    std::map<std::string, DeviceCloud::Ptr>& entity_annotations_by_id(id_t id) { return annotations_[id]; }
    std::set<id_t> entity_neighbors(id_t id) const
    {
        std::lock_guard<std::mutex> lk(adj_mtx_);
        auto                        it = adjacency_.find(id);
        return it == adjacency_.end() ? std::set<id_t>() : it->second;
    }
    void add_edge(id_t a, id_t b)
    {
        std::lock_guard<std::mutex> lk(adj_mtx_);
        adjacency_[a].insert(b);
        adjacency_[b].insert(a);
    }
    void clear()
    {
        annotations_.clear();
        adjacency_.clear();
    }
    /** drops one named annotation from every entity (a module removing what IT put there) */
    void erase_annotation_everywhere(const std::string& name)
    {
        ent_mtx_.lock();
        for (auto& kv : annotations_) kv.second.erase(name);
        ent_mtx_.unlock();
    }

   private:
    std::shared_mutex                                          ent_mtx_, fac_mtx_;
    mutable std::mutex                                         adj_mtx_;
    std::map<id_t, std::map<std::string, DeviceCloud::Ptr>>    annotations_;
    std::map<id_t, std::set<id_t>>                             adjacency_;
};

/** In-process back-end: hands out KF / factor ids and records what it was told. */
class SimpleBackEnd : public BackEndBase
{
   public:
    explicit SimpleBackEnd(WorldModel::Ptr wm = nullptr) : wm_(std::move(wm)) {}
    std::future<ProposeKF_Output> addKeyFrame(const ProposeKF_Input& i) override
    {
        std::lock_guard<std::mutex> lk(mtx_);
        ProposeKF_Output            o;
        o.success = true;
        o.new_kf_id = next_kf_++;
        kf_stamps.push_back(i.timestamp);
        std::promise<ProposeKF_Output> p;
        p.set_value(o);
        return p.get_future();
    }
    std::future<AddFactor_Output> addFactor(Factor& f) override
    {
        std::lock_guard<std::mutex> lk(mtx_);
        AddFactor_Output            o;
        o.success = true;
        o.new_factor_id = factors.size();
        factors.push_back(f);
        if (wm_) wm_->add_edge(f.from_kf, f.to_kf);
        std::promise<AddFactor_Output> p;
        p.set_value(o);
        return p.get_future();
    }
    std::future<void> advertiseUpdatedLocalization(const AdvertiseUpdatedLocalization_Input& l) override
    {
        std::lock_guard<std::mutex> lk(mtx_);
        localizations.push_back(l);
        std::promise<void> p;
        p.set_value();
        return p.get_future();
    }
    std::mutex                                      mtx_;
    std::vector<double>                             kf_stamps;
    std::vector<Factor>                             factors;
    std::vector<AdvertiseUpdatedLocalization_Input> localizations;

   private:
    WorldModel::Ptr wm_;
    id_t            next_kf_ = 0;
};

/** mola::FrontEndBase / ExecutableBase: lifecycle + the two inherited members
 *  the reference reads (raw_sensor_label_, slam_backend_). */
class FrontEndBase
{
   public:
    virtual ~FrontEndBase() = default;
    /** [U] reads `raw_sensor_label` (and finds the back-end) before initialize() */
    void initialize_common(const Yaml& cfg)
    {
        if (cfg.has("raw_sensor_label")) raw_sensor_label_ = cfg["raw_sensor_label"].as_string();
    }
    virtual void initialize(const Yaml& cfg) = 0;
    virtual void spinOnce() = 0;
    virtual void onNewObservation(CObservation::Ptr& o) = 0;

    std::string                  raw_sensor_label_ = "lidar";
    std::shared_ptr<BackEndBase> slam_backend_;
};

}  // namespace mola
