// ICP_B200.h -- mp2p_icp::ICP whose align() runs on a B200 through libb200icp.so.
//
// Drop-in at the reference's ICP seam: LidarOdometry builds its ICP objects by class name
// (`icp_class`, /root/reference/src/LidarOdometry.cpp:62-75), configures them with
// initialize_solvers / _matchers / _quality_evaluators (cpp:80-87) and calls
// icp->align(from, to, guess, params, result) (cpp:869-871).  Selecting
//     icp_class: 'mola::ICP_B200'
// in params/icp-settings-*.yaml routes that call to the device; nothing else changes.
//
// The object keeps the solver / matcher / quality lists the base class builds from the YAML and
// translates them (dynamic_cast on the stock classes, reading their public parameters) into a
// b200icp_params_t; combinations the device path does not implement make align() throw, naming the
// class, like an unknown icp_class does at cpp:70-75.
#pragma once
#include <b200icp.h>
#include <mola_b200/DeviceCloudCache.h>
#include <mp2p_icp/ICP.h>

#include <mutex>
#include <vector>

namespace mola
{
class ICP_B200 : public mp2p_icp::ICP
{
    DEFINE_MRPT_OBJECT(ICP_B200, mola)

   public:
    ICP_B200() = default;
    ~ICP_B200() override;

    /** pc1 = reference cloud ("from"), pc2 = the cloud that moves ("to"); guess and result = pose of pc2 wrt pc1.
     *  Fills quality, optimal_tf {mean, cov}, nIterations, terminationReason -- the fields read at cpp:873-888.
     *  Re-entrant: called concurrently from both thread pools on one shared object (h:167-172, cpp:711-729). */
    void align(const mp2p_icp::metric_map_t& pc1, const mp2p_icp::metric_map_t& pc2,
               const mrpt::math::TPose3D& init_guess_m2_wrt_m1, const mp2p_icp::Parameters& p,
               mp2p_icp::Results& result) override;

    /** CUDA device of this object (default 0; set before the first align). */
    void setDevice(int device) { device_ = device; }
    /** Point layer registered (default metric_map_t::PT_LAYER_RAW). */
    void setLayer(const std::string& name) { layer_ = name; }

    /** The translation of (Parameters, solvers, matchers, quality evaluators) into the C ABI's parameter block;
     *  throws std::runtime_error naming what is not supported.  Public for tests. */
    b200icp_params_t translate(const mp2p_icp::Parameters& p) const;

    std::size_t cachedClouds() const { return clouds_.size(); }
    std::size_t contexts() const { return ctxs_.size(); }
    std::size_t uploads() const { return clouds_.uploads(); }

   private:
    b200icp_t* context_for(const b200icp_params_t& q);

    struct Ctx
    {
        b200icp_params_t q;
        b200icp_t*       h;
    };
    std::mutex                   mtx_;
    std::vector<Ctx>             ctxs_;  // one device context per distinct solver / matcher / quality configuration
    mola_b200::DeviceCloudCache  clouds_;
    int                          device_ = 0;
    std::string                  layer_  = mp2p_icp::metric_map_t::PT_LAYER_RAW;
};
}  // namespace mola
