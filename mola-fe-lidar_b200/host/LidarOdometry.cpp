// LidarOdometry.cpp -- host-side restatement of the control flow of
// mola::LidarOdometry (reference: src/LidarOdometry.cpp), calling the CUDA
// path through the C ABI of include/b200icp.h.  Each function cites the
// reference lines it follows; quirks are kept.
#include "LidarOdometry.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <random>
#include <stdexcept>
#include <thread>

namespace mola
{
static const std::string ANNOTATION_NAME_PC_LAYERS = "lidar-pointcloud-layers";  // cpp:41

static inline double DEG2RAD(double d) { return d * 3.14159265358979323846 / 180.0; }

static void check_rc(int rc, const char* what)
{
    if (rc != B200ICP_OK) throw std::runtime_error(std::string(what) + ": " + b200icp_last_error());
}

// ---------------------------------------------------------------- pose graph
void NetworkOfPoses3D::getAdjacencyMatrix(std::map<id_t, std::set<id_t>>& adj) const
{
    adj.clear();
    for (const auto& e : edges)
    {
        adj[e.first.first].insert(e.first.second);
        adj[e.first.second].insert(e.first.first);
    }
}

void NetworkOfPoses3D::dijkstra_nodes_estimate(std::map<id_t, size_t>& topo)
{
    // unit edge weights: breadth-first spanning tree from the root, poses
    // composed along the tree (edge (a,b) holds the pose of b wrt a)
    std::map<id_t, std::set<id_t>> adj;
    getAdjacencyMatrix(adj);
    nodes.clear();
    topo.clear();
    if (root == INVALID_ID) return;
    nodes[root] = pose_identity();
    topo[root] = 0;
    std::vector<id_t> frontier{root};
    while (!frontier.empty())
    {
        std::vector<id_t> next;
        for (const id_t a : frontier)
            for (const id_t b : adj[a])
            {
                if (nodes.count(b)) continue;
                CPose3D    rel;
                const auto fwd = edges.find({a, b});
                if (fwd != edges.end())
                    rel = fwd->second;
                else
                {  // traversed against its direction: invert
                    const CPose3D& e = edges.at({b, a});
                    b2::pose_inverse_compose(e, pose_identity(), rel);
                }
                CPose3D pb;
                b2::pose_compose(nodes[a], rel, pb);
                nodes[b] = pb;
                topo[b] = topo[a] + 1;
                next.push_back(b);
            }
        frontier.swap(next);
    }
}

// ------------------------------------------------------------------ lifecycle
LidarOdometry::LidarOdometry() = default;

LidarOdometry::~LidarOdometry()
{
    worker_pool_prefetch_.clear();
    worker_pool_.clear();
    worker_pool_past_KFs_.clear();
    release_icp_objects();
}

// Every device cloud this module made holds its ICP object's context (uploads, spills and reloads go through
// it): they go first -- the last scan, the key-frame clouds annotated in the world model (only THIS module's
// annotation; a world model injected with setWorldModel() may be shared with other modules) -- and only then
// the ICP objects.  Callers make sure no pool task is running (destructor: pools cleared; initialize: waitIdle).
void LidarOdometry::release_icp_objects()
{
    {
        std::lock_guard<std::mutex> lk(state_mtx_);
        state_ = MethodState();
    }
    if (worldmodel_) worldmodel_->erase_annotation_everywhere(ANNOTATION_NAME_PC_LAYERS);
    {
        std::lock_guard<std::mutex> lk(kf_store_mtx_);
        kf_store_.clear();
    }
    for (auto& kv : params_.icp)
    {
        b200icp_destroy(kv.second.icp);
        kv.second.icp = nullptr;
    }
    params_.icp.clear();
}

// cpp:57-88
static void load_icp_set_of_params(LidarOdometry::Parameters::ICP_case& out, const Yaml& cfg, int device)
{
    if (!cfg.isMap()) throw std::runtime_error("ICP settings block is missing or not a map");
    // icp_class / params / solvers / matchers / quality: the C ABI re-checks
    // the same entries (ENSURE_YAML_ENTRY_EXISTS cpp:77-86) and class names
    check_rc(b200icp_params_from_yaml(cfg.dump().c_str(), &out.icpParameters), "load_icp_set_of_params");
    check_rc(b200icp_create(&out.icpParameters, device, &out.icp), "b200icp_create");
}

// cpp:90-149
void LidarOdometry::initialize(const Yaml& c)
{
    auto numICPThreads = std::thread::hardware_concurrency() / 2;
    if (numICPThreads < 2) numICPThreads = 2;
    worker_pool_past_KFs_.resize(numICPThreads);  // cpp:94-96

    if (!c.has("params")) throw std::runtime_error("yaml: required entry `params` not found");
    const Yaml& cfg = c["params"];  // cpp:102

    cfg.load_req("min_dist_xyz_between_keyframes", params_.min_dist_xyz_between_keyframes);  // cpp:105
    if (cfg.has("min_rotation_between_keyframes"))  // YAML_LOAD_OPT_DEG cpp:106
        params_.min_rotation_between_keyframes = DEG2RAD(cfg["min_rotation_between_keyframes"].as_double());
    cfg.load_opt("min_time_between_scans", params_.min_time_between_scans);
    cfg.load_opt("min_icp_goodness", params_.min_icp_goodness);
    cfg.load_opt("min_icp_goodness_lc", params_.min_icp_goodness_lc);
    cfg.load_opt("min_dist_to_matching", params_.min_dist_to_matching);
    cfg.load_opt("max_dist_to_matching", params_.max_dist_to_matching);
    cfg.load_opt("max_dist_to_loop_closure", params_.max_dist_to_loop_closure);
    cfg.load_opt("max_nearby_align_checks", params_.max_nearby_align_checks);
    cfg.load_opt("min_topo_dist_to_consider_loopclosure", params_.min_topo_dist_to_consider_loopclosure);
    cfg.load_opt("loop_closure_montecarlo_samples", params_.loop_closure_montecarlo_samples);
    cfg.load_opt("viz_decor_decimation", params_.viz_decor_decimation);
    cfg.load_opt("viz_decor_pointsize", params_.viz_decor_pointsize);
    // additive keys of this implementation
    cfg.load_opt("b200_device", params_.device);
    cfg.load_opt("b200_extra_edge_checks", params_.extra_edge_checks);
    cfg.load_opt("b200_kf_store_budget_mb", params_.kf_store_budget_mb);
    cfg.load_opt("b200_prefetch_uploads", params_.prefetch_uploads);
    {
        unsigned int seed = (unsigned int)params_.montecarlo_seed;
        cfg.load_opt("b200_montecarlo_seed", seed);
        params_.montecarlo_seed = seed;
    }

    waitIdle();  // a second initialize(): nothing may still be using the objects released next
    release_icp_objects();
    cfg.at("icp_settings_with_vel");  // ENSURE_YAML_ENTRY_EXISTS cpp:122
    load_icp_set_of_params(params_.icp[AlignKind::LidarOdometry], cfg["icp_settings_with_vel"], params_.device);
    load_icp_set_of_params(params_.icp[AlignKind::NearbyAlign], cfg["icp_settings_without_vel"], params_.device);
    load_icp_set_of_params(params_.icp[AlignKind::LoopClosure], cfg["icp_settings_loop_closure"], params_.device);

    // clouds are indexed once for the largest search radius any case uses
    cloud_search_radius_ = 0.f;
    for (const auto& kv : params_.icp)
        cloud_search_radius_ = std::max(cloud_search_radius_, (float)kv.second.icpParameters.distance_threshold);

    // cpp:135-140: generators + filter pipeline. The shipped YAML defines
    // neither: absent => default generator, empty pipeline.
    const Yaml& gen = cfg["pointcloud_generator"];
    if (!gen.isNull())
    {
        if (!gen.isSeq()) throw std::runtime_error("pointcloud_generator must be a sequence");
        for (const auto& g : gen.seq)
        {
            const std::string cls = g.at("class_name").as_string();
            if (cls != "mp2p_icp_filters::Generator")
                throw std::runtime_error("pointcloud_generator class_name=`" + cls + "` is not registered");
        }
    }
    params_.voxel_decimation_resolution = 0.0;
    params_.edges_planes_enabled = false;
    const Yaml& flt = cfg["pointcloud_filter"];
    if (!flt.isNull())
    {
        if (!flt.isSeq()) throw std::runtime_error("pointcloud_filter must be a sequence");
        for (const auto& f : flt.seq)
        {
            const std::string cls = f.at("class_name").as_string();
            const Yaml&       p = f.at("params");
            if (cls == "mp2p_icp_filters::FilterDecimateVoxels")
            {
                p.load_req("voxel_filter_resolution", params_.voxel_decimation_resolution);
                p.load_opt("use_voxel_average", params_.voxel_use_average);
            }
            else if (cls == "mp2p_icp_filters::FilterEdgesPlanes" || cls == "mola::lidar_segmentation::FilterEdgesPlanes")
            {   // keys of pointcloud_filter_params (kitti-default.yaml:23-32); defaults = the shipped values
                auto& e = params_.edges_planes;
                b200icp_edges_planes_defaults(&e);
                p.load_opt("voxel_filter_resolution", e.voxel_filter_resolution);
                p.load_opt("full_pointcloud_decimation", e.full_pointcloud_decimation);
                p.load_opt("voxel_filter_decimation", e.voxel_filter_decimation);
                p.load_opt("voxel_filter_max_e2_e0", e.voxel_filter_max_e2_e0);
                p.load_opt("voxel_filter_max_e1_e0", e.voxel_filter_max_e1_e0);
                p.load_opt("voxel_filter_min_e2_e0", e.voxel_filter_min_e2_e0);
                p.load_opt("voxel_filter_min_e1_e0", e.voxel_filter_min_e1_e0);
                p.load_opt("b200_min_points_per_voxel", e.min_points_per_voxel);
                std::string layer = "planes";
                p.load_opt("b200_register_layer", layer);
                if (layer == "edges") params_.edges_planes_layer = 0;
                else if (layer == "planes") params_.edges_planes_layer = 1;
                else if (layer == "full_decim") params_.edges_planes_layer = 2;
                else throw std::runtime_error("b200_register_layer must be edges, planes or full_decim");
                params_.edges_planes_enabled = true;
            }
            else
                throw std::runtime_error("pointcloud_filter class_name=`" + cls +
                                         "` is not registered (known: mp2p_icp_filters::FilterDecimateVoxels, "
                                         "mp2p_icp_filters::FilterEdgesPlanes)");
        }
    }
    // attach to world model, if present (cpp:144-146): the harness may have
    // injected one; else keep a private one so that KF clouds have a home
    if (!worldmodel_) worldmodel_ = std::make_shared<WorldModel>();
}

void LidarOdometry::spinOnce() { ProfilerEntry tleg(profiler_, "spinOnce"); }  // cpp:150-158

void LidarOdometry::reset()  // cpp:160
{
    worker_pool_prefetch_.waitIdle();
    prefetch_last_tim_ = -1.0;
    std::lock_guard<std::mutex> lk(state_mtx_);
    state_ = MethodState();
}

void LidarOdometry::waitIdle()
{
    worker_pool_prefetch_.waitIdle();
    worker_pool_.waitIdle();
    worker_pool_past_KFs_.waitIdle();
}

// cpp:162-187
void LidarOdometry::onNewObservation(CObservation::Ptr& o)
{
    ProfilerEntry tleg(profiler_, "onNewObservation");
    if (!o) throw std::runtime_error("onNewObservation: null observation");  // ASSERT_(o)
    if (o->sensorLabel != raw_sensor_label_) return;  // only "my" sensor source

    const auto queued = worker_pool_.pendingTasks();
    profiler_.registerUserMeasure("onNewObservation.queue_length", (double)queued);
    if (queued > 10)
    {
        profiler_.registerUserMeasure("onNewObservation.drop_observation", 1);
        std::lock_guard<std::mutex> lk(state_mtx_);
        state_.n_dropped++;
        return;
    }
    profiler_.enter("delay_onNewObs_to_process");
    CObservation::Ptr obs = o;
    if (params_.prefetch_uploads)
    {
        // stage 1 on its own thread and CUDA stream: the scan is on the device, indexed, by the time the
        // 1-thread pool gets to it.  The time gate of cpp:201-212 is a function of the timestamps alone, so the
        // prefetch thread replays it and does not upload scans that will be skipped.
        auto prom = std::make_shared<std::promise<DeviceCloud::Ptr>>();
        obs->prefetched = prom->get_future().share();
        worker_pool_prefetch_.enqueue([this, obs, prom]() {
            try
            {
                const double t = obs->timestamp;
                if (prefetch_last_tim_ >= 0 && (t - prefetch_last_tim_) < params_.min_time_between_scans)
                {
                    prom->set_value(nullptr);
                    return;
                }
                prefetch_last_tim_ = t;
                prom->set_value(make_cloud(*obs));
            }
            catch (...)
            {
                prom->set_exception(std::current_exception());
            }
        });
    }
    worker_pool_.enqueue([this, obs]() mutable { doProcessNewObservation(obs); });
}

// apply_generators + apply_filter_pipeline (cpp:215-224) on the device
DeviceCloud::Ptr LidarOdometry::make_cloud(const CObservation& o)
{
    b200icp_t*       ctx = params_.icp.at(AlignKind::LidarOdometry).icp;
    b200icp_cloud_t* raw = nullptr;
    const bool       filtered = params_.voxel_decimation_resolution > 0 || params_.edges_planes_enabled;
    {   // observation -> device cloud (apply_generators, cpp:215-217); the search index is built for the cloud
        // that gets registered: this one, or the output of the filter stage below
        ProfilerEntry tle0(profiler_, "doProcessNewObservation.0.upload_and_index");
        if (filtered)
            check_rc(b200icp_cloud_upload_raw(ctx, o.xs(), o.ys(), o.zs(), o.size(), &raw), "b200icp_cloud_upload_raw");
        else
            check_rc(b200icp_cloud_upload(ctx, o.xs(), o.ys(), o.zs(), o.size(), cloud_search_radius_, &raw),
                     "b200icp_cloud_upload");
    }
    auto              raw_ptr = std::make_shared<DeviceCloud>(raw, ctx, cloud_search_radius_);
    ProfilerEntry     tle1(profiler_, "doProcessNewObservation.1.filter_pointclouds");
    if (params_.edges_planes_enabled)
    {   // FilterEdgesPlanes: three layers on the device; the configured one is what gets registered
        b200icp_cloud_t* layers[3] = {nullptr, nullptr, nullptr};
        check_rc(b200icp_filter_edges_planes(ctx, raw, &params_.edges_planes, cloud_search_radius_, layers, nullptr,
                                             nullptr),
                 "b200icp_filter_edges_planes");
        for (int l = 0; l < 3; l++)
            if (l != params_.edges_planes_layer) b200icp_cloud_free(layers[l]);
        return std::make_shared<DeviceCloud>(layers[params_.edges_planes_layer], ctx, cloud_search_radius_);
    }
    if (filtered)
    {
        b200icp_cloud_t* dec = nullptr;
        check_rc(b200icp_voxel_decimate(ctx, raw, (float)params_.voxel_decimation_resolution,
                                        params_.voxel_use_average ? 1 : 0, cloud_search_radius_, &dec, nullptr),
                 "b200icp_voxel_decimate");
        return std::make_shared<DeviceCloud>(dec, ctx, cloud_search_radius_);
    }
    return raw_ptr;
}

// cpp:190-514
void LidarOdometry::doProcessNewObservation(CObservation::Ptr& o)
{
    try
    {
        if (!o) throw std::runtime_error("doProcessNewObservation: null observation");
        ProfilerEntry tleg(profiler_, "doProcessNewObservation");
        profiler_.leave("delay_onNewObs_to_process");

        // Only process pointclouds that are sufficiently apart in time (cpp:201-212)
        const double this_obs_tim = o->timestamp;
        if (state_.last_obs_tim >= 0 && (this_obs_tim - state_.last_obs_tim) < params_.min_time_between_scans)
        {
            state_.n_dropped++;
            return;
        }

        // Extract points from observation + filter/segment (cpp:214-226): already under way on the prefetch
        // thread when the scan came through onNewObservation
        DeviceCloud::Ptr this_obs_points;
        if (o->prefetched.valid()) this_obs_points = o->prefetched.get();
        if (!this_obs_points) this_obs_points = make_cloud(*o);

        profiler_.enter("doProcessNewObservation.2.copy_vars");
        // Store for next step (cpp:230-234)
        auto last_obs_tim = state_.last_obs_tim;
        auto last_points = state_.last_points;
        state_.last_obs_tim = this_obs_tim;
        state_.last_points = this_obs_points;
        profiler_.leave("doProcessNewObservation.2.copy_vars");

        if (this_obs_points->empty()) return;  // cpp:238-245
        state_.n_processed++;

        bool create_keyframe = false;
        // First time we cannot do ICP since we need at least two pointclouds (cpp:249-257)
        if (!last_points || last_points->empty())
            create_keyframe = true;
        else
        {
            profiler_.enter("doProcessNewObservation.2c.prepare_icp_in");
            // Use velocity model for the initial guess (cpp:265-275)
            double dt = .0;
            if (last_obs_tim >= 0) dt = this_obs_tim - last_obs_tim;

            ICP_Output icp_out;
            ICP_Input  icp_in;
            icp_in.init_guess_to_wrt_from =
                TPose3D{state_.last_iter_twist.vx * dt, state_.last_iter_twist.vy * dt,
                        state_.last_iter_twist.vz * dt, state_.last_iter_twist.wz * dt, 0, 0};
            icp_in.to_pc = this_obs_points;
            icp_in.from_pc = last_points;
            icp_in.from_id = state_.last_kf;
            icp_in.to_id = INVALID_ID;  // current data, not a new KF (yet)
            icp_in.debug_str = "lidar_odom";
            // If we don't have a valid twist estimation, use the other set (cpp:287-290)
            b200icp_call_params_of(state_.last_iter_twist_is_good
                                       ? &params_.icp[AlignKind::LidarOdometry].icpParameters
                                       : &params_.icp[AlignKind::NearbyAlign].icpParameters,
                                   &icp_in.icp_params);
            profiler_.leave("doProcessNewObservation.2c.prepare_icp_in");
            {
                ProfilerEntry tle(profiler_, "doProcessNewObservation.3.icp_latest");
                run_one_icp(icp_in, icp_out);  // cpp:299
            }
            const CPose3D rel_pose = icp_out.found_pose_to_wrt_from.getMeanVal();
            const TPose3D rp = asTPose(rel_pose);
            // Update velocity model (cpp:305-311)
            state_.last_iter_twist.vx = rp.x / dt;
            state_.last_iter_twist.vy = rp.y / dt;
            state_.last_iter_twist.vz = rp.z / dt;
            state_.last_iter_twist.wz = rp.yaw / dt;
            state_.last_iter_twist_is_good = true;
            state_.last_icp_out = icp_out;

            // Create a new KF if the distance since the last one is large enough (cpp:321-337)
            CPose3D acc;
            b2::pose_compose(state_.accum_since_last_kf, rel_pose, acc);
            state_.accum_since_last_kf = acc;
            const double dist_eucl_since_last = pose_norm(state_.accum_since_last_kf);
            double       lg[6];
            b2::se3_log(state_.accum_since_last_kf, lg);
            const double rot_since_last = std::sqrt(lg[3] * lg[3] + lg[4] * lg[4] + lg[5] * lg[5]);
            create_keyframe = (icp_out.goodness > params_.min_icp_goodness &&
                               (dist_eucl_since_last > params_.min_dist_xyz_between_keyframes ||
                                rot_since_last > params_.min_rotation_between_keyframes));
        }

        if (create_keyframe && slam_backend_)
        {
            // 1) New KeyFrame (cpp:345-370)
            BackEndBase::ProposeKF_Input kf;
            kf.timestamp = this_obs_tim;
            profiler_.enter("doProcessNewObservation.3a.addKeyFrame");
            auto kf_out = slam_backend_->addKeyFrame(kf).get();
            if (!kf_out.success || !kf_out.new_kf_id) throw std::runtime_error("addKeyFrame failed");
            const id_t new_kf_id = kf_out.new_kf_id.value();
            if (new_kf_id == INVALID_ID) throw std::runtime_error("addKeyFrame: invalid id");
            profiler_.leave("doProcessNewObservation.3a.addKeyFrame");

            // Add point cloud to the KF annotations in the map (cpp:372-389);
            // the render decorations of cpp:390-426 are dropped
            if (!worldmodel_) throw std::runtime_error("no WorldModel");
            {
                worldmodel_->entities_lock_for_write();
                ProfilerEntry tle(profiler_, "doProcessNewObservation.4.writePCsToWorldModel");
                worldmodel_->entity_annotations_by_id(new_kf_id).emplace(ANNOTATION_NAME_PC_LAYERS, this_obs_points);
                worldmodel_->entities_unlock_for_write();
            }
            kf_store_add(this_obs_points);
            // 2) New SE(3) constraint between consecutive Keyframes (cpp:432-470)
            if (state_.last_kf != INVALID_ID)
            {
                FactorRelativePose3 fPose3;
                fPose3.from_kf = state_.last_kf, fPose3.to_kf = new_kf_id;
                fPose3.rel_pose = asTPose(state_.accum_since_last_kf);
                fPose3.noise_model_diag_xyz_ = 0.10;
                fPose3.noise_model_diag_rot_ = DEG2RAD(1.0);
                auto factor_out = slam_backend_->addFactor(fPose3).get();
                if (!factor_out.success || !factor_out.new_factor_id ||
                    factor_out.new_factor_id == INVALID_FID)
                    throw std::runtime_error("addFactor failed");
                {
                    std::lock_guard<std::mutex> lck(local_pose_graph_mtx);
                    state_.local_pose_graph.graph.insertEdgeAtEnd(state_.last_kf, new_kf_id,
                                                                  state_.accum_since_last_kf);
                }
            }
            // Reset accumulators (cpp:472-474)
            state_.accum_since_last_kf = pose_identity();
            state_.last_kf = new_kf_id;
        }

        // publish our **current** vehicle pose (cpp:477-491)
        if (slam_backend_)
        {
            ProfilerEntry tle(profiler_, "doProcessNewObservation.5.advertiseUpdatedLocalization");
            BackEndBase::AdvertiseUpdatedLocalization_Input new_loc;
            new_loc.timestamp = this_obs_tim;
            new_loc.reference_kf = state_.last_kf;
            new_loc.pose = asTPose(state_.accum_since_last_kf);
            slam_backend_->advertiseUpdatedLocalization(new_loc);
        }

        // try to align this new KF against a few past KFs as well (cpp:493-508)
        bool can_check_for_other_matches = true;
        {
            std::lock_guard<std::mutex> lck(local_pose_graph_mtx);
            can_check_for_other_matches = !state_.local_pose_graph.graph.edges.empty();
        }
        if (can_check_for_other_matches && params_.extra_edge_checks)
        {
            ProfilerEntry tle(profiler_, "doProcessNewObservation.6.checkForNearbyKFs");
            checkForNearbyKFs();
        }
    }
    catch (const std::exception& e)
    {
        profiler_.registerUserMeasure(std::string("exception: ") + e.what(), 1);  // cpp:510-513
    }
}

// cpp:516-744
void LidarOdometry::checkForNearbyKFs()
{
    using euclidean_dist_t = double;
    std::map<euclidean_dist_t, std::pair<id_t, topological_dist_t>> KF_distances;
    id_t current_kf_id{INVALID_ID};
    {
        std::lock_guard<std::mutex> lck(local_pose_graph_mtx);
        auto& lpg = state_.local_pose_graph.graph;
        current_kf_id = state_.last_kf;
        // Dijkstra from the current KF: spanning tree -> relative poses and
        // topological distances to all other nodes (cpp:537-542)
        lpg.root = current_kf_id;
        std::map<id_t, size_t> topolog_dists;
        lpg.dijkstra_nodes_estimate(topolog_dists);
        // Sort KFs by distance (cpp:545-552)
        for (const auto& kfs : lpg.nodes)
            KF_distances[pose_norm(kfs.second)] = std::make_pair(kfs.first, topolog_dists.at(kfs.first));
        std::map<id_t, std::set<id_t>> adj;
        lpg.getAdjacencyMatrix(adj);
        // Remove too distant KFs (cpp:558-569)
        while (lpg.nodes.size() > params_.max_KFs_local_graph && !KF_distances.empty())
        {
            const auto id_to_remove = KF_distances.rbegin()->second.first;
            KF_distances.erase(std::prev(KF_distances.end()));
            lpg.nodes.erase(id_to_remove);
            for (const auto other_id : adj[id_to_remove])
            {
                lpg.edges.erase(std::make_pair(id_to_remove, other_id));
                lpg.edges.erase(std::make_pair(other_id, id_to_remove));
            }
        }
    }
    // nodes at an intermediary distance (cpp:574-576)
    auto it1 = KF_distances.lower_bound(params_.min_dist_to_matching);
    auto it2 = KF_distances.upper_bound(std::max(params_.max_dist_to_loop_closure, params_.max_dist_to_matching));

    std::vector<ICP_Input::Ptr>                nearby_checks;
    std::map<euclidean_dist_t, ICP_Input::Ptr> loop_closure_checks;
    for (auto it = it1; it != it2; ++it)
    {
        const double             kf_eucl_dist = it->first;
        const auto               kf_id = it->second.first;
        const topological_dist_t kf_topo_d = it->second.second;
        bool                     edge_already_exists = false;
        const bool is_potential_loop_closure = (kf_topo_d >= params_.min_topo_dist_to_consider_loopclosure);
        // Only explore KFs farther than this threshold if they are LCs (cpp:591-594)
        if (!is_potential_loop_closure && kf_eucl_dist > params_.max_dist_to_matching) continue;
        // Already sent out for checking? (cpp:596-605)
        const auto pair_ids = std::make_pair(std::min(kf_id, current_kf_id), std::max(kf_id, current_kf_id));
        {
            std::lock_guard<std::mutex> lck(local_pose_graph_mtx);
            if (state_.local_pose_graph.checked_KF_pairs.count(pair_ids) != 0) edge_already_exists = true;
        }
        // cpp:610-631. Kept as in the reference: finding an existing factor
        // assigns `false` (a no-op), so such pairs are NOT discarded here.
        if (!edge_already_exists && worldmodel_)
        {
            worldmodel_->entities_lock_for_read();
            worldmodel_->factors_lock_for_read();
            const auto connected = worldmodel_->entity_neighbors(kf_id);
            if (connected.count(current_kf_id) != 0) edge_already_exists = false;
            worldmodel_->factors_unlock_for_read();
            worldmodel_->entities_unlock_for_read();
        }
        if (!edge_already_exists)
        {
            auto d = std::make_shared<ICP_Input>();
            d->to_id = kf_id;
            d->from_id = current_kf_id;
            // Retrieve the point clouds from the Map (cpp:640-669)
            worldmodel_->entities_lock_for_read();
            d->to_pc = worldmodel_->entity_annotations_by_id(d->to_id).at(ANNOTATION_NAME_PC_LAYERS);
            d->from_pc = worldmodel_->entity_annotations_by_id(d->from_id).at(ANNOTATION_NAME_PC_LAYERS);
            worldmodel_->entities_unlock_for_read();
            {
                std::lock_guard<std::mutex> lck(local_pose_graph_mtx);
                d->init_guess_to_wrt_from = asTPose(state_.local_pose_graph.graph.nodes[kf_id]);  // cpp:674-675
            }
            if (!is_potential_loop_closure)
            {
                d->align_kind = AlignKind::NearbyAlign;
                d->debug_str = "extra_edge";
                b200icp_call_params_of(&params_.icp[d->align_kind].icpParameters, &d->icp_params);
                nearby_checks.emplace_back(std::move(d));
            }
            else
            {
                d->align_kind = AlignKind::LoopClosure;
                d->debug_str = "loop_closure";
                b200icp_call_params_of(&params_.icp[d->align_kind].icpParameters, &d->icp_params);
                loop_closure_checks[kf_eucl_dist] = std::move(d);
            }
        }
    }
    // Nearby checks: send a maximum of "N" (cpp:703-722)
    const size_t nNearbyChecks = nearby_checks.size();
    const size_t nearbyCheckDecim =
        std::max(static_cast<size_t>(1U), nNearbyChecks / std::max(1u, params_.max_nearby_align_checks));
    for (size_t idx = 0; idx < nNearbyChecks; idx += nearbyCheckDecim)
    {
        const auto d = nearby_checks[idx];
        worker_pool_past_KFs_.enqueue([this, d]() { doCheckForNonAdjacentKFs(d); });
        std::lock_guard<std::mutex> lck(local_pose_graph_mtx);
        state_.local_pose_graph.checked_KF_pairs.insert(
            std::make_pair(std::min(d->to_id, d->from_id), std::max(d->to_id, d->from_id)));
    }
    // Loop closures: just send the one with the smallest distance (cpp:723-741)
    if (!loop_closure_checks.empty())
    {
        const auto d = loop_closure_checks.begin()->second;
        worker_pool_past_KFs_.enqueue([this, d]() { doCheckForNonAdjacentKFs(d); });
        std::lock_guard<std::mutex> lck(local_pose_graph_mtx);
        state_.local_pose_graph.checked_KF_pairs.insert(
            std::make_pair(std::min(d->to_id, d->from_id), std::max(d->to_id, d->from_id)));
    }
}

// cpp:746-849
void LidarOdometry::doCheckForNonAdjacentKFs(ICP_Input::Ptr d)
{
    try
    {
        ProfilerEntry tleg(profiler_, "doCheckForNonAdjacentKFs");
        ICP_Output    icp_out;
        if (d->align_kind != AlignKind::LoopClosure)
        {
            ProfilerEntry tle(profiler_, "doCheckForNonAdjacentKFs.run_icp");
            run_one_icp(*d, icp_out);  // cpp:759
        }
        else
        {
            ProfilerEntry tle(profiler_, "doCheckForNonAdjacentKFs.run_icp_loop_closure");
            // a small montecarlo sampling, keep the best attempt (cpp:767-787).
            // The N registrations are independent: one batched launch.
            const double  std_xyz = params_.max_dist_to_loop_closure * 0.1;
            const double  std_rot = DEG2RAD(2.0);
            const TPose3D original_guess = d->init_guess_to_wrt_from;
            const size_t  N = params_.loop_closure_montecarlo_samples;
            // the reference draws from an unseeded generator (cpp:773); here
            // the draws are reproducible per (seed, from, to)
            std::mt19937_64 rnd(params_.montecarlo_seed * 0x9E3779B97F4A7C15ull + d->from_id * 1000003ull + d->to_id);
            std::normal_distribution<double> gauss(0.0, 1.0);
            std::vector<double>              guesses(6 * N);
            for (size_t i = 0; i < N; i++)
            {
                guesses[6 * i + 0] = original_guess.x + gauss(rnd) * std_xyz;
                guesses[6 * i + 1] = original_guess.y + gauss(rnd) * std_xyz;
                guesses[6 * i + 2] = original_guess.z + gauss(rnd) * std_xyz;
                guesses[6 * i + 3] = original_guess.yaw + gauss(rnd) * std_rot;
                guesses[6 * i + 4] = original_guess.pitch;
                guesses[6 * i + 5] = original_guess.roll;
            }
            CloudPin pin_from(d->from_pc), pin_to(d->to_pc);  // resident (re-uploaded if spilled) for the batch
            if (!pin_from.h || !pin_to.h) throw std::runtime_error(std::string("cloud reload: ") + b200icp_last_error());
            std::vector<const b200icp_cloud_t*> fr(N, pin_from.h), to(N, pin_to.h);
            std::vector<b200icp_result_t>       res(N);
            check_rc(b200icp_align_batch(params_.icp.at(d->align_kind).icp, N, fr.data(), to.data(),
                                         guesses.data(), res.data()),
                     "b200icp_align_batch");
            count_icp(N);
            kf_store_enforce_budget();
            for (size_t i = 0; i < N; i++)
            {
                if (res[i].quality > icp_out.goodness)
                {  // cpp:785-786
                    icp_out.goodness = res[i].quality;
                    memcpy(icp_out.found_pose_to_wrt_from.mean.R, res[i].R, sizeof(res[i].R));
                    memcpy(icp_out.found_pose_to_wrt_from.mean.t, res[i].t, sizeof(res[i].t));
                    memcpy(icp_out.found_pose_to_wrt_from.cov, res[i].cov, sizeof(res[i].cov));
                    icp_out.n_iterations = res[i].n_iterations;
                    icp_out.termination_reason = res[i].termination_reason;
                }
            }
            {
                MonteCarloRecord rec;
                rec.from_id = d->from_id, rec.to_id = d->to_id;
                rec.guesses = guesses;
                for (size_t i = 0; i < N; i++) rec.goodness.push_back(res[i].quality);
                rec.best_goodness = icp_out.goodness;
                b2::pose_to_ypr(icp_out.found_pose_to_wrt_from.mean, rec.best_pose);
                std::lock_guard<std::mutex> lk(state_mtx_);
                last_mc_ = std::move(rec);
            }
            // d->init_guess_to_wrt_from keeps the LAST perturbed guess in the
            // reference (cpp:777-781 writes through d): same here
            if (N)
                d->init_guess_to_wrt_from = TPose3D{guesses[6 * (N - 1)],     guesses[6 * (N - 1) + 1],
                                                    guesses[6 * (N - 1) + 2], guesses[6 * (N - 1) + 3],
                                                    guesses[6 * (N - 1) + 4], guesses[6 * (N - 1) + 5]};
        }
        const CPose3D rel_pose = icp_out.found_pose_to_wrt_from.getMeanVal();
        const double  icp_goodness = icp_out.goodness;
        // Accept the new edge? (cpp:794-816)
        const CPose3D init_guess = to_CPose3D(d->init_guess_to_wrt_from);
        CPose3D       diff;
        b2::pose_inverse_compose(init_guess, rel_pose, diff);  // rel_pose - init_guess
        const double pos_correction = pose_norm(diff);
        const double correction_percent = pos_correction / (pose_norm(init_guess) + 0.01);
        const double goodness_thres =
            (d->align_kind == AlignKind::LoopClosure ? params_.min_icp_goodness_lc : params_.min_icp_goodness);
        if (icp_goodness > goodness_thres &&
            (correction_percent < 0.2 || d->align_kind == AlignKind::LoopClosure) && slam_backend_)
        {
            FactorRelativePose3 fPose3;  // cpp:818-830
            fPose3.from_kf = d->from_id, fPose3.to_kf = d->to_id;
            fPose3.rel_pose = asTPose(rel_pose);
            auto factor_out = slam_backend_->addFactor(fPose3).get();
            if (!factor_out.success || !factor_out.new_factor_id || factor_out.new_factor_id == INVALID_FID)
                throw std::runtime_error("addFactor failed");
            std::lock_guard<std::mutex> lck(local_pose_graph_mtx);  // cpp:832-837
            state_.local_pose_graph.graph.insertEdgeAtEnd(d->from_id, d->to_id, rel_pose);
        }
    }
    catch (const std::exception& e)
    {
        profiler_.registerUserMeasure(std::string("exception: ") + e.what(), 1);  // cpp:845-848
    }
}

// ---- key-frame cloud store ---------------------------------------------------
void LidarOdometry::kf_store_add(const DeviceCloud::Ptr& c)
{
    {
        std::lock_guard<std::mutex> lk(kf_store_mtx_);
        kf_store_.push_back(c);
    }
    kf_store_enforce_budget();
}

void LidarOdometry::kf_store_enforce_budget()
{
    if (!(params_.kf_store_budget_mb > 0)) return;
    const double                  budget = params_.kf_store_budget_mb * 1024.0 * 1024.0;
    std::lock_guard<std::mutex>   lk(kf_store_mtx_);
    std::vector<DeviceCloud::Ptr> alive;
    for (auto it = kf_store_.begin(); it != kf_store_.end();)
    {
        if (auto p = it->lock())
        {
            alive.push_back(std::move(p));
            ++it;
        }
        else
            it = kf_store_.erase(it);
    }
    double total = 0;
    for (const auto& c : alive) total += (double)c->device_bytes();
    // least recently used first (a cloud is "used" when it is created and whenever a registration pins it)
    std::sort(alive.begin(), alive.end(),
              [](const DeviceCloud::Ptr& a, const DeviceCloud::Ptr& b) { return a->last_use() < b->last_use(); });
    for (const auto& c : alive)
    {
        if (total <= budget) break;
        const double bytes = (double)c->device_bytes();
        if (bytes > 0 && c->spill()) total -= bytes;
    }
}

// cpp:851-895
void LidarOdometry::run_one_icp(const ICP_Input& in, ICP_Output& out)
{
    ProfilerEntry tle(profiler_, "run_one_icp");
    if (!in.from_pc || !in.to_pc) throw std::runtime_error("run_one_icp: null point cloud");  // ASSERT_ cpp:860-861

    TPose3D          current_solution = in.init_guess_to_wrt_from;
    b200icp_result_t icp_result;
    const double     guess[6] = {current_solution.x,   current_solution.y,     current_solution.z,
                             current_solution.yaw, current_solution.pitch, current_solution.roll};
    // the ICP object of `align_kind` -- its matchers, solvers and quality evaluators -- runs with the
    // mp2p_icp::Parameters the caller selected (cpp:869-871 passes in.icp_params next to the shared object)
    b200icp_t* icp = params_.icp.at(in.align_kind).icp;
    int        rc;
    {
        CloudPin pin_from(in.from_pc), pin_to(in.to_pc);  // resident (re-uploaded if spilled) for the call
        if (!pin_from.h || !pin_to.h) throw std::runtime_error(std::string("cloud reload: ") + b200icp_last_error());
        rc = b200icp_align_with(icp, pin_from.h, pin_to.h, guess, &in.icp_params, &icp_result);
    }
    check_rc(rc, "b200icp_align");
    kf_store_enforce_budget();  // a reload may have pushed the store over its budget
    count_icp(1);

    if (icp_result.quality > 0)
    {  // Keep as init value for next stage (cpp:873-877)
        current_solution = TPose3D{icp_result.pose[0], icp_result.pose[1], icp_result.pose[2],
                                   icp_result.pose[3], icp_result.pose[4], icp_result.pose[5]};
    }
    memcpy(out.found_pose_to_wrt_from.mean.R, icp_result.R, sizeof(icp_result.R));  // cpp:879
    memcpy(out.found_pose_to_wrt_from.mean.t, icp_result.t, sizeof(icp_result.t));
    memcpy(out.found_pose_to_wrt_from.cov, icp_result.cov, sizeof(icp_result.cov));
    out.goodness = icp_result.quality;  // cpp:880
    out.n_iterations = icp_result.n_iterations;
    out.termination_reason = icp_result.termination_reason;
}

}  // namespace mola
