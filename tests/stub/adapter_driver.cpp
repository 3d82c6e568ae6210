// Drives the mp2p_icp adapters the way LidarOdometry.cpp:57-88 and 869-880 drive an ICP object:
// class factory by name -> initialize_solvers/matchers/quality_evaluators -> align(); against the mock
// upstream headers and the device test double.  Prints one line per check; exit code 0 = all passed.
#include <mola_b200/FilterEdgesPlanes_B200.h>
#include <mola_b200/ICP_B200.h>
#include <mola_b200/Matcher_B200.h>
#include <mp2p_icp/Matcher_Point2Plane.h>

#include <cmath>
#include <cstdio>
#include <thread>

using mrpt::containers::yaml;
extern "C" unsigned long fake_align_calls(void);
extern "C" unsigned long fake_upload_calls(void);
extern "C" unsigned long fake_last_call_max_iterations(void);
extern "C" const b200icp_edges_planes_params_t* fake_last_edges_planes_params(void);

static int g_fail = 0;
#define CHECK(cond)                                                       \
    do                                                                    \
    {                                                                     \
        if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); g_fail++; } \
        else std::printf("ok   %s\n", #cond);                             \
    } while (0)

static mp2p_icp::metric_map_t make_map(float x0, float y0, float z0, int n)
{
    auto pts = mrpt::maps::CSimplePointsMap::Create();
    for (int i = 0; i < n; i++) pts->insertPointFast(x0 + 0.01f * i, y0, z0);
    mp2p_icp::metric_map_t m;
    m.layers[mp2p_icp::metric_map_t::PT_LAYER_RAW] = pts;
    return m;
}

static yaml seq_of(const std::string& cls, const yaml& params)
{
    yaml e = yaml::Map();
    e["class"] = yaml(cls);
    e["params"] = params;
    yaml s = yaml::Sequence();
    s.push_back(e);
    return s;
}

int main()
{
    // --- load_icp_set_of_params (cpp:57-88) with icp_class: 'mola::ICP_B200'
    auto obj = mrpt::rtti::classFactory("mola::ICP_B200");
    auto icp = mrpt::ptr_cast<mp2p_icp::ICP>::from(obj);
    CHECK(icp != nullptr);
    CHECK(mrpt::rtti::classFactory("mola::NoSuchClass") == nullptr);

    yaml sp = yaml::Map(); sp["maxIterations"] = yaml(20);
    icp->initialize_solvers(seq_of("mp2p_icp::Solver_GaussNewton", sp));
    yaml mp = yaml::Map();
    mp["distanceThreshold"] = yaml(0.70), mp["planeEigenThreshold"] = yaml(0.07), mp["knn"] = yaml(6);
    mp["runFromIteration"] = yaml(0), mp["runUpToIteration"] = yaml(0);
    icp->initialize_matchers(seq_of("mp2p_icp::Matcher_Point2Plane", mp));
    yaml qp = yaml::Map(); qp["thresholdDistance"] = yaml(0.10);
    icp->initialize_quality_evaluators(seq_of("mp2p_icp::QualityEvaluator_PairedRatio", qp));

    mp2p_icp::Parameters prm;
    prm.maxIterations = 100, prm.minAbsStep_trans = 5e-5, prm.minAbsStep_rot = 1e-5;
    prm.pairingsWeightParameters.scale_outlier_threshold = 1.1;

    auto* b200 = dynamic_cast<mola::ICP_B200*>(icp.get());
    CHECK(b200 != nullptr);
    const b200icp_params_t q = b200->translate(prm);
    CHECK(q.max_iterations == 100 && q.min_abs_step_trans == 5e-5 && q.min_abs_step_rot == 1e-5);
    CHECK(q.solver_kind == B200ICP_SOLVER_GAUSS_NEWTON && q.solver_max_iterations == 20);
    CHECK(q.matcher_kind == B200ICP_MATCHER_POINT2PLANE && q.knn == 6 && q.distance_threshold == 0.70 &&
          q.plane_eigen_threshold == 0.07);
    CHECK(q.quality_threshold_distance == 0.10 && q.scale_outlier_threshold == 1.1);

    // --- run_one_icp (cpp:851-895): align + the Results fields the reference reads
    const auto from = make_map(1.f, 2.f, 0.f, 50), to = make_map(1.5f, 2.25f, 0.f, 40);
    mp2p_icp::Results res;
    icp->align(from, to, mrpt::math::TPose3D(0, 0, 0, 0.1, 0, 0), prm, res);
    CHECK(std::fabs(res.optimal_tf.mean.x() - 0.5) < 1e-6 && std::fabs(res.optimal_tf.mean.y() - 0.25) < 1e-6);
    CHECK(std::fabs(res.optimal_tf.mean.yaw() - 0.1) < 1e-12);      // the double echoes the guess's yaw
    CHECK(res.quality == 1.0 && res.nIterations == 3);
    CHECK(res.terminationReason == mp2p_icp::IterTermReason::Stalled);
    CHECK(res.optimal_tf.cov(0, 0) == 1e-4 && res.optimal_tf.cov(0, 1) == 0.0);

    // --- clouds are cached by map identity: a second align uploads nothing, a modified map is re-uploaded
    const unsigned long up0 = fake_upload_calls();
    icp->align(from, to, mrpt::math::TPose3D(), prm, res);
    CHECK(fake_upload_calls() == up0);
    auto pts = std::dynamic_pointer_cast<mrpt::maps::CPointsMap>(to.layers.at("raw"));
    pts->insertPointFast(9.f, 9.f, 9.f);
    icp->align(from, to, mrpt::math::TPose3D(), prm, res);
    CHECK(fake_upload_calls() == up0 + 1);
    CHECK(b200->cachedClouds() == 2);

    // --- other per-call Parameters on the same object (cpp:287-290 picks icp_params per scan)
    mp2p_icp::Parameters prm2 = prm;
    prm2.maxIterations = 50;
    icp->align(from, to, mrpt::math::TPose3D(), prm2, res);
    CHECK(res.quality == 1.0);
    CHECK(fake_last_call_max_iterations() == 50);  // travelled with the call ...
    CHECK(b200->contexts() == 1);                  // ... no second device object, no second upload
    CHECK(b200->cachedClouds() == 2);

    // --- concurrent align on one shared object (h:167-172, cpp:711-729)
    {
        const unsigned long a0 = fake_align_calls();
        std::vector<std::thread> th;
        for (int t = 0; t < 8; t++)
            th.emplace_back([&] { mp2p_icp::Results r; for (int i = 0; i < 20; i++) icp->align(from, to, mrpt::math::TPose3D(), prm, r); });
        for (auto& t : th) t.join();
        CHECK(fake_align_calls() == a0 + 160);
    }

    // --- unsupported combination: named, like an unknown icp_class at cpp:70-75
    {
        auto icp2 = mrpt::ptr_cast<mp2p_icp::ICP>::from(mrpt::rtti::classFactory("mola::ICP_B200"));
        icp2->initialize_solvers(seq_of("mp2p_icp::Solver_GaussNewton", sp));
        icp2->initialize_matchers(seq_of("mola::Matcher_B200", mp));  // a matcher the device loop does not know
        icp2->initialize_quality_evaluators(seq_of("mp2p_icp::QualityEvaluator_PairedRatio", qp));
        bool threw = false;
        try { mp2p_icp::Results r; icp2->align(from, to, mrpt::math::TPose3D(), prm, r); }
        catch (const std::exception& e) { threw = std::string(e.what()).find("mola::Matcher_B200") != std::string::npos; }
        CHECK(threw);
    }

    // --- Matcher_B200 inside a stock ICP's matcher list (cpp:83-84)
    {
        auto m = mrpt::ptr_cast<mp2p_icp::Matcher>::from(mrpt::rtti::classFactory("mola::Matcher_B200"));
        CHECK(m != nullptr);
        yaml mp2 = mp; mp2["runFromIteration"] = yaml(2);
        m->initialize(mp2);
        auto* mb = dynamic_cast<mola::Matcher_B200*>(m.get());
        CHECK(mb && mb->knn == 6 && mb->distanceThreshold == 0.70 && mb->runFromIteration == 2);
        mp2p_icp::Pairings pr;
        mp2p_icp::MatchState ms;
        mp2p_icp::MatchContext mc;
        mc.icpIteration = 0;
        CHECK(!m->match(from, to, mrpt::poses::CPose3D(), mc, ms, pr) && pr.empty());  // gated by runFromIteration
        mc.icpIteration = 2;
        CHECK(m->match(from, to, mrpt::poses::CPose3D(0.1, 0, 0), mc, ms, pr));
        // the double pairs every second local point with plane (centroid = the point moved by the pose, normal z)
        CHECK(pr.paired_pt2pl.size() == (to.point_layer("raw")->size() + 1) / 2);
        const auto& p0 = pr.paired_pt2pl[0];
        CHECK(std::fabs(p0.pl_global.centroid.x - (1.5 + 0.1)) < 1e-6 && p0.pt_local.x == 1.5f);
        CHECK(p0.pl_global.plane.coefs[2] == 1.0 && std::fabs(p0.pl_global.plane.coefs[3] + p0.pl_global.centroid.z) < 1e-12);
    }
    // --- FilterEdgesPlanes_B200 in the filter pipeline (cpp:139-140, 223-224)
    {
        auto f = mrpt::ptr_cast<mp2p_icp_filters::FilterBase>::from(mrpt::rtti::classFactory("mola::FilterEdgesPlanes_B200"));
        CHECK(f != nullptr);
        yaml fp = yaml::Map();
        fp["voxel_filter_resolution"] = yaml(0.5), fp["voxel_filter_decimation"] = yaml(2);
        fp["voxel_filter_min_e2_e0"] = yaml(100);
        f->initialize(fp);
        mp2p_icp::metric_map_t m = to;  // a map with the "raw" layer
        const std::size_t n = m.point_layer("raw")->size();
        mp2p_icp_filters::apply_filter_pipeline({f}, m);
        CHECK(m.point_layer("edges")->size() == n / 4 && m.point_layer("planes")->size() == n / 3 &&
              m.point_layer("full_decim")->size() == n / 2 && m.point_layer("raw")->size() == n);
        const auto* ep = fake_last_edges_planes_params();
        CHECK(ep->voxel_filter_resolution == 0.5f && ep->voxel_filter_decimation == 2 && ep->voxel_filter_min_e2_e0 == 100.f);
        CHECK(ep->full_pointcloud_decimation == 10 && ep->voxel_filter_max_e2_e0 == 30.f);  // the shipped defaults
        mp2p_icp::metric_map_t empty_map;
        bool threw = false;
        try { f->filter(empty_map); } catch (const std::exception&) { threw = true; }
        CHECK(threw);  // no input layer
    }
    std::printf("%s (%d failed)\n", g_fail ? "FAILED" : "ALL OK", g_fail);
    return g_fail ? 1 : 0;
}
