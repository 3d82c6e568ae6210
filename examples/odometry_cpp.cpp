// odometry_cpp.cpp -- the module driven from C++, the way mola-launcher drives
// the reference (include/mola-fe-lidar/LidarOdometry.h:29-43): initialize(Yaml)
// once, then onNewObservation() per scan; key-frames and factors arrive at the
// back-end.  Links lib/libmola_fe_lidar_b200.so (which calls the CUDA path
// through the C ABI of include/b200icp.h).  Build and run:
//
//   g++ -std=c++17 -O2 -I include -I mola-fe-lidar_b200/host examples/odometry_cpp.cpp \
//       -L mola-fe-lidar_b200/lib -lmola_fe_lidar_b200 -lb200icp \
//       -Wl,-rpath,$PWD/mola-fe-lidar_b200/lib -lpthread -o /tmp/odometry_cpp
//   /tmp/odometry_cpp mola-fe-lidar_b200            # argument: the package directory (YAML files)
//
// The scans here are a toy scene (a room seen from a sensor moving 1 m per
// scan); a real caller hands over the float buffers of its point cloud.
#include <cmath>
#include <cstdio>
#include <filesystem>
#include <random>
#include <string>

#include "LidarOdometry.h"

using namespace mola;

static CObservation::Ptr make_scan(double sensor_x, double stamp, std::mt19937& rng)
{
    auto o = std::make_shared<CObservation>();
    o->sensorLabel = "lidar";
    o->timestamp = stamp;
    std::uniform_real_distribution<float> u(-1.f, 1.f);
    for (int i = 0; i < 20000; i++)
    {  // floor + two walls of a 40 m x 12 m corridor, in the SENSOR frame
        float x = 20.f * u(rng), y = 6.f * u(rng), z = -1.7f;
        const int s = i % 3;
        if (s == 1) y = 6.f, z = 1.5f * u(rng);
        if (s == 2) y = -6.f, z = 1.5f * u(rng);
        if (s == 0 && (i % 7) == 0) x = 10.f, y = 3.f * u(rng), z = 1.f * u(rng);  // a box face: constrains x
        o->x.push_back(x - (float)sensor_x), o->y.push_back(y), o->z.push_back(z);
    }
    return o;
}

int main(int argc, char** argv)
{
    const std::string pkg = std::filesystem::absolute(argc > 1 ? argv[1] : "mola-fe-lidar_b200").string();
    try
    {
        yaml_lite::Options opt;
        opt.module_dirs["mola-fe-lidar"] = pkg;  // what $(mola-dir mola-fe-lidar) resolves to
        const Yaml cfg = yaml_lite::parse("raw_sensor_label: lidar\nparams:\n  $include{" + pkg +
                                              "/params/kitti-default.yaml}\n",
                                          opt);
        auto wm = std::make_shared<WorldModel>();
        auto backend = std::make_shared<SimpleBackEnd>(wm);
        LidarOdometry lo;
        lo.slam_backend_ = backend;
        lo.setWorldModel(wm);
        lo.initialize_common(cfg);
        lo.initialize(cfg);

        std::mt19937 rng(1);
        for (int i = 0; i < 8; i++)
        {
            auto obs = make_scan(1.0 * i, 0.1 * i, rng);
            lo.onNewObservation(obs);  // asynchronous, like the reference (worker pool of one thread)
            lo.waitIdle();
            const auto st = lo.stateCopy();
            std::printf("scan %d: processed %zu, registrations %zu, goodness %.3f, key-frames %zu\n", i,
                        st.n_processed, st.n_icp, st.last_icp_out.goodness, backend->kf_stamps.size());
        }
        std::printf("factors: %zu\n", backend->factors.size());
        for (const auto& f : backend->factors)
            std::printf("  KF %llu -> KF %llu: x %.3f y %.3f yaw %.4f\n", (unsigned long long)f.from_kf,
                        (unsigned long long)f.to_kf, f.rel_pose.x, f.rel_pose.y, f.rel_pose.yaw);
        return 0;
    }
    catch (const std::exception& e)
    {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
