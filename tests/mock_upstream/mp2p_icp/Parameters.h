#pragma once
#include <cstdint>
namespace mp2p_icp
{
struct WeightParameters
{
    bool   use_scale_outlier_detector = true;
    double scale_outlier_threshold    = 1.20;
    bool   use_robust_kernel          = false;
    double robust_kernel_param        = 0.05;  // radians
    double robust_kernel_scale        = 400.0;
};
struct Parameters
{
    uint32_t         maxIterations    = 40;
    double           minAbsStep_trans = 5e-4;
    double           minAbsStep_rot   = 1e-4;
    WeightParameters pairingsWeightParameters;
};
}  // namespace mp2p_icp
