#pragma once
#include <mp2p_icp/Pairings.h>
#include <mrpt/poses/CPose3DPDFGaussian.h>
namespace mp2p_icp
{
enum class IterTermReason : uint8_t { Undefined = 0, NoPairings, SolverError, MaxIterations, Stalled };
struct Results
{
    mrpt::poses::CPose3DPDFGaussian optimal_tf;
    size_t                          nIterations       = 0;
    IterTermReason                  terminationReason = IterTermReason::Undefined;
    double                          quality           = 0;
    Pairings                        finalPairings;
};
}  // namespace mp2p_icp
