#pragma once
#include <mrpt/containers/yaml.h>
#include <mrpt/rtti/CObject.h>
#include <vector>
namespace mp2p_icp
{
class QualityEvaluator : public mrpt::rtti::CObject
{
   public:
    using Ptr = std::shared_ptr<QualityEvaluator>;
    virtual void initialize(const mrpt::containers::yaml&) {}
};
class QualityEvaluator_PairedRatio : public QualityEvaluator
{
    DEFINE_MRPT_OBJECT(QualityEvaluator_PairedRatio, mp2p_icp)
   public:
    void initialize(const mrpt::containers::yaml& p) override { thresholdDistance = p.getOrDefault<double>("thresholdDistance", thresholdDistance); }
    double thresholdDistance = 0.10;
};
struct QualityEvaluatorEntry { QualityEvaluator::Ptr obj; double relativeWeight = 1.0; };
using quality_eval_list_t = std::vector<QualityEvaluatorEntry>;
}  // namespace mp2p_icp
