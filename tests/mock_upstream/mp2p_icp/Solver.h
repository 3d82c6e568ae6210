#pragma once
#include <mrpt/containers/yaml.h>
#include <mrpt/rtti/CObject.h>
#include <vector>
namespace mp2p_icp
{
class Solver : public mrpt::rtti::CObject
{
   public:
    using Ptr = std::shared_ptr<Solver>;
    virtual void initialize(const mrpt::containers::yaml&) {}
};
class Solver_GaussNewton : public Solver
{
    DEFINE_MRPT_OBJECT(Solver_GaussNewton, mp2p_icp)
   public:
    void initialize(const mrpt::containers::yaml& p) override { maxIterations = p.getOrDefault<uint32_t>("maxIterations", maxIterations); }
    uint32_t maxIterations = 5;
};
class Solver_Horn : public Solver
{
    DEFINE_MRPT_OBJECT(Solver_Horn, mp2p_icp)
};
using solver_list_t = std::vector<Solver::Ptr>;
}  // namespace mp2p_icp
