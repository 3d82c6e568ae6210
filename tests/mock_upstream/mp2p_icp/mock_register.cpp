// mock of the upstream libraries' own class registration (what loading libmp2p_icp.so does)
#include <mp2p_icp/ICP.h>
#include <mp2p_icp/Matcher_Point2Plane.h>
#include <mrpt/core/initializer.h>
IMPLEMENTS_MRPT_OBJECT(ICP, mrpt::rtti::CObject, mp2p_icp)
IMPLEMENTS_MRPT_OBJECT(Matcher_Point2Plane, Matcher, mp2p_icp)
IMPLEMENTS_MRPT_OBJECT(Matcher_Points_DistanceThreshold, Matcher, mp2p_icp)
IMPLEMENTS_MRPT_OBJECT(Solver_GaussNewton, Solver, mp2p_icp)
IMPLEMENTS_MRPT_OBJECT(Solver_Horn, Solver, mp2p_icp)
IMPLEMENTS_MRPT_OBJECT(QualityEvaluator_PairedRatio, QualityEvaluator, mp2p_icp)
MRPT_INITIALIZER(register_mock_mp2p_icp)
{
    using namespace mp2p_icp;
    mrpt::rtti::registerClass(CLASS_ID(ICP));
    mrpt::rtti::registerClass(CLASS_ID(Matcher_Point2Plane));
    mrpt::rtti::registerClass(CLASS_ID(Matcher_Points_DistanceThreshold));
    mrpt::rtti::registerClass(CLASS_ID(Solver_GaussNewton));
    mrpt::rtti::registerClass(CLASS_ID(Solver_Horn));
    mrpt::rtti::registerClass(CLASS_ID(QualityEvaluator_PairedRatio));
}
