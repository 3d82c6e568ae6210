// Registers the B200 classes in MRPT's class registry when the library is loaded, next to where a
// MOLA module registers itself (/root/reference/src/LidarOdometry.cpp:44-53), so that
// `icp_class: 'mola::ICP_B200'` (cpp:62-68), `class: mola::Matcher_B200` (cpp:83-84) and
// `class_name: mola::FilterEdgesPlanes_B200` in the filter pipeline (cpp:139-140) resolve.
#include <mola_b200/FilterEdgesPlanes_B200.h>
#include <mola_b200/ICP_B200.h>
#include <mola_b200/Matcher_B200.h>
#include <mrpt/core/initializer.h>

MRPT_INITIALIZER(do_register_mola_b200)
{
    mrpt::rtti::registerClass(CLASS_ID(mola::ICP_B200));
    mrpt::rtti::registerClass(CLASS_ID(mola::Matcher_B200));
    mrpt::rtti::registerClass(CLASS_ID(mola::FilterEdgesPlanes_B200));
}
