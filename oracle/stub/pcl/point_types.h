/* Stand-in for <pcl/point_types.h> so the reference's lesson_16.{h,cu} compile
 * where PCL is not installed.  The reference only needs fixed-width ints, libm and
 * the (no-op here) point-struct registration macro.  Test infrastructure only. */
#pragma once
#include <cstdint>
#include <cmath>
#define POINT_CLOUD_REGISTER_POINT_STRUCT(...)
