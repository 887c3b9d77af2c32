// TEST INFRASTRUCTURE (oracle build shim) -- not product code.
// Minimal stand-in for <pcl/point_types.h>: only the 48-byte pcl::PointXYZINormal
// layout the reference's include/ikd-Tree/ikd_Tree.{h,cpp} needs.  PCL is not
// installed in this image; the layout follows PCL's documented struct
// (x,y,z,pad | normal_x,normal_y,normal_z,pad | intensity,curvature,pad,pad).
#pragma once
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <algorithm>
#include <vector>
#include <unistd.h>
namespace pcl {
struct alignas(16) PointXYZINormal {
    float x = 0.f, y = 0.f, z = 0.f, _pad0 = 1.f;
    float normal_x = 0.f, normal_y = 0.f, normal_z = 0.f, _pad1 = 0.f;
    float intensity = 0.f, curvature = 0.f, _pad2 = 0.f, _pad3 = 0.f;
};
static_assert(sizeof(PointXYZINormal) == 48, "PointXYZINormal must be 48 bytes");
}  // namespace pcl
