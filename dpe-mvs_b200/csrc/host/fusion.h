// fusion.h — PLY export of the fused point cloud: ExportPointCloud, DPE.cpp:532-572.
// The fusion itself runs on the device (csrc/dpe_fusion.cu, dpe_fuse_* of the C ABI).
#pragma once
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>

namespace dpe_host {

// binary little-endian PLY: x y z float + diffuse_blue/green/red uchar per vertex
bool write_ply(const std::string& path, const float* xyz, const uint8_t* bgr, size_t n);
bool write_ply_records(const std::string& path, const std::vector<const uint8_t*>& parts, const std::vector<size_t>& counts);

}  // namespace dpe_host
