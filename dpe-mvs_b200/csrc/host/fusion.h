// fusion.h — depth-map fusion into a point cloud:
//   RunFusion        DPE.cpp:1220-1370      ExportPointCloud  DPE.cpp:532-572
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include "io.h"
#include "../../../include/dpe_b200.h"

namespace dpe_host {

struct FusionInput {
  int width = 0, height = 0, n_views = 0;
  const std::vector<std::vector<float>>* depth = nullptr;    // per view H*W (0 = invalid)
  const std::vector<std::vector<float>>* normal = nullptr;   // per view H*W*3, world space
  const std::vector<std::vector<uint8_t>>* state = nullptr;  // per view H*W PixelState
  const std::vector<std::vector<uint8_t>>* bgr = nullptr;    // per view H*W*3
  const std::vector<CamFile>* cams = nullptr;
  std::vector<std::vector<int>> src;                          // per view: source view indices (-1 = not a problem)
};

struct FusedPoint { float x, y, z; uint8_t b, g, r; };

void fuse_views(const FusionInput& in, std::vector<FusedPoint>* cloud);
bool write_ply(const std::string& path, const std::vector<FusedPoint>& cloud);
bool run_fusion(const FusionInput& in, const std::string& ply_path);

}  // namespace dpe_host
