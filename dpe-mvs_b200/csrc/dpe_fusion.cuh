// dpe_fusion.cuh — launch interface of the device fusion (dpe_fusion.cu).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/dpe_b200.h"

namespace dpe {

struct FuseView {
  const float4* planes;  // H*W (world normal, depth); depth 0 = invalid; nullptr = view has no maps (source-only image)
  const uint8_t* state;  // PixelState
  const uint8_t* bgr;    // H*W*3
  uint16_t* mask;        // H*W: 0 = free, e = fused into a point while view e - 1 was the reference view
  const uint8_t* block;  // H*W or nullptr: <dense>/blocks/mask_<id>.jpg; below 128 = not fused as a reference pixel (DPE.cpp:1296)
  float K[9], R[9], t[3], C[3];
};
struct FuseSrcList {
  int n;
  int id[DPE_MAX_SRC];   // view indices, -1 = skip
};
struct FusedPointDev {
  float x, y, z;
  uint32_t bgr;          // b | g << 8 | r << 16
};

void launch_fuse_view(const FuseView* views, int i, const FuseSrcList& srcs, int W, int H, FusedPointDev* pts, uint8_t* accept,
                      int num_sms, cudaStream_t stream);
// appends the accepted points of one view to `cloud` (device) behind running[0] points; running = {total so far,
// start of this view, views dropped for lack of capacity}; warp_counts: fuse_append_blocks(num_sms) * 8 ints
int fuse_append_blocks(int num_sms);
void launch_fuse_append(const FusedPointDev* pts, const uint8_t* accept, int n, int* warp_counts, unsigned long long* running,
                        unsigned long long capacity, FusedPointDev* cloud, int num_sms, cudaStream_t stream);

}  // namespace dpe
