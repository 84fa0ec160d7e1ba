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
size_t fuse_select_temp_bytes(int n);
void launch_fuse_select(void* temp, size_t temp_bytes, const FusedPointDev* pts, const uint8_t* accept, FusedPointDev* out, int* n_out,
                        int n, cudaStream_t stream);

}  // namespace dpe
