// dpe_types.h — plain structs shared by the CUDA kernels, the C-ABI layer and the
// host-side simulator used by the CPU tests.  Semantics follow the reference's
// Camera / PatchMatchParams / DataPassHelper (main.h:50-59, 78-106; DPE.h:52-86) but
// not their layout: everything that is constant per (reference view, source view,
// scale) is folded on the host in double precision once per scene.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/dpe_b200.h"
#include "dpe_rng.h"

namespace dpe {

// Per (reference view, source view, scale) constants.
//   Homography of plane (n,d) [n.X + d = 0 in ref-camera coords]  (DPE.cu:453-513):
//     H = Ks (Rrel - trel n^T / d) Kr^-1 = A - b (x) m,   m = Kr^-T n / d
//   Forward projection of ref pixel p at depth z (DPE.cu:881-913):  z * A p~ + b
//   Backward projection of src pixel q at depth z:                  z * Ai q~ + bi
struct SrcConst {
  float A[9];
  float b[3];
  float Ai[9];
  float bi[3];
  float baseline;        // |C_ref - C_src|  (DPE.cu:2640-2645)
  // fp32 camera data for the reference-order arithmetic (StageArgs::exact): relative pose as the reference's
  // ComputeHomography forms it (DPE.cu:455-481), and the source camera itself (geometric consistency)
  float Rrel[9], trel[3];
  float sR[9], st[3], sc3[3];
  float sK[9];           // source intrinsics at this scale, all nine entries as the reference multiplies by them
  float width, height;   // source image size at this scale
  int src_view;          // index of the source view in the scene = its layer in the scale's layered texture
  unsigned long long tex;  // host simulator: HostImage* of the source image (unused on the GPU)
  const float* depth;      // source depth map in the committed atlas (nullptr if none)
};

struct RefConst {
  int W, H;
  float fx, cx, fy, cy;
  float R[9];
  float t[3];
  float c[3];   // camera centre as ReadCamera stores it (DPE.cpp:362-367)
  float K9[9];  // intrinsics at this scale, all nine entries
  float depth_min, depth_max;  // PatchMatch range: cam file min*0.6, max*1.2 (DPE.cpp:788-789)
  int n_src;
  int view;
  SrcConst src[DPE_MAX_SRC];
};

// kernel variants (StageArgs::variants, test hook dpe_debug_set_variants)
enum {
  DPE_VARIANT_LIGHT_FULL_IMAGE = 4     // label boundary / nearest strong / anchor search / plane fit over the whole image instead of the WEAK lists
};

// Kernel argument block for one (view, stage).
#define DPE_RC_SLOTS 5  // folded-camera blocks in constant memory: one per stream of a stage + one for the test hooks

struct StageArgs {
  const RefConst* rc;    // set inside the kernels: &c_rc[slot]
  int slot;              // which constant-memory RefConst block this view-stage uses
  const float* ref_img;  // W*H float, row-major (the reference image at this scale)
  int W, H;
  // PatchMatch state (DPE.h:61-81)
  float4* planes;      // camera-space plane (n, d) while sweeping; (world n, depth) after extract
  float* costs;
  uint32_t* selected;  // selected_views bitmask
  uint4* view_w;       // 32 x 4-bit sampled view weights (view_weight_cuda, DPE.cu:1548)
  uint8_t* state;      // PixelState (weak_info_cuda)
  // weak / edge path
  float4* fit_planes;
  int* radius;
  const uint8_t* edge;      // edges_k
  const uint8_t* edge_low;  // edges of the coarsest scale (edge_low_res_cuda)
  const uint32_t* edge_low_bits;  // the same map, one bit per pixel, rows of low_words 32-bit words (nullptr: use edge_low)
  int low_words;
  int low_w, low_h;
  short2* edge_neigh;       // 8 per pixel
  float* complexity;        // complex_cuda
  const int32_t* label;     // labels_k
  short2* label_boundary;   // 8 per pixel
  uint8_t* weak_reliable;
  short2* nearest_strong;
  short2* neighbours;       // 9 per pixel (not compacted; neighbours_map is the identity)
  // WEAK pixels of this stage, compacted per colour after GenNeighbours: weak_list[c * list_stride + i],
  // i < weak_count[c], colour c = (x + y) & 1
  int* weak_list;
  int* weak_count;
  int list_stride;
  int* weak_scan;  // scratch of the ordered compaction: per-warp counts / offsets (compact_scan_entries ints)
  // state carried in from the previous stage (possibly at the previous scale)
  const float4* prev_planes;  // (world normal, depth)
  const uint8_t* prev_state;
  const uint32_t* prev_selected;
  int prev_W, prev_H;
  // outputs of the stage
  float4* out_planes;  // (world normal, depth), depth zeroed when out of range
  uint8_t* out_state;
  uint32_t* out_selected;
  float* atlas_out;  // this view's slot in the depth atlas being written
  // stage parameters
  int run_state, geom, use_apd, top_k, weak_peak_radius, rotate_time;
  float ransac_threshold, geom_factor;
  int iter, colour;
  unsigned char* debug_accept;  // test hook (dpe_debug_stop_after at a strong-sweep step): per pixel, which candidate the sweep took
  int exact;     // 1: homography, source coordinates, bilateral weights and geometric consistency in the reference's fp32 operation order
  int cost_raw;  // cost arithmetic: 1 = moments on raw intensities like the reference, 0 = centred (dpe_core.cuh)
  int ref_race;  // 1: edge-mode direction 4 samples its own colour like the reference, racy;
                 // 2: the same positions read from a copy of the maps taken before the launch (what a reference thread
                 //    sees when the pixels up-left of it are still in flight): deterministic
  const float4* snap_planes;  // ref_race == 2: planes / costs as they were before this half-sweep
  const float* snap_costs;
  int variants;  // DPE_VARIANT_* bits (dpe_debug_set_variants): earlier forms of a kernel, kept for A/B tests
  Xorwow* rng;  // per-pixel XORWOW state of this stage (dpe_rng.h), starts as curand_init(seed, y, x)
  unsigned long long* eval_units;  // optional counter (36-tap units)
  int tiles_x, tiles_y;
};

}  // namespace dpe
