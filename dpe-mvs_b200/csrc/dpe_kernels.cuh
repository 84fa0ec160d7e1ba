// dpe_kernels.cuh — launch interface between the C-ABI layer (dpe_capi.cu) and the
// sm_100a kernels (dpe_kernels.cu).
#pragma once
#include "dpe_types.h"

namespace dpe {

struct KernelParams {
  StageArgs a;
  RefConst rc;
};

struct LaunchCfg {
  int num_sms;
  long long* launch_counter;  // host-side count of kernel launches
};

// selects the scale whose layered image texture the following kernels sample (device-wide constant:
// all kernels in flight on a device must be of the same scale — a stage is)
void launch_set_scale_tex(unsigned long long tex, cudaStream_t stream);
// copies a view-stage's folded cameras into constant-memory block `slot`, ordered on `stream` in front of the
// kernels that use it (StageArgs::slot); the host copy may be reused as soon as the call returns
void launch_set_ref_const(int slot, const RefConst* rc, cudaStream_t stream);

// stage kernels (one view, one stage); all asynchronous on `stream`
void launch_load(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t stream);
void launch_init(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t stream);
void launch_strong(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t stream);  // one colour, one iter
void launch_extract(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t stream);
void launch_median(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t stream);  // one colour
void launch_classify_refine(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t stream);
void launch_finish(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t stream);
// weak / edge path (DPE.cu kernels 2-5, 8, 9)
void launch_edge_info(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t stream);
void launch_nearest_strong(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t stream);
void launch_gen_neighbours(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t stream);
// WEAK pixels of each colour into raster-ordered lists; before the anchor search (its input) and again after it
// (it demotes unreliable pixels).  compact_scan_entries = ints StageArgs::weak_scan needs.
void launch_compact_weak(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t stream);
int compact_scan_entries(int num_sms);
void launch_fit_plane(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t stream);
void launch_weak(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t stream);  // one colour, one iter

// .npy payloads of a view from its carried maps (any output may be nullptr)
void launch_export(const float4* planes, const uint8_t* state, float* depth, float* normal3, int8_t* weak, int n,
                   const LaunchCfg& cfg, cudaStream_t stream);

// viz=True images of a view (depth ramp, normals, pixel classes) as interleaved BGR
void launch_viz(const float4* planes, const uint8_t* state, float depth_min, float depth_max, uint8_t* bgr_depth, uint8_t* bgr_normal,
                uint8_t* bgr_weak, int n, const LaunchCfg& cfg, cudaStream_t stream);

// scene preparation
void launch_u8_to_f32(const uint8_t* src, float* dst, int n, const LaunchCfg& cfg, cudaStream_t stream);
// cv::resize(INTER_LINEAR) of a float image (DPE.cpp:808)
void launch_relative_pose(const float* in, float* out, int n, cudaStream_t stream);
void launch_resize_linear(const float* src, int sw, int sh, float* dst, int dw, int dh,
                          const LaunchCfg& cfg, cudaStream_t stream);

// gate-1 hooks
void launch_cost_eval(const KernelParams& P, int n_pix, const int* xy, const float4* planes, int mode,
                      unsigned long long point_tex, float* out, const LaunchCfg& cfg, cudaStream_t stream);
void launch_geom_eval(const KernelParams& P, int n_pix, const int* xy, const float4* planes, float* out,
                      const LaunchCfg& cfg, cudaStream_t stream);

// NCC throughput study (variants of the tap loop; see dpe_kernels.cu)
void launch_ncc_bench(const KernelParams& P, const float4* world_planes, int n_cand, int variant, float* out,
                      const LaunchCfg& cfg, cudaStream_t stream);

// micro-benchmarks
void launch_probe_tex(unsigned long long tex, int w, int h, int iters, float* sink, int blocks, int threads,
                      const LaunchCfg& cfg, cudaStream_t stream);
void launch_probe_tex_pattern(unsigned long long tex, int w, int h, int iters, int layout, const float m[4], float* sink,
                              int blocks, int threads, const LaunchCfg& cfg, cudaStream_t stream);
void launch_probe_fma(int iters, float* sink, int blocks, int threads, const LaunchCfg& cfg,
                      cudaStream_t stream);
void launch_probe_weights(unsigned long long tex, int n, float* out, const LaunchCfg& cfg, cudaStream_t stream);

}  // namespace dpe
