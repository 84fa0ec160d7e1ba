// dpe_fusion.cu — depth-map fusion on the device (RunFusion, DPE.cpp:1220-1370).
//
// The reference fuses on one CPU core: for every view in order, for every pixel in raster order, project the
// pixel's 3-D point into each source view, check reprojection error (< 2 px), relative depth difference
// (< 1 %) and normal angle (< 10 deg), accept the point if sum exp(-(e + 200 dd + 10 ang)) exceeds
// 0.3 (0.45 for WEAK pixels) per consistent view, and mark the source pixels it used so that later
// pixels / views skip them.  Here views are still processed in order (one launch per view: a later view
// sees every mark of an earlier one), the pixels of a view in parallel.  A mark records WHICH view set it
// (view index + 1), and a launch ignores the marks of its own view, so what a pixel sees never depends on
// how the launch was scheduled: the cloud is the same from run to run.  The one semantic difference: two
// pixels of the SAME view that land on the same source pixel both use it, where the reference's raster order
// gives it to the first — a slightly denser cloud, same points.  With several GPUs every rank fuses its own
// block of views against its own marks (marks do not cross ranks), which adds duplicates at the seams.
// Accepted points are appended to the cloud on the device, view after view, in raster order like the
// reference's (ordered compaction below); the cloud is copied out once at the end.
#include <cuda_runtime.h>

#include "dpe_fusion.cuh"

namespace dpe {

__device__ __forceinline__ void fuse_world_point(const FuseView& c, float x, float y, float depth, float X[3]) {  // DPE.cpp:1170-1194
  const float px = depth * (x - c.K[2]) / c.K[0], py = depth * (y - c.K[5]) / c.K[4], pz = depth;
  X[0] = c.R[0] * px + c.R[3] * py + c.R[6] * pz + c.C[0];
  X[1] = c.R[1] * px + c.R[4] * py + c.R[7] * pz + c.C[1];
  X[2] = c.R[2] * px + c.R[5] * py + c.R[8] * pz + c.C[2];
}
__device__ __forceinline__ void fuse_project(const FuseView& c, const float X[3], float* u, float* v, float* d) {  // DPE.cpp:1196-1206
  const float tx = c.R[0] * X[0] + c.R[1] * X[1] + c.R[2] * X[2] + c.t[0];
  const float ty = c.R[3] * X[0] + c.R[4] * X[1] + c.R[5] * X[2] + c.t[1];
  const float tz = c.R[6] * X[0] + c.R[7] * X[1] + c.R[8] * X[2] + c.t[2];
  *d = c.K[6] * tx + c.K[7] * ty + c.K[8] * tz;
  *u = (c.K[0] * tx + c.K[1] * ty + c.K[2] * tz) / *d;
  *v = (c.K[3] * tx + c.K[4] * ty + c.K[5] * tz) / *d;
}

__global__ void __launch_bounds__(256) k_fuse_view(const FuseView* __restrict__ views, const int i, const FuseSrcList srcs, const int W,
                                                   const int H, FusedPointDev* __restrict__ pts, uint8_t* __restrict__ accept) {
  const FuseView& ref = views[i];
  const uint16_t epoch = (uint16_t)(i + 1);
  const int total = W * H;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    accept[idx] = 0;
    if (ref.block != nullptr && ref.block[idx] < 128) continue;  // block mask (DPE.cpp:1296-1298): gates the reference pixel only
    if (ref.mask[idx] != 0) continue;  // marked by an earlier view (a view is never its own source)
    const float4 rp = ref.planes[idx];
    const float ref_depth = rp.w;
    if (ref_depth <= 0.0f) continue;
    const int r = idx / W, c = idx - r * W;
    const float rn0 = rp.x, rn1 = rp.y, rn2 = rp.z;
    float X[3];
    fuse_world_point(ref, (float)c, (float)r, ref_depth, X);
    int num_consistent = 0;
    float dynamic_consistency = 0.0f;
    int used[DPE_MAX_SRC];
    for (int j = 0; j < srcs.n; ++j) {
      used[j] = -1;
      const int s = srcs.id[j];
      if (s < 0) continue;
      const FuseView& sv = views[s];
      if (sv.planes == nullptr) continue;
      float u, v, pd;
      fuse_project(sv, X, &u, &v, &pd);
      const int sr = (int)(v + 0.5f), sc = (int)(u + 0.5f);
      if (!(sc >= 0 && sc < W && sr >= 0 && sr < H)) continue;
      const int sidx = sr * W + sc;
      const uint16_t sm = sv.mask[sidx];
      if (sm != 0 && sm != epoch) continue;  // marks of this launch do not count (see the header comment)
      const float4 sp4 = sv.planes[sidx];
      const float sd = sp4.w;
      if (sd <= 0.0f) continue;
      float Y[3], bu, bv;
      fuse_world_point(sv, (float)sc, (float)sr, sd, Y);
      fuse_project(ref, Y, &bu, &bv, &pd);
      const float reproj = sqrtf((c - bu) * (c - bu) + (r - bv) * (r - bv));
      const float rel = fabsf(pd - ref_depth) / ref_depth;
      float angle = acosf(rn0 * sp4.x + rn1 * sp4.y + rn2 * sp4.z);
      if (angle != angle) angle = 0.0f;
      if (reproj < 2.0f && rel < 0.01f && angle < 0.174533f) {
        used[j] = sidx;
        dynamic_consistency += expf(-(reproj + 200 * rel + angle * 10));
        num_consistent++;
      }
    }
    const float factor = (ref.state[idx] == DPE_WEAK ? 0.45f : 0.3f);
    if (num_consistent >= 1 && dynamic_consistency > factor * num_consistent) {
      const uint8_t* px = ref.bgr + 3 * (size_t)idx;
      float col[3] = {(float)px[0], (float)px[1], (float)px[2]};
      for (int j = 0; j < srcs.n; ++j) {
        if (used[j] < 0) continue;
        const FuseView& sv = views[srcs.id[j]];
        sv.mask[used[j]] = epoch;
        const uint8_t* sp = sv.bgr + 3 * (size_t)used[j];
        col[0] += sp[0]; col[1] += sp[1]; col[2] += sp[2];
      }
      FusedPointDev p;
      p.x = X[0]; p.y = X[1]; p.z = X[2];
      const uint32_t b = (uint32_t)(uint8_t)(col[0] / (num_consistent + 1)), g = (uint32_t)(uint8_t)(col[1] / (num_consistent + 1)),
                     rr = (uint32_t)(uint8_t)(col[2] / (num_consistent + 1));
      p.bgr = b | (g << 8) | (rr << 16);
      pts[idx] = p;
      accept[idx] = 1;
    }
  }
}

void launch_fuse_view(const FuseView* views, int i, const FuseSrcList& srcs, int W, int H, FusedPointDev* pts, uint8_t* accept,
                      int num_sms, cudaStream_t stream) {
  k_fuse_view<<<num_sms * 8, 256, 0, stream>>>(views, i, srcs, W, H, pts, accept);
}

// ---- appending a view's accepted points to the cloud on the device ----------------------------------------------
// Ordered compaction in three small passes (count per warp, one-CTA scan, scatter) with the running total kept on
// the device: points of a view land behind the points of the views before it, in raster order like the
// reference's cloud (DPE.cpp:1286-1364), and the host never has to read a count back between views.
constexpr int FC_THREADS = 256;
__device__ __forceinline__ void fuse_range(int total, int& begin, int& end) {
  const int n_warps = gridDim.x * (FC_THREADS / 32);
  int chunk = (total + n_warps - 1) / n_warps;
  chunk = (chunk + 31) & ~31;
  const int w = blockIdx.x * (FC_THREADS / 32) + (threadIdx.x >> 5);
  begin = min(w * chunk, total);
  end = min(begin + chunk, total);
}
__global__ void __launch_bounds__(FC_THREADS) k_fuse_count(const uint8_t* __restrict__ accept, int total, int* __restrict__ warp_counts) {
  int begin, end;
  fuse_range(total, begin, end);
  const int lane = threadIdx.x & 31;
  int n = 0;
  for (int base = begin; base < end; base += 32) {
    const int i = base + lane;
    n += __popc(__ballot_sync(0xffffffffu, i < end && accept[i] != 0));
  }
  if (lane == 0) warp_counts[blockIdx.x * (FC_THREADS / 32) + (threadIdx.x >> 5)] = n;
}
// exclusive scan of the warp counts, offset by the running total; the running total moves on (clamped to capacity)
__global__ void __launch_bounds__(1024) k_fuse_scan(int* __restrict__ warp_counts, int n, unsigned long long* __restrict__ running,
                                                    unsigned long long capacity) {
  __shared__ unsigned long long s_sum[32];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int per = (n + 1023) / 1024;
  const int b = min(t * per, n), e = min(b + per, n);
  unsigned long long loc = 0;
  for (int i = b; i < e; ++i) loc += (unsigned long long)warp_counts[i];
  unsigned long long v = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned long long u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
  if (lane == 31) s_sum[wid] = v;
  const unsigned long long pre = v - loc;
  __syncthreads();
  if (wid == 0) {
    const unsigned long long own = s_sum[lane];
    unsigned long long w = own;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += u; }
    s_sum[lane] = w - own;
    if (lane == 31) {  // w = this view's point count
      const unsigned long long base = running[0];
      running[1] = base;                                              // where this view's points start
      running[0] = base + w <= capacity ? base + w : base;            // a view that does not fit is dropped (running[2] counts)
      if (base + w > capacity) running[2] += 1;
    }
  }
  __syncthreads();
  unsigned long long run = pre + s_sum[wid];
  for (int i = b; i < e; ++i) {
    const int c = warp_counts[i];
    warp_counts[i] = (int)run;  // offset inside the view (< P)
    run += (unsigned long long)c;
  }
}
__global__ void __launch_bounds__(FC_THREADS) k_fuse_scatter(const FusedPointDev* __restrict__ pts, const uint8_t* __restrict__ accept, int total,
                                                             const int* __restrict__ warp_offsets, const unsigned long long* __restrict__ running,
                                                             unsigned long long capacity, FusedPointDev* __restrict__ cloud) {
  int begin, end;
  fuse_range(total, begin, end);
  const unsigned long long base = running[1];
  if (running[0] == base) return;  // nothing accepted, or the view did not fit
  const int lane = threadIdx.x & 31;
  int off = warp_offsets[blockIdx.x * (FC_THREADS / 32) + (threadIdx.x >> 5)];
  const unsigned lt = (1u << lane) - 1u;
  for (int b = begin; b < end; b += 32) {
    const int i = b + lane;
    const bool a = i < end && accept[i] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, a);
    if (a) cloud[base + (unsigned long long)(off + __popc(m & lt))] = pts[i];
    off += __popc(m);
  }
}
int fuse_append_blocks(int num_sms) { return num_sms * 4; }
void launch_fuse_append(const FusedPointDev* pts, const uint8_t* accept, int n, int* warp_counts, unsigned long long* running,
                        unsigned long long capacity, FusedPointDev* cloud, int num_sms, cudaStream_t stream) {
  const int blocks = fuse_append_blocks(num_sms);
  k_fuse_count<<<blocks, FC_THREADS, 0, stream>>>(accept, n, warp_counts);
  k_fuse_scan<<<1, 1024, 0, stream>>>(warp_counts, blocks * (FC_THREADS / 32), running, capacity);
  k_fuse_scatter<<<blocks, FC_THREADS, 0, stream>>>(pts, accept, n, warp_counts, running, capacity, cloud);
}

}  // namespace dpe
