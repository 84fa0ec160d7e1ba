// dpe_kernels.cu — hand-written sm_100a kernels of the PatchMatch path.
//
// Kernel shape:
//  * every image-space kernel is persistent: grid = #SMs x resident CTAs, each CTA loops
//    over 32-wide pixel tiles;
//  * per tile the reference-image window (tile + 5-pixel halo) is staged into shared
//    memory with cp.async, and each thread derives the 36 hypothesis-invariant
//    (weight, weight*ref) pairs of its pixel from it ONCE into a [tap][thread] shared
//    table — the reference recomputes them (exp + sqrt + ref fetch per tap) for every one
//    of the ~100 (hypothesis, view) evaluations a pixel makes per sweep;
//  * per (hypothesis, view) the homography is built from the folded cameras in constant memory (c_rc[slot]) —
//    in the reference's own operation order by default, 9 FMAs (A - b m^T) with DPE_COST_REFERENCE — the 36
//    source taps go through the texture unit's bilinear filter, and the three moment sums are FMAs;
//  * a red/black sweep writes one colour and reads the other, so it is race-free and deterministic; the one
//    exception is the reference's edge-mode direction 4, which samples the colour being written: read from a
//    pre-sweep copy (ref_race 2, deterministic), live in the reference's launch geometry (ref_race 1,
//    k_half_refgeom), or moved onto the other colour (ref_race 0).
#include <cuda_pipeline.h>
#include "dpe_core.cuh"
#include "dpe_consts.h"
#include "dpe_weak.cuh"
#include "dpe_kernels.cuh"

namespace dpe {

constexpr int NT = 128;      // threads per CTA
#ifndef DPE_CTAS_PER_SM
#define DPE_CTAS_PER_SM 4
#endif
constexpr int CTAS_PER_SM = DPE_CTAS_PER_SM;  // resident CTAs per SM of the image-tile kernels (register budget 65536 / (128 * n))
constexpr int TILE_W = 32;
constexpr int HALO = 5;
constexpr int SMW = TILE_W + 2 * HALO;  // 42

// The layered image texture of the scale a stage runs at.  It lives in the constant bank at a fixed
// address (set by launch_set_scale_tex before the stage's kernels are queued) so that the handle of every
// TEX instruction is warp-uniform by construction, also inside non-inlined cost functions; a handle that
// arrives through a pointer or a struct makes the compiler wrap each fetch in a serialising
// per-unique-handle loop.
__constant__ cudaTextureObject_t c_scale_tex;

// Folded cameras of the view-stages in flight: one slot per stream of the stage (dpe_capi.cu queues a
// stream-ordered copy into the slot in front of a view-stage's kernels).  Kernels and the non-inlined cost
// functions index c_rc[slot] directly, which the compiler turns into constant-bank loads with a uniform index —
// the 8 KB block used to ride along as a __grid_constant__ kernel parameter, and a device function that received
// a pointer into it could only read it with generic loads.
__constant__ RefConst c_rc[DPE_RC_SLOTS];

struct DevEnv {
  const float2* tbl;  // &table[threadIdx.x]; tap t lives at tbl[t * NT]
  const float* img;   // reference image (for the far-away anchor patches of the weak path)
  int W, H;
  int slot;           // which c_rc block this view-stage uses
  __device__ __forceinline__ const RefConst& rc() const { return c_rc[slot]; }
  __device__ __forceinline__ const SrcConst& src(int v) const { return c_rc[slot].src[v]; }
  __device__ __forceinline__ float ref(int x, int y) const {
    return __ldg(&img[(size_t)iclamp(y, 0, H - 1) * W + iclamp(x, 0, W - 1)]);
  }
  __device__ __forceinline__ float tex(const SrcConst& sc, float u, float v) const {
    return tex2DLayered<float>(c_scale_tex, u, v, sc.src_view);
  }
  __device__ __forceinline__ float2 pw(int t) const { return tbl[t * NT]; }
};

// reference-image window in shared memory; clamp addressing reproduces the reference's
// exact-texel tex2D fetches at image borders
template <int TILE_H>
struct RefTile {
  static constexpr int SMH = TILE_H + 2 * HALO;
  float* s;
  int x0, y0;
  __device__ __forceinline__ void stage(const float* __restrict__ img, int W, int H, int tx0, int ty0) {
    x0 = tx0; y0 = ty0;
    for (int i = threadIdx.x; i < SMW * SMH; i += blockDim.x) {
      const int gx = iclamp(tx0 - HALO + (i % SMW), 0, W - 1);
      const int gy = iclamp(ty0 - HALO + (i / SMW), 0, H - 1);
      __pipeline_memcpy_async(&s[i], &img[(size_t)gy * W + gx], sizeof(float));
    }
    __pipeline_commit();
  }
  __device__ __forceinline__ float operator()(int x, int y) const {
    return s[(y - y0 + HALO) * SMW + (x - x0 + HALO)];
  }
};

struct TblStore {
  float2* tbl;
  __device__ __forceinline__ void operator()(int t, float w, float wr) const { tbl[t * NT] = make_float2(w, wr); }
};

__device__ __forceinline__ void flush_evals(unsigned long long* counter, unsigned evals) {
  if (counter == nullptr) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) evals += __shfl_down_sync(0xffffffffu, evals, o);
  if ((threadIdx.x & 31) == 0 && evals) atomicAdd(counter, (unsigned long long)evals);
}

// ---- full-image kernels: tile = 32 x 4 pixels, one thread per pixel ---------------------
enum FullOp { OP_INIT = 0, OP_CLASSIFY = 1 };

template <int OP>
__global__ void __launch_bounds__(NT, CTAS_PER_SM) k_full(const __grid_constant__ StageArgs A) {
  __shared__ float2 s_tbl[36 * NT];
  __shared__ float s_tile[SMW * (4 + 2 * HALO)];
  StageArgs a = A;
  a.rc = &c_rc[a.slot];
  const int n_tiles = a.tiles_x * a.tiles_y;
  unsigned evals = 0;
  RefTile<4> tile;
  tile.s = s_tile;
  DevEnv env;
  env.tbl = s_tbl + threadIdx.x;
  env.img = a.ref_img; env.W = a.W; env.H = a.H; env.slot = a.slot;
  TblStore st;
  st.tbl = s_tbl + threadIdx.x;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int tx0 = (t % a.tiles_x) * TILE_W, ty0 = (t / a.tiles_x) * 4;
    __syncthreads();
    tile.stage(a.ref_img, a.W, a.H, tx0, ty0);
    __pipeline_wait_prior(0);
    __syncthreads();
    const int x = tx0 + (threadIdx.x & 31), y = ty0 + (threadIdx.x >> 5);
    if (x < a.W && y < a.H) {
      const PatchStats ps = build_patch(tile, x, y, st, a.cost_raw != 0, a.exact != 0);
      if (OP == OP_INIT) init_pixel(env, ps, a, x, y, evals);
      else classify_refine_pixel(env, ps, a, x, y, evals);
    }
  }
  flush_evals(a.eval_units, evals);
}

// ---- red/black half sweeps: tile = 32 x 8 pixels, one thread per pixel of one colour ----
enum HalfOp { OP_STRONG = 0, OP_STRONG_EDGE = 1 };

template <int OP>
__global__ void __launch_bounds__(NT, CTAS_PER_SM) k_half(const __grid_constant__ StageArgs A) {
  __shared__ float2 s_tbl[36 * NT];
  __shared__ float s_tile[SMW * (8 + 2 * HALO)];
  StageArgs a = A;
  a.rc = &c_rc[a.slot];
  const int n_tiles = a.tiles_x * a.tiles_y;
  unsigned evals = 0;
  RefTile<8> tile;
  tile.s = s_tile;
  DevEnv env;
  env.tbl = s_tbl + threadIdx.x;
  env.img = a.ref_img; env.W = a.W; env.H = a.H; env.slot = a.slot;
  TblStore st;
  st.tbl = s_tbl + threadIdx.x;
  float cost_arr[9 * DPE_MAX_IMAGES];
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int tx0 = (t % a.tiles_x) * TILE_W, ty0 = (t / a.tiles_x) * 8;
    __syncthreads();
    tile.stage(a.ref_img, a.W, a.H, tx0, ty0);
    __pipeline_wait_prior(0);
    __syncthreads();
    const int lx = threadIdx.x & 31;
    const int x = tx0 + lx, y = ty0 + 2 * (threadIdx.x >> 5) + ((lx + a.colour) & 1);
    if (x < a.W && y < a.H) {
      if (a.state[y * a.W + x] != DPE_WEAK) {
        const PatchStats ps = build_patch(tile, x, y, st, a.cost_raw != 0, a.exact != 0);
        strong_update_pixel<OP == OP_STRONG_EDGE>(env, ps, a, x, y, cost_arr, evals);
      }
    }
  }
  flush_evals(a.eval_units, evals);
}

// The edge-mode strong sweep in the REFERENCE's launch geometry, for dpe_set_reference_race(ctx, 1): one CTA of
// 16 warps per 32 x 32 pixel block, warp w on the row pair 2w / 2w+1, CTAs handed out in raster order, one per SM
// (128 registers x 512 threads) — BlackPixelUpdateStrong<<<(W/32, H/2/16), (32,16)>>> (DPE.cu:3129-3148, 3199-3201).
// What a direction-4 read of the launch's own colour sees depends on which threads have already written: warps whose
// pixels are cheap (views outside the source images) finish long before their block mates, earlier waves of blocks
// before later ones.  The same geometry reproduces those systematic parts of the reference's race; the persistent
// 8-row tiles of k_half do not.  The per-thread table keeps its stride of NT: four groups of 128 threads.
constexpr int REFGEOM_THREADS = 512;
constexpr int REFGEOM_SMEM = 4 * 36 * NT * (int)sizeof(float2) + SMW * (32 + 2 * HALO) * (int)sizeof(float);
template <int OP>
__global__ void __launch_bounds__(REFGEOM_THREADS, 1) k_half_refgeom(const __grid_constant__ StageArgs A) {
  extern __shared__ float2 s_dyn[];
  float2* s_tbl = s_dyn + (threadIdx.x / NT) * 36 * NT + (threadIdx.x % NT);
  StageArgs a = A;
  a.rc = &c_rc[a.slot];
  unsigned evals = 0;
  RefTile<32> tile;
  tile.s = reinterpret_cast<float*>(s_dyn + 4 * 36 * NT);
  DevEnv env;
  env.tbl = s_tbl;
  env.img = a.ref_img; env.W = a.W; env.H = a.H; env.slot = a.slot;
  TblStore st;
  st.tbl = s_tbl;
  float cost_arr[9 * DPE_MAX_IMAGES];
  const int tx0 = (blockIdx.x % a.tiles_x) * TILE_W, ty0 = (blockIdx.x / a.tiles_x) * 32;
  tile.stage(a.ref_img, a.W, a.H, tx0, ty0);
  __pipeline_wait_prior(0);
  __syncthreads();
  const int lx = threadIdx.x & 31;
  const int x = tx0 + lx, y = ty0 + 2 * (threadIdx.x >> 5) + ((lx + a.colour) & 1);
  if (x < a.W && y < a.H) {
    if (a.state[y * a.W + x] != DPE_WEAK) {
      const PatchStats ps = build_patch(tile, x, y, st, a.cost_raw != 0, a.exact != 0);
      strong_update_pixel<OP == OP_STRONG_EDGE>(env, ps, a, x, y, cost_arr, evals);
    }
  }
  flush_evals(a.eval_units, evals);
}

// ---- WEAK pixels: compacted lists ------------------------------------------------------------
// Only a few per cent of the pixels are WEAK, scattered over the image; sweeping them through
// the image-tile kernel leaves most lanes idle behind a handful of long-running ones (the
// reference does exactly that, DPE.cu:1864-1898).  The stage therefore compacts the WEAK pixels
// of each colour once (after GenNeighbours has demoted the unreliable ones) and the weak sweep
// runs over the dense list; so do the nearest-strong search, the label-boundary walk, the anchor search and the
// plane fit, which only WEAK pixels take part in (k_list below).
struct GlobalRef {
  const float* img;
  int W, H;
  __device__ __forceinline__ float operator()(int x, int y) const {
    return __ldg(&img[(size_t)iclamp(y, 0, H - 1) * W + iclamp(x, 0, W - 1)]);
  }
};

// Ordered compaction in three small passes (count per warp, scan, scatter): every warp owns a contiguous run
// of pixels, so each colour's list comes out in raster order — neighbouring list entries are neighbouring WEAK
// pixels, which keeps the warps of the list kernels on similar work and their reads close together — and is
// the same from run to run.
constexpr int CW_THREADS = 256;
constexpr int CW_BLOCKS_PER_SM = 4;
__device__ __forceinline__ void compact_range(const StageArgs& a, int& begin, int& end) {
  const int total = a.W * a.H;
  const int n_warps = gridDim.x * (CW_THREADS / 32);
  int chunk = (total + n_warps - 1) / n_warps;
  chunk = (chunk + 31) & ~31;
  const int w = blockIdx.x * (CW_THREADS / 32) + (threadIdx.x >> 5);
  begin = imin(w * chunk, total);
  end = imin(begin + chunk, total);
}
__global__ void __launch_bounds__(CW_THREADS) k_weak_count(const __grid_constant__ StageArgs A) {
  const StageArgs& a = A;
  int begin, end;
  compact_range(a, begin, end);
  const int lane = threadIdx.x & 31;
  int n0 = 0, n1 = 0;
  for (int base = begin; base < end; base += 32) {
    const int i = base + lane;
    const bool weak = i < end && a.state[i] == DPE_WEAK;
    const int colour = ((i % a.W) + (i / a.W)) & 1;
    n0 += __popc(__ballot_sync(0xffffffffu, weak && colour == 0));
    n1 += __popc(__ballot_sync(0xffffffffu, weak && colour == 1));
  }
  if (lane == 0) {
    const int w = blockIdx.x * (CW_THREADS / 32) + (threadIdx.x >> 5);
    a.weak_scan[2 * w] = n0; a.weak_scan[2 * w + 1] = n1;
  }
}
// exclusive scan of the per-warp counts (n entries per colour, interleaved), one CTA
__global__ void __launch_bounds__(1024) k_weak_scan(int* __restrict__ scan, const int n, int* __restrict__ count) {
  __shared__ int s_sum[2][32];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int per = (n + 1023) / 1024;
  const int b = imin(t * per, n), e = imin(b + per, n);
  int loc[2] = {0, 0};
  for (int i = b; i < e; ++i) { loc[0] += scan[2 * i]; loc[1] += scan[2 * i + 1]; }
  int pre[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    int v = loc[c];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
    if (lane == 31) s_sum[c][wid] = v;
    pre[c] = v - loc[c];
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int own = s_sum[c][lane];
      int v = own;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
      s_sum[c][lane] = v - own;
      if (lane == 31) count[c] = v;
    }
  }
  __syncthreads();
  int run[2] = {pre[0] + s_sum[0][wid], pre[1] + s_sum[1][wid]};
  for (int i = b; i < e; ++i) {
    const int c0 = scan[2 * i], c1 = scan[2 * i + 1];
    scan[2 * i] = run[0]; scan[2 * i + 1] = run[1];
    run[0] += c0; run[1] += c1;
  }
}
__global__ void __launch_bounds__(CW_THREADS) k_weak_scatter(const __grid_constant__ StageArgs A) {
  const StageArgs& a = A;
  int begin, end;
  compact_range(a, begin, end);
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (CW_THREADS / 32) + (threadIdx.x >> 5);
  int off[2] = {a.weak_scan[2 * w], a.weak_scan[2 * w + 1]};
  const unsigned lt = (1u << lane) - 1u;
  for (int base = begin; base < end; base += 32) {
    const int i = base + lane;
    const bool weak = i < end && a.state[i] == DPE_WEAK;
    const int colour = ((i % a.W) + (i / a.W)) & 1;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const unsigned m = __ballot_sync(0xffffffffu, weak && colour == c);
      if (weak && colour == c) a.weak_list[c * a.list_stride + off[c] + __popc(m & lt)] = i;
      off[c] += __popc(m);
    }
  }
}

// One WARP per WEAK pixel.  A WEAK pixel scores 8 anchor planes x N views plus ~7 hypotheses x
// sampled views, each over up to 108 taps; one thread per pixel leaves that as a single long
// dependent chain on a few thousand threads.  Here the (hypothesis, view) pairs of a phase are
// spread over the 32 lanes, the reference-side table (WeakTab) is built once per pixel in shared
// memory, and the serial parts (view sampling with its RNG draws, accept chains) run on values every
// lane holds.  Per-pair arithmetic and all summation orders are those of weak_update_pixel
// (dpe_weak.cuh), which stays the definition (CPU simulator) — tests compare the two.
struct WarpEnv {
  const float2* tbl36;  // strong-patch table of the centre pixel (final re-score), shared memory
  const float* img;
  int W, H;
  int slot;
  __device__ __forceinline__ const RefConst& rc() const { return c_rc[slot]; }
  __device__ __forceinline__ const SrcConst& src(int v) const { return c_rc[slot].src[v]; }
  __device__ __forceinline__ float ref(int x, int y) const {
    return __ldg(&img[(size_t)iclamp(y, 0, H - 1) * W + iclamp(x, 0, W - 1)]);
  }
  __device__ __forceinline__ float tex(const SrcConst& sc, float u, float v) const {
    return tex2DLayered<float>(c_scale_tex, u, v, sc.src_view);
  }
  __device__ __forceinline__ float2 pw(int t) const { return tbl36[t]; }
};

struct WeakWarpSmem {
  WeakTab T;
  float cost[8 * DPE_MAX_IMAGES];   // candidate costs (cost_array of DPE.cu:1690)
  float hcost[7 * DPE_MAX_IMAGES];  // per-view costs of: current plane, fit plane, 5 refinement hypotheses
  float4 hplane[7];
  float4 cand[8];
  float2 tbl36[36];
};

struct Tbl36Store {
  float2* t;
  __device__ __forceinline__ void operator()(int i, float w, float wr) const { t[i] = make_float2(w, wr); }
};

__device__ __forceinline__ float bcast(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// PHASE_SYNC(): the warps of a CTA (each on its own pixel) enter every phase together, so that an SM
// executes a few code regions at a time instead of one per warp — the weak sweep's code is far larger
// than the instruction cache close to the schedulers, and un-synchronised warps spent two thirds of
// their time waiting for instruction fetches (profiles/r01_ncu_weak_warp.txt).
#define PHASE_SYNC() __syncthreads()
__device__ void weak_update_warp(const StageArgs& a, const RefConst& rc, const int x, const int y, WeakWarpSmem& S, int& taps,
                                 const bool live) {
  const int lane = threadIdx.x & 31;
  const bool writer = live && lane == 0;  // the lane that publishes this pixel's results
  const int W = a.W, H = a.H, N = rc.n_src, center = y * W + x;
  const int iter = a.iter;
  WarpEnv env{S.tbl36, a.ref_img, W, H, a.slot};
  GlobalRef ref{a.ref_img, W, H};
  WeakTab& T = S.T;
  // ---- set-up: anchors, strong-patch table + statistics (lane 0), deformable table (lanes 0..8)
  int first0 = 0, inc0 = 2;
  if (lane < DPE_NEIGHBOUR_NUM) init_weak_tab_entry(a, center, T, lane, first0, inc0);
  first0 = __shfl_sync(0xffffffffu, first0, 0); inc0 = __shfl_sync(0xffffffffu, inc0, 0);
  // strong-patch table (build_patch, dpe_core.cuh): one tap per lane, summed by lane 0 in tap order
  PatchStats ps;
  ps.r0 = ref(x, y);
  ps.c0 = a.cost_raw ? 0.0f : ps.r0;
  ps.exact = a.exact;
  for (int t = lane; t < 36; t += 32) {
    const int ixx = t / 6, jy = t - ixx * 6;
    const int i = 2 * ixx - 5, j = 2 * jy - 5;
    const float r = ref(x + i, y + j);
    const float w = bilateral_weight(i, j, r, ps.r0);
    const float rp = r - ps.c0;
    S.tbl36[t] = make_float2(w, mul_rn(w, rp));
    S.hcost[t] = rp;
  }
  __syncwarp();
  {
    float sw = 0.f, swr = 0.f, swrr = 0.f;
    if (lane == 0) {
      for (int ixx = 0; ixx < 6; ++ixx) {
        float sw_c = 0.f, swr_c = 0.f, swrr_c = 0.f;
        for (int jy = 0; jy < 6; ++jy) {
          const float2 ww = S.tbl36[ixx * 6 + jy];
          sw_c += ww.x; swr_c += ww.y; swrr_c = fmaf(ww.y, S.hcost[ixx * 6 + jy], swrr_c);
        }
        sw += sw_c; swr += swr_c; swrr += swrr_c;
      }
    }
    sw = bcast(sw, 0); swr = bcast(swr, 0); swrr = bcast(swrr, 0);
    ps.inv_sw = fast_rcp(sw);
    ps.mean_r = ps.inv_sw * swr;
    ps.var_r = fmaf(ps.inv_sw, swrr, -mul_rn(ps.mean_r, ps.mean_r));
  }
  __syncwarp();
  if (lane < DPE_NEIGHBOUR_NUM) build_weak_patch(ref, ps.r0, ps.c0, T, lane, lane == 0 ? first0 : -5, lane == 0 ? inc0 : 5);
  for (int i = lane; i < 8 * DPE_MAX_IMAGES; i += 32) S.cost[i] = 0.f;
  __syncwarp();
  if (lane == 0) S.cost[0] = 2.0f;
  // candidates = planes of the STRONG anchors
  bool my_flag = false;
  int my_pos = 0;
  if (lane < 8) {
    const short2 np = T.anchor[lane + 1];
    if (!(np.x == -1 || np.y == -1)) {
      const int npc = np.x + np.y * W;
      if (a.state[npc] == DPE_STRONG) { my_flag = true; my_pos = npc; S.cand[lane] = a.planes[npc]; }
    }
  }
  const unsigned flags = __ballot_sync(0xffffffffu, my_flag) & 0xffu;
  __syncwarp();
  PHASE_SYNC();
  // ---- phase 1: 8 candidates x N views
  for (int p = lane; p < 8 * N; p += 32) {
    const int j = p / N, v = p - j * N;
    if ((flags >> j) & 1u) {
      const float4 cpl = S.cand[j];
      const float3 m = plane_to_m(rc, cpl);
      S.cost[j * DPE_MAX_IMAGES + v] = ncc_new(env, ps.c0, T, rc.src[v], v, m, x, y, W, H, taps, a.exact ? &rc : nullptr, cpl);
    }
  }
  __syncwarp();
  PHASE_SYNC();
  // ---- phase 2: priors + view sampling (serial, lane 0 holds the pixel's RNG state)
  ViewW vw;
  float weight_norm = 0.f;
  uint32_t sel_bits = 0u;
  Rng rng;
  if (lane == 0) {
    float priors[DPE_MAX_IMAGES];
    for (int v = 0; v < N; ++v) priors[v] = 0.f;
    for (int i = 0; i < 8; ++i) {
      const short2 np = T.anchor[i + 1];
      if (np.x == -1 || np.y == -1) continue;
      const uint32_t sv = T.asel[i + 1];
      for (int v = 0; v < N; ++v) priors[v] += ((sv >> v) & 1u) ? 0.9f : 0.1f;
    }
    rng.load(a.rng + center);
    sample_views(S.cost, priors, N, iter, rng, vw, weight_norm, sel_bits);
    if (live) a.view_w[center] = vw.pack();
  }
  vw.lo = __shfl_sync(0xffffffffu, vw.lo, 0); vw.hi = __shfl_sync(0xffffffffu, vw.hi, 0);
  weight_norm = bcast(weight_norm, 0);
  sel_bits = __shfl_sync(0xffffffffu, sel_bits, 0);
  // sampled views, ascending
  int nv = 0;
  int vlist_lo = 0;  // this lane's entry of the list (lane i holds the i-th sampled view)
  for (int v = 0; v < N; ++v)
    if (vw.get(v) > 0) { if (nv == lane) vlist_lo = v; nv++; }
  // ---- phase 3: final candidate costs (lane j), arg-min with "<=" (last minimum wins)
  float my_final = 0.f;
  if (lane < 8) {
    const bool fl = (flags >> lane) & 1u;
    float4 cand = make_float4(0.f, 0.f, 1.f, 1.f);
    if (fl) cand = S.cand[lane];
    float f = 0.f;
    for (int v = 0; v < N; ++v) {
      const int w = vw.get(v);
      if (w > 0) {
        float c = S.cost[lane * DPE_MAX_IMAGES + v];
        if (a.geom) c += a.geom_factor * (fl ? geom_cost(a, rc, rc.src[v], cand, x, y) : 3.0f);
        f += w * c;
      }
    }
    my_final = f / weight_norm;
  }
  int min_idx = 0;
  float min_final = bcast(my_final, 0);
  for (int j = 1; j < 8; ++j) {
    const float fj = bcast(my_final, j);
    if (fj <= min_final) { min_final = fj; min_idx = j; }
  }
  PHASE_SYNC();
  // ---- phase 4: current plane and fit plane over the sampled views
  float4 plane_now = a.planes[center];
  const float4 fit = a.fit_planes[center];
  const bool has_fit = !(fit.x == 0 && fit.y == 0 && fit.z == 0);
  if (lane == 0) { S.hplane[0] = plane_now; S.hplane[1] = fit; }
  __syncwarp();
  // lane i < nv knows the i-th sampled view; spread it to a small shared list through registers
  int vlist[DPE_MAX_SRC > 15 ? 15 : DPE_MAX_SRC];  // at most 15 draws => at most 15 sampled views
#pragma unroll
  for (int i = 0; i < 15; ++i) vlist[i] = __shfl_sync(0xffffffffu, vlist_lo, i);
  auto view_at = [&](const int i) {
    int v = vlist[0];
#pragma unroll
    for (int q = 1; q < 15; ++q) if (i == q) v = vlist[q];
    return v;
  };
  auto score = [&](const int h0, const int nh) {
    for (int p = lane; p < nh * nv; p += 32) {
      const int hi = p / nv;
      const int v = view_at(p - hi * nv);
      const float4 pl = S.hplane[h0 + hi];
      const float3 m = plane_to_m(rc, pl);
      float cv = ncc_new(env, ps.c0, T, rc.src[v], v, m, x, y, W, H, taps, a.exact ? &rc : nullptr, pl);
      if (a.geom) cv += a.geom_factor * geom_cost(a, rc, rc.src[v], pl, x, y);
      S.hcost[(h0 + hi) * DPE_MAX_IMAGES + v] = cv;
    }
    __syncwarp();
  };
  auto weighted = [&](const int h) {  // same order as weighted_cost_weak
    float c = 0.f;
    for (int v = 0; v < N; ++v) {
      const int w = vw.get(v);
      if (w > 0) c += w * S.hcost[h * DPE_MAX_IMAGES + v];
    }
    return c / weight_norm;
  };
  score(0, has_fit ? 2 : 1);
  float cost_now = weighted(0);
  const float cost_before = cost_now;
  float depth_now = depth_from_plane(rc, plane_now, x, y);
  if ((flags >> min_idx) & 1u) {
    const float4 cand = S.cand[min_idx];
    const float db = depth_from_plane(rc, cand, x, y);
    if (db >= rc.depth_min && db <= rc.depth_max && min_final < cost_now) {
      depth_now = db; plane_now = cand; cost_now = min_final;
      if (writer) a.selected[center] = sel_bits;
    }
  }
  PHASE_SYNC();
  // ---- phase 5: PlaneHypothesisRefinementWeak (DPE.cu:1120-1212)
  if (has_fit) {
    const float dmin = rc.depth_min, dmax = rc.depth_max;
    {
      const float c = weighted(1);
      const float db = depth_from_plane(rc, fit, x, y);
      if (db >= dmin && db <= dmax && c < cost_now) { depth_now = db; plane_now = fit; cost_now = c; }
    }
    __syncwarp();
    if (lane == 0) {
      const float depth_rand = rng.uniform() * (dmax - dmin) + dmin;
      const float4 n_rand = random_normal(rc, x, y, rng, depth_now);
      const float lo = (1 - 0.02f) * depth_now, hi = (1 + 0.02f) * depth_now;
      const float depth_pert = rng.uniform() * (hi - lo) + lo;
      const float4 n_pert = perturbed_normal(rc, x, y, plane_now, rng, (float)(0.02f * 3.14159265358979323846));
      for (int i = 0; i < 5; ++i) {
        const float d = (i == 0 || i == 2) ? depth_rand : (i == 4 ? depth_pert : depth_now);
        float4 n = (i == 1 || i == 2) ? n_rand : (i == 3 ? n_pert : plane_now);
        n.w = dist2origin(rc, x, y, d, n);
        S.hplane[2 + i] = n;
      }
    }
    __syncwarp();
    score(2, 5);
    for (int i = 0; i < 5; ++i) {
      const float4 n = S.hplane[2 + i];
      const float c = weighted(2 + i);
      const float db = depth_from_plane(rc, n, x, y);
      if (db >= dmin && db <= dmax && c < cost_now) { depth_now = db; plane_now = n; cost_now = c; }
    }
  }
  PHASE_SYNC();
  if (writer) rng.store(a.rng + center);
  float4 final_plane = a.planes[center];
  if (a.run_state == DPE_REFINE_INIT) {
    if (cost_now < cost_before - 0.1) { final_plane = plane_now; if (writer) a.planes[center] = plane_now; }  // double, DPE.cu:1835
  } else {
    final_plane = plane_now;
    if (writer) a.planes[center] = plane_now;
  }
  // ---- phase 6: costs[] re-scored with the plain 6x6 NCC (DPE.cu:1845-1861)
  __syncwarp();
  {
    const float3 m = plane_to_m(rc, final_plane);
    if (lane < nv) {
      const int v = vlist_lo;
      const float cv = ncc_old(env, ps, v, final_plane, m, x, y);
      S.hcost[v] = cv;
      if (cv < 2.0f) taps += 36;
    }
    __syncwarp();
    if (writer) a.costs[center] = weighted(0);
  }
  __syncwarp();
}

constexpr int NTW = 128;  // threads per CTA of the weak sweep: 4 warps = 4 pixels in phase lock-step
#ifndef DPE_WEAK_CTAS_PER_SM
#define DPE_WEAK_CTAS_PER_SM 6
#endif
constexpr int WEAK_CTAS_PER_SM = DPE_WEAK_CTAS_PER_SM;
__global__ void __launch_bounds__(NTW, WEAK_CTAS_PER_SM) k_weak_list(const __grid_constant__ StageArgs A) {
  __shared__ WeakWarpSmem s_w[NTW / 32];
  StageArgs a = A;
  a.rc = &c_rc[a.slot];
  const int count = a.weak_count[a.colour];
  const int* list = a.weak_list + a.colour * a.list_stride;
  int taps = 0;
  const int warps_per_cta = NTW / 32;
  // consecutive list entries go to different CTAs first, so a short list still spreads over all SMs
  // CTA-uniform loop (the phases synchronise the CTA): group g = one list entry per
  // warp.  In the ragged last group a warp without an entry shadows the last one with its stores off.
  const int wid = threadIdx.x >> 5;
  const int groups = (count + warps_per_cta - 1) / warps_per_cta;
  for (int g = blockIdx.x; g < groups; g += gridDim.x) {
    const int i = g * warps_per_cta + wid;
    const bool live = i < count;
    const int center = list[live ? i : count - 1];
    weak_update_warp(a, c_rc[a.slot], center % a.W, center / a.W, s_w[wid], taps, live);
  }
  flush_evals(a.eval_units, (unsigned)((taps + 18) / 36));
}
#undef PHASE_SYNC

// ---- light per-pixel kernels ------------------------------------------------------------
enum LightOp { L_EXTRACT = 0, L_MEDIAN = 1, L_FINISH = 2, L_EDGE_INFO = 3, L_NEAREST = 4, L_NEIGH = 5, L_FIT = 6, L_LOAD = 7,
               L_LABEL_BOUNDARY = 8, L_FIT_COPY = 9 };

template <int OP>
__global__ void __launch_bounds__(256) k_light(const __grid_constant__ StageArgs A) {
  StageArgs a = A;
  a.rc = &c_rc[a.slot];
  const int total = (OP == L_MEDIAN) ? a.W * ((a.H + 1) / 2) : a.W * a.H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int x, y;
    if (OP == L_MEDIAN) {
      x = i % a.W;
      y = 2 * (i / a.W) + ((x + a.colour) & 1);
      if (y >= a.H) continue;
      if (a.state[y * a.W + x] == DPE_WEAK) continue;  // DPE.cu:2081
    } else {
      x = i % a.W;
      y = i / a.W;
    }
    if (OP == L_EXTRACT) extract_pixel(a, x, y);
    else if (OP == L_MEDIAN) median_pixel(a, x, y);
    else if (OP == L_FINISH) finish_pixel(a, x, y);
    else if (OP == L_EDGE_INFO) edge_info_pixel(a, x, y);
    else if (OP == L_NEAREST) nearest_strong_pixel(a, x, y);
    else if (OP == L_NEIGH) gen_neighbours_pixel(a, x, y);
    else if (OP == L_FIT) fit_plane_pixel(a, x, y);
    else if (OP == L_LOAD) load_pixel(a, x, y);
    else if (OP == L_LABEL_BOUNDARY) label_boundary_pixel(a, x, y);
    else if (OP == L_FIT_COPY) { if (a.state[i] != DPE_WEAK) a.fit_planes[i] = a.planes[i]; }  // DPE.cu:2945
  }
}

// The same per-pixel functions over the compacted WEAK lists (both colours): one thread per WEAK pixel, warps
// full of WEAK pixels instead of one here and there in an image-sized launch.  A pixel another thread demoted
// after the list was built (NeigbourUpdate) is skipped by the functions' own state test.
template <int OP, int THREADS>
__global__ void __launch_bounds__(THREADS) k_list(const __grid_constant__ StageArgs A) {
  StageArgs a = A;
  a.rc = &c_rc[a.slot];
  const int n0 = a.weak_count[0], total = n0 + a.weak_count[1];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int center = i < n0 ? a.weak_list[i] : a.weak_list[a.list_stride + (i - n0)];
    const int x = center % a.W, y = center / a.W;
    if (OP == L_NEAREST) nearest_strong_pixel(a, x, y);
    else if (OP == L_LABEL_BOUNDARY) label_boundary_pixel(a, x, y);
    else if (OP == L_NEIGH) gen_neighbours_pixel(a, x, y);
    else if (OP == L_FIT) fit_plane_pixel(a, x, y);
  }
}

void launch_set_ref_const(int slot, const RefConst* rc, cudaStream_t stream) {
  // n_src <= 31 sources are in use; the tail of src[] need not travel
  const size_t bytes = offsetof(RefConst, src) + (size_t)rc->n_src * sizeof(SrcConst);
  cudaMemcpyToSymbolAsync(c_rc, rc, bytes, (size_t)slot * sizeof(RefConst), cudaMemcpyHostToDevice, stream);
}

void launch_set_scale_tex(unsigned long long tex, cudaStream_t stream) {
  cudaMemcpyToSymbolAsync(c_scale_tex, &tex, sizeof(tex), 0, cudaMemcpyHostToDevice, stream);
}

// ---- launchers -----------------------------------------------------------------------------
static inline int persistent_grid(int n_tiles, int num_sms, int per_sm) {
  const int g = num_sms * per_sm;
  return n_tiles < g ? (n_tiles > 0 ? n_tiles : 1) : g;
}
static inline void count(const LaunchCfg& cfg) {
  if (cfg.launch_counter) ++*cfg.launch_counter;
}

void launch_init(const KernelParams& P0, const LaunchCfg& cfg, cudaStream_t stream) {
  KernelParams P = P0;
  P.a.tiles_x = (P.a.W + TILE_W - 1) / TILE_W;
  P.a.tiles_y = (P.a.H + 3) / 4;
  k_full<OP_INIT><<<persistent_grid(P.a.tiles_x * P.a.tiles_y, cfg.num_sms, CTAS_PER_SM), NT, 0, stream>>>(P.a);
  count(cfg);
}
void launch_classify_refine(const KernelParams& P0, const LaunchCfg& cfg, cudaStream_t stream) {
  KernelParams P = P0;
  P.a.tiles_x = (P.a.W + TILE_W - 1) / TILE_W;
  P.a.tiles_y = (P.a.H + 3) / 4;
  k_full<OP_CLASSIFY><<<persistent_grid(P.a.tiles_x * P.a.tiles_y, cfg.num_sms, CTAS_PER_SM), NT, 0, stream>>>(P.a);
  count(cfg);
}
void launch_strong(const KernelParams& P0, const LaunchCfg& cfg, cudaStream_t stream) {
  KernelParams P = P0;
  P.a.tiles_x = (P.a.W + TILE_W - 1) / TILE_W;
  P.a.tiles_y = (P.a.H + 7) / 8;
  const int g = persistent_grid(P.a.tiles_x * P.a.tiles_y, cfg.num_sms, CTAS_PER_SM);
  if (P.a.use_apd && P.a.ref_race == 2 && P.a.snap_planes) {
    const size_t n = (size_t)P.a.W * P.a.H;
    cudaMemcpyAsync((void*)P.a.snap_planes, P.a.planes, n * sizeof(float4), cudaMemcpyDeviceToDevice, stream);
    cudaMemcpyAsync((void*)P.a.snap_costs, P.a.costs, n * sizeof(float), cudaMemcpyDeviceToDevice, stream);
  }
  if (P.a.use_apd && P.a.ref_race == 1) {
    cudaFuncSetAttribute(k_half_refgeom<OP_STRONG_EDGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, REFGEOM_SMEM);  // per device
    P.a.tiles_y = (P.a.H + 31) / 32;
    k_half_refgeom<OP_STRONG_EDGE><<<P.a.tiles_x * P.a.tiles_y, REFGEOM_THREADS, REFGEOM_SMEM, stream>>>(P.a);
  } else if (P.a.use_apd) k_half<OP_STRONG_EDGE><<<g, NT, 0, stream>>>(P.a);
  else k_half<OP_STRONG><<<g, NT, 0, stream>>>(P.a);
  count(cfg);
}
void launch_weak(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t stream) {
  k_weak_list<<<cfg.num_sms * WEAK_CTAS_PER_SM, NTW, 0, stream>>>(P.a);
  count(cfg);
}
void launch_compact_weak(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t stream) {
  const int blocks = cfg.num_sms * CW_BLOCKS_PER_SM;
  k_weak_count<<<blocks, CW_THREADS, 0, stream>>>(P.a);
  k_weak_scan<<<1, 1024, 0, stream>>>(P.a.weak_scan, blocks * (CW_THREADS / 32), P.a.weak_count);
  k_weak_scatter<<<blocks, CW_THREADS, 0, stream>>>(P.a);
  count(cfg); count(cfg); count(cfg);
}
int compact_scan_entries(int num_sms) { return 2 * num_sms * CW_BLOCKS_PER_SM * (CW_THREADS / 32); }
template <int OP>
static void launch_light(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t stream) {
  k_light<OP><<<cfg.num_sms * 8, 256, 0, stream>>>(P.a);
  count(cfg);
}
template <int OP, int THREADS>
static void launch_list(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t stream, int blocks_per_sm) {
  k_list<OP, THREADS><<<cfg.num_sms * blocks_per_sm, THREADS, 0, stream>>>(P.a);
  count(cfg);
}
static inline bool full_image(const KernelParams& P) { return (P.a.variants & DPE_VARIANT_LIGHT_FULL_IMAGE) != 0; }
void launch_extract(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t s) { launch_light<L_EXTRACT>(P, cfg, s); }
void launch_median(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t s) { launch_light<L_MEDIAN>(P, cfg, s); }
void launch_finish(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t s) { launch_light<L_FINISH>(P, cfg, s); }
// GenEdgeInform: the per-pixel part over the image, the WEAK-only label-boundary walk over the list
void launch_edge_info(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t s) {
  launch_light<L_EDGE_INFO>(P, cfg, s);
  if (full_image(P)) launch_light<L_LABEL_BOUNDARY>(P, cfg, s);
  else launch_list<L_LABEL_BOUNDARY, 128>(P, cfg, s, 16);
}
void launch_nearest_strong(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t s) {
  if (full_image(P)) { launch_light<L_NEAREST>(P, cfg, s); return; }
  // every pixel that is not WEAK holds (-1, -1) (DPE.cu:2858)
  cudaMemsetAsync(P.a.nearest_strong, 0xFF, (size_t)P.a.W * P.a.H * sizeof(short2), s);
  launch_list<L_NEAREST, 128>(P, cfg, s, 16);
}
// The anchor search keeps ~3 KB of arrays per thread in local memory: how many of its warps are resident decides
// whether their stacks stay in L2 (126 MB) or thrash through DRAM (round-2 ncu at 32 warps/SM: 12 GB of DRAM reads
// in one 10 ms launch).
#ifndef DPE_NEIGH_BLOCKS_PER_SM
#define DPE_NEIGH_BLOCKS_PER_SM 16
#endif
void launch_gen_neighbours(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t s) {
  if (full_image(P)) launch_light<L_NEIGH>(P, cfg, s);
  else launch_list<L_NEIGH, 64>(P, cfg, s, DPE_NEIGH_BLOCKS_PER_SM);
}
void launch_fit_plane(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t s) {
  if (full_image(P)) { launch_light<L_FIT>(P, cfg, s); return; }
  launch_light<L_FIT_COPY>(P, cfg, s);
  launch_list<L_FIT, 64>(P, cfg, s, 16);
}
void launch_load(const KernelParams& P, const LaunchCfg& cfg, cudaStream_t s) { launch_light<L_LOAD>(P, cfg, s); }

// ---- result export ---------------------------------------------------------------------------
// The arrays of the .npy files, from a view's carried maps (world normal, depth) + state: depth zeroed where the
// pixel is UNKNOWN (ZeroDepthForUnknown, main.cpp:36-46), normals packed to 3 floats, the state remapped
// {UNKNOWN 0, WEAK 1, STRONG 2} (main.cpp:190-197).  Pure streaming: 17 B read, up to 17 B written per pixel.
__global__ void __launch_bounds__(256) k_export(const float4* __restrict__ planes, const uint8_t* __restrict__ state,
                                                float* __restrict__ depth, float* __restrict__ normal3,
                                                int8_t* __restrict__ weak, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = planes[i];
    const uint8_t st = state[i];
    if (depth) depth[i] = st == DPE_UNKNOWN ? 0.0f : p.w;
    if (normal3) { normal3[3 * i] = p.x; normal3[3 * i + 1] = p.y; normal3[3 * i + 2] = p.z; }
    if (weak) weak[i] = st == DPE_WEAK ? 1 : (st == DPE_STRONG ? 2 : 0);
  }
}
void launch_export(const float4* planes, const uint8_t* state, float* depth, float* normal3, int8_t* weak, int n,
                   const LaunchCfg& cfg, cudaStream_t stream) {
  k_export<<<cfg.num_sms * 8, 256, 0, stream>>>(planes, state, depth, normal3, weak, n);
  count(cfg);
}

// ---- viz: the colour-mapped images ShowDepthMap / ShowNormalMap / ShowWeakImage write per view-stage when
// viz=True (DPE.cpp:384-503, main.cpp:448-454), rendered on the device as interleaved BGR; JPEG encoding is the
// caller's (nvJPEG, host/io.cpp).  planes = the view's carried (world normal, depth) map.
__global__ void __launch_bounds__(256) k_viz(const float4* __restrict__ planes, const uint8_t* __restrict__ state, float depth_min,
                                             float depth_max, uint8_t* __restrict__ bgr_depth, uint8_t* __restrict__ bgr_normal,
                                             uint8_t* __restrict__ bgr_weak, int n) {
  const float delta = depth_max - depth_min;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = planes[i];
    // ShowDepthMap: blue -> cyan -> green -> yellow -> red ramp over (depth_max - d) / (depth_max - depth_min)
    uint8_t b = 0, g = 0, r = 0;
    if (!(p.w < depth_min || p.w > depth_max || p.w != p.w)) {
      float v = (depth_max - p.w) / delta;
      v = fminf(fmaxf(v, 0.f), 1.f) * 255.f;
      if (v <= 51.f) { b = 255; g = (uint8_t)(v * 5.f); r = 0; }
      else if (v <= 102.f) { v -= 51.f; b = (uint8_t)(255.f - v * 5.f); g = 255; r = 0; }
      else if (v <= 153.f) { v -= 102.f; b = 0; g = 255; r = (uint8_t)(v * 5.f); }
      else if (v <= 204.f) { v -= 153.f; b = 0; g = (uint8_t)(255 - (uint8_t)(v * 128.0 / 51 + 0.5)); r = 255; }
      else { v -= 204.f; b = 0; g = (uint8_t)(127 - (uint8_t)(v * 127.0 / 51 + 0.5)); r = 255; }
    }
    bgr_depth[3 * i] = b; bgr_depth[3 * i + 1] = g; bgr_depth[3 * i + 2] = r;
    // ShowNormalMap: normalised normal * 127.5 + 127.5, rounded and saturated (cv::Mat::convertTo)
    const float norm = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);
    const float inv = norm == 0.f ? 0.f : 1.0f / norm;
    const float c3[3] = {p.x * inv, p.y * inv, p.z * inv};
#pragma unroll
    for (int k = 0; k < 3; ++k) bgr_normal[3 * i + k] = (uint8_t)fminf(fmaxf(rintf(c3[k] * 127.5f + 127.5f), 0.f), 255.f);
    // ShowWeakImage: WEAK white, STRONG green, UNKNOWN red
    const uint8_t st = state[i];
    bgr_weak[3 * i] = st == DPE_WEAK ? 255 : 0;
    bgr_weak[3 * i + 1] = st == DPE_UNKNOWN ? 0 : 255;
    bgr_weak[3 * i + 2] = st == DPE_STRONG ? 0 : 255;
  }
}
void launch_viz(const float4* planes, const uint8_t* state, float depth_min, float depth_max, uint8_t* bgr_depth, uint8_t* bgr_normal,
                uint8_t* bgr_weak, int n, const LaunchCfg& cfg, cudaStream_t stream) {
  k_viz<<<cfg.num_sms * 8, 256, 0, stream>>>(planes, state, depth_min, depth_max, bgr_depth, bgr_normal, bgr_weak, n);
  count(cfg);
}

// ---- scene preparation ---------------------------------------------------------------------
__global__ void k_u8_to_f32(const uint8_t* __restrict__ src, float* __restrict__ dst, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = (float)src[i];
}
void launch_u8_to_f32(const uint8_t* src, float* dst, int n, const LaunchCfg& cfg, cudaStream_t stream) {
  k_u8_to_f32<<<cfg.num_sms * 8, 256, 0, stream>>>(src, dst, n);
  count(cfg);
}

// cv::resize INTER_LINEAR for CV_32FC1: source coordinate (d + 0.5) * scale - 0.5 with
// scale = src/dst computed in double, floor, clamp of the tap index at the borders,
// horizontal then vertical blend in float (the order OpenCV's separable resize uses).
__global__ void k_resize_linear(const float* __restrict__ src, int sw, int sh, float* __restrict__ dst, int dw,
                                int dh) {
  const double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < dw * dh; i += gridDim.x * blockDim.x) {
    const int dx = i % dw, dy = i / dw;
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = (int)floorf(fx);
    fx -= sx;
    if (sx < 0) { fx = 0.f; sx = 0; }
    if (sx >= sw - 1) { fx = 0.f; sx = sw - 1; }
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = (int)floorf(fy);
    fy -= sy;
    if (sy < 0) { fy = 0.f; sy = 0; }
    if (sy >= sh - 1) { fy = 0.f; sy = sh - 1; }
    const int sx1 = imin(sx + 1, sw - 1), sy1 = imin(sy + 1, sh - 1);
    const float a0 = 1.f - fx, a1 = fx, b0 = 1.f - fy, b1 = fy;
    // products and sums rounded separately, like the host code this restates (cv::resize on the CPU in the
    // reference, DPE.cpp:808): a fused multiply-add here changes the last bit of the coarse images whenever
    // the scale is not exactly 2, and the raw-moment NCC turns that into 1e-4 of cost
    const float r0 = __fadd_rn(__fmul_rn(src[(size_t)sy * sw + sx], a0), __fmul_rn(src[(size_t)sy * sw + sx1], a1));
    const float r1 = __fadd_rn(__fmul_rn(src[(size_t)sy1 * sw + sx], a0), __fmul_rn(src[(size_t)sy1 * sw + sx1], a1));
    dst[i] = __fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, b1));
  }
}
// Relative pose of every (reference, source) pair of the scene, computed by the DEVICE compiler from the same
// expressions ComputeHomography evaluates per call (DPE.cu:455-481): the reference's nvcc contracts these
// multiply-adds, a host compiler would not, and the last bit of R_rel / t_rel is enough to move a tap across a
// 1/256 filter-weight bin.  in: n x (ref R[9], ref t[3], src R[9], src t[3], ref c[3], src c[3]);
// out: n x (R_rel[9], t_rel[3], baseline).
__global__ void k_relative_pose(const float* __restrict__ in, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* c = in + (size_t)i * 30;
  float* o = out + (size_t)i * 13;
  relative_pose_ref(c, c + 9, c + 12, c + 21, o, o + 9);
  // baseline of the pair as DepthToWeak / LocalRefine form it per pixel (DPE.cu:2640-2645, 2783-2788): float
  // differences of the stored camera centres, the squared norm promoted to double, sqrtf of it — which under
  // --use_fast_math is the approximate single-precision root; a host sqrtf can differ from it in the last bit, and
  // the disparity steps of both kernels (and with them the depths LocalRefine writes) hang on that bit
  const float* rc3 = c + 24;
  const float* sc3 = c + 27;
  float c_dist[3];
  c_dist[0] = rc3[0] - sc3[0];
  c_dist[1] = rc3[1] - sc3[1];
  c_dist[2] = rc3[2] - sc3[2];
  double temp_val = c_dist[0] * c_dist[0] + c_dist[1] * c_dist[1] + c_dist[2] * c_dist[2];
  o[12] = sqrtf(temp_val);
}
void launch_relative_pose(const float* in, float* out, int n, cudaStream_t stream) {
  k_relative_pose<<<(n + 127) / 128, 128, 0, stream>>>(in, out, n);
}

void launch_resize_linear(const float* src, int sw, int sh, float* dst, int dw, int dh, const LaunchCfg& cfg,
                          cudaStream_t stream) {
  k_resize_linear<<<cfg.num_sms * 8, 256, 0, stream>>>(src, sw, sh, dst, dw, dh);
  count(cfg);
}

// ---- gate-1 hooks ---------------------------------------------------------------------------
// mode 0: the product path (shared patch table + hardware bilinear).  mode 1: same
// arithmetic, but each source tap is an exact fp32 bilinear blend of 4 point-sampled
// texels (texture `point_tex` array is not needed: tex2Dgather is not available for
// float textures with linear filtering, so the 4 taps are fetched at texel centres, where
// the linear filter returns the texel itself).
struct DevEnvExact {
  const float2* tbl;
  int slot;
  __device__ __forceinline__ const RefConst& rc() const { return c_rc[slot]; }
  __device__ __forceinline__ const SrcConst& src(int v) const { return c_rc[slot].src[v]; }
  __device__ __forceinline__ float tex(const SrcConst& sc, float u, float v) const {
    const float xb = u - 0.5f, yb = v - 0.5f;
    const float fx0 = floorf(xb), fy0 = floorf(yb);
    const float ax = xb - fx0, ay = yb - fy0;
    const cudaTextureObject_t t = c_scale_tex;
    const int l = sc.src_view;
    const float t00 = tex2DLayered<float>(t, fx0 + 0.5f, fy0 + 0.5f, l), t10 = tex2DLayered<float>(t, fx0 + 1.5f, fy0 + 0.5f, l);
    const float t01 = tex2DLayered<float>(t, fx0 + 0.5f, fy0 + 1.5f, l), t11 = tex2DLayered<float>(t, fx0 + 1.5f, fy0 + 1.5f, l);
    return (1.f - ay) * ((1.f - ax) * t00 + ax * t10) + ay * ((1.f - ax) * t01 + ax * t11);
  }
  __device__ __forceinline__ float2 pw(int t) const { return tbl[t * NT]; }
};

__global__ void __launch_bounds__(NT) k_cost_eval(const __grid_constant__ StageArgs A, int n_pix,
                                                  const int* __restrict__ xy, const float4* __restrict__ planes,
                                                  int mode, float* __restrict__ out) {
  __shared__ float2 s_tbl[36 * NT];
  const RefConst& rc = c_rc[A.slot];
  const int i = blockIdx.x * NT + threadIdx.x;
  if (i >= n_pix) return;
  const int x = xy[2 * i], y = xy[2 * i + 1];
  GlobalRef ref{A.ref_img, A.W, A.H};
  TblStore st{s_tbl + threadIdx.x};
  const PatchStats ps = build_patch(ref, x, y, st, A.cost_raw != 0, A.exact != 0);
  const float3 m = plane_to_m(rc, planes[i]);
  if (mode == 0) {
    DevEnv env{s_tbl + threadIdx.x, A.ref_img, A.W, A.H, A.slot};
    for (int v = 0; v < rc.n_src; ++v) out[(size_t)i * rc.n_src + v] = ncc_old(env, ps, v, planes[i], m, x, y);
  } else {
    DevEnvExact env{s_tbl + threadIdx.x, A.slot};
    for (int v = 0; v < rc.n_src; ++v) out[(size_t)i * rc.n_src + v] = ncc_old(env, ps, v, planes[i], m, x, y);
  }
}
void launch_cost_eval(const KernelParams& P, int n_pix, const int* xy, const float4* planes, int mode,
                      unsigned long long, float* out, const LaunchCfg& cfg, cudaStream_t stream) {
  k_cost_eval<<<(n_pix + NT - 1) / NT, NT, 0, stream>>>(P.a, n_pix, xy, planes, mode, out);
  count(cfg);
}

__global__ void k_geom_eval(const __grid_constant__ StageArgs A, int n_pix, const int* __restrict__ xy,
                            const float4* __restrict__ planes, float* __restrict__ out) {
  const RefConst& rc = c_rc[A.slot];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pix) return;
  for (int v = 0; v < rc.n_src; ++v)
    out[(size_t)i * rc.n_src + v] = geom_cost(A, rc, rc.src[v], planes[i], xy[2 * i], xy[2 * i + 1]);
}
void launch_geom_eval(const KernelParams& P, int n_pix, const int* xy, const float4* planes, float* out,
                      const LaunchCfg& cfg, cudaStream_t stream) {
  k_geom_eval<<<(n_pix + 127) / 128, 128, 0, stream>>>(P.a, n_pix, xy, planes, out);
  count(cfg);
}

// ---- NCC throughput study --------------------------------------------------------------------
// The candidate loop of the strong sweep in isolation: every pixel of a view scores the planes
// of n_cand neighbours (carried (world normal, depth) maps of the last stage) against all source
// views.  ROWS = tap rows fetched before any is consumed (texture results in flight per thread =
// 6 * ROWS); MINB = CTAs per SM the register allocation is tuned for.
template <int ROWS, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_ncc_bench(const __grid_constant__ StageArgs A, const float4* __restrict__ world_planes,
                                                        int n_cand, float* __restrict__ out) {
  __shared__ float2 s_tbl[36 * NT];
  __shared__ float s_tile[SMW * (8 + 2 * HALO)];
  StageArgs a = A;
  const RefConst& rc = c_rc[A.slot];
  const int tiles_x = (a.W + TILE_W - 1) / TILE_W, tiles_y = (a.H + 7) / 8;
  RefTile<8> tile;
  tile.s = s_tile;
  DevEnv env;
  env.tbl = s_tbl + threadIdx.x;
  env.img = a.ref_img; env.W = a.W; env.H = a.H; env.slot = a.slot;
  TblStore st;
  st.tbl = s_tbl + threadIdx.x;
  const int offx[8] = {0, 0, -1, 1, -3, 3, 5, -7}, offy[8] = {-1, 1, 0, 0, 5, -5, 3, 2};
  for (int t = blockIdx.x; t < tiles_x * tiles_y; t += gridDim.x) {
    const int tx0 = (t % tiles_x) * TILE_W, ty0 = (t / tiles_x) * 8;
    __syncthreads();
    tile.stage(a.ref_img, a.W, a.H, tx0, ty0);
    __pipeline_wait_prior(0);
    __syncthreads();
    const int lx = threadIdx.x & 31;
    const int x = tx0 + lx, y = ty0 + 2 * (threadIdx.x >> 5) + (lx & 1);
    if (x < a.W && y < a.H) {
      const PatchStats ps = build_patch(tile, x, y, st, a.cost_raw != 0, a.exact != 0);
      float acc = 0.f;
      for (int c = 0; c < n_cand; ++c) {
        const int nx = iclamp(x + offx[c & 7], 0, a.W - 1), ny = iclamp(y + offy[c & 7], 0, a.H - 1);
        const float4 pw = world_planes[ny * a.W + nx];
        float4 pl = world_to_cam_normal(rc, pw);
        pl.w = dist2origin(rc, nx, ny, pw.w > 0.f ? pw.w : 1.0f, pl);
        const float3 m = plane_to_m(rc, pl);
        for (int v = 0; v < rc.n_src; ++v) {
          acc += ncc_old(env, ps, v, pl, m, x, y);
        }
      }
      out[y * a.W + x] = acc;
    }
  }
}
void launch_ncc_bench(const KernelParams& P, const float4* world_planes, int n_cand, int variant, float* out,
                      const LaunchCfg& cfg, cudaStream_t stream) {
  const int tiles = ((P.a.W + TILE_W - 1) / TILE_W) * ((P.a.H + 7) / 8);
  auto g = [&](int per_sm) { return persistent_grid(tiles, cfg.num_sms, per_sm); };
  switch (variant) {
    case 0: k_ncc_bench<0, 4><<<g(4), NT, 0, stream>>>(P.a, world_planes, n_cand, out); break;
    case 1: k_ncc_bench<0, 3><<<g(3), NT, 0, stream>>>(P.a, world_planes, n_cand, out); break;
    default: break;
  }
  count(cfg);
}

// ---- micro-benchmarks (roofline denominators) ----------------------------------------------
// Filtered fetch rate with an access pattern like the NCC's: each lane walks a 6x6 tap
// grid with ~1-pixel steps around a base that differs by one pixel between lanes.
__global__ void k_probe_tex(cudaTextureObject_t tex, int w, int h, int iters, float* sink) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  float bx = (float)((gid * 1) % (w - 16)) + 0.37f;
  float by = (float)(((gid / 32) * 2) % (h - 16)) + 0.61f;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 6; ++j)
#pragma unroll
      for (int i = 0; i < 6; ++i) acc += tex2D<float>(tex, bx + 1.93f * i, by + 1.97f * j);
    bx += 0.11f; by += 0.07f;
    if (bx > w - 16) bx -= (w - 32);
    if (by > h - 16) by -= (h - 32);
  }
  if (acc == 1234.5678f) sink[0] = acc;
}
void launch_probe_tex(unsigned long long tex, int w, int h, int iters, float* sink, int blocks, int threads,
                      const LaunchCfg& cfg, cudaStream_t stream) {
  k_probe_tex<<<blocks, threads, 0, stream>>>((cudaTextureObject_t)tex, w, h, iters, sink);
  count(cfg);
}
// Lane-layout / footprint study: `layout` maps the 32 lanes of a warp to reference pixels
// (0: 32 consecutive x; 1: red/black zig-zag over two rows, the half-sweep layout; 2: 8x4 block;
// 3: one colour of an 8-wide x 8-tall block; 4: 16x2 block), the 2x2 matrix m maps reference
// offsets to source offsets (rotation / scale / shear of a plane-induced homography), taps are
// the NCC's {-5,-3,..,5}^2.
__global__ void k_probe_tex_pattern(cudaTextureObject_t tex, int w, int h, int iters, int layout, float m00, float m01,
                                    float m10, float m11, float* sink) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lx, ly;
  if (layout == 0) { lx = lane; ly = 0; }
  else if (layout == 1) { lx = lane; ly = lane & 1; }
  else if (layout == 2) { lx = lane & 7; ly = lane >> 3; }
  else if (layout == 3) { lx = lane & 7; ly = 2 * (lane >> 3) + (lane & 1); }
  else { lx = lane & 15; ly = lane >> 4; }
  const int span = w - 96;
  float bx = 48.f + (float)((warp * 37) % span) + lx * m00 + ly * m01 + 0.37f;
  float by = 48.f + (float)((warp * 11) % (h - 96)) + lx * m10 + ly * m11 + 0.61f;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = -5; j <= 5; j += 2)
#pragma unroll
      for (int i = -5; i <= 5; i += 2) acc += tex2D<float>(tex, bx + m00 * i + m01 * j, by + m10 * i + m11 * j);
    bx += 0.11f; by += 0.07f;
  }
  if (acc == 1234.5678f) sink[0] = acc;
}
void launch_probe_tex_pattern(unsigned long long tex, int w, int h, int iters, int layout, const float m[4], float* sink,
                              int blocks, int threads, const LaunchCfg& cfg, cudaStream_t stream) {
  k_probe_tex_pattern<<<blocks, threads, 0, stream>>>((cudaTextureObject_t)tex, w, h, iters, layout, m[0], m[1], m[2], m[3], sink);
  count(cfg);
}
__global__ void k_probe_fma(int iters, float* sink) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
  float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float b = 1.000001f, c = 1e-7f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
      a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
    }
  }
  const float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 1234.5678f) sink[0] = s;
}
void launch_probe_fma(int iters, float* sink, int blocks, int threads, const LaunchCfg& cfg, cudaStream_t stream) {
  k_probe_fma<<<blocks, threads, 0, stream>>>(iters, sink);
  count(cfg);
}
// texture is 2x1: texel 0 = 0.0, texel 1 = 1.0; sample at x = 0.5 + i/n
__global__ void k_probe_weights(cudaTextureObject_t tex, int n, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n) out[i] = tex2D<float>(tex, 0.5f + (float)i / (float)n, 0.5f);
}
void launch_probe_weights(unsigned long long tex, int n, float* out, const LaunchCfg& cfg, cudaStream_t stream) {
  k_probe_weights<<<(n + 256) / 256, 256, 0, stream>>>((cudaTextureObject_t)tex, n, out);
  count(cfg);
}

}  // namespace dpe
