// dpe_hostsim.cu — TEST INFRASTRUCTURE: CPU simulator of the kernel logic (built by oracle/Makefile).
//
// This is not a product path and not an oracle: it compiles the same per-pixel
// __host__ __device__ code as the CUDA kernels (dpe_core.cuh / dpe_weak.cuh) for the host,
// with a software model of the texture unit, so that the *logic* of a stage (sampling
// patterns, view selection, refinement, classifier, weak path) can be exercised by the
// `-m "not gpu"` tests on a box without a GPU.  It is built into its own library
// (libdpe_hostsim.so) that only tests/ loads; nothing in the product links or falls back
// to it — the C-ABI library fails with DPE_ERR_NO_DEVICE when there is no GPU.
#include <stdint.h>
#include <stdlib.h>
#include <vector>
#include "dpe_core.cuh"
#include "dpe_weak.cuh"
#include "dpe_consts.h"

using namespace dpe;

namespace {

struct HostImage {
  const float* d;
  int W, H;
  int quant;  // 0: exact fp32 weights, 1: weights rounded to 1/256, 2: truncated to 1/256
};

inline float host_tex(const HostImage* im, float u, float v) {
  // CUDA linear filtering, unnormalised coordinates, clamp addressing:
  // xB = x - 0.5, i = floor(xB), alpha = frac(xB) in 1.8 fixed point
  const float xb = u - 0.5f, yb = v - 0.5f;
  const float fx0 = floorf(xb), fy0 = floorf(yb);
  float ax = xb - fx0, ay = yb - fy0;
  if (im->quant == 1) { ax = floorf(ax * 256.f + 0.5f) * (1.f / 256.f); ay = floorf(ay * 256.f + 0.5f) * (1.f / 256.f); }
  else if (im->quant == 2) { ax = floorf(ax * 256.f) * (1.f / 256.f); ay = floorf(ay * 256.f) * (1.f / 256.f); }
  const int x0 = iclamp((int)fx0, 0, im->W - 1), x1 = iclamp((int)fx0 + 1, 0, im->W - 1);
  const int y0 = iclamp((int)fy0, 0, im->H - 1), y1 = iclamp((int)fy0 + 1, 0, im->H - 1);
  if (!(u == u) || !(v == v)) return 0.f;
  const float t00 = im->d[(size_t)y0 * im->W + x0], t10 = im->d[(size_t)y0 * im->W + x1];
  const float t01 = im->d[(size_t)y1 * im->W + x0], t11 = im->d[(size_t)y1 * im->W + x1];
  return (1.f - ay) * ((1.f - ax) * t00 + ax * t10) + ay * ((1.f - ax) * t01 + ax * t11);
}

struct HostEnv {
  float2 tbl[36];
  const float* img;
  int W, H;
  const RefConst* rcp;
  const RefConst& rc() const { return *rcp; }
  const SrcConst& src(int v) const { return rcp->src[v]; }
  float tex(const SrcConst& sc, float u, float v) const { return host_tex((const HostImage*)sc.tex, u, v); }
  float2 pw(int t) const { return tbl[t]; }
  float ref(int x, int y) const { return img[(size_t)iclamp(y, 0, H - 1) * W + iclamp(x, 0, W - 1)]; }
};

struct HostRef {
  const float* img; int W, H;
  float operator()(int x, int y) const { return img[(size_t)iclamp(y, 0, H - 1) * W + iclamp(x, 0, W - 1)]; }
};
struct HostStore {
  float2* tbl;
  void operator()(int t, float w, float wr) const { tbl[t] = make_float2(w, wr); }
};

}  // namespace

extern "C" {

// Bilateral NCC of fixed hypotheses through the simulator (used to cross-check the
// simulator itself against the float64 oracle).
int dpe_hostsim_cost_eval(int W, int H, int full_w, int full_h, const float* ref_img, int n_src,
                          const float* const* src_imgs, const float* K, const float* R, const float* t, int quant,
                          int n_pix, const int* xy, const float* planes, float* out) {
  std::vector<HostCam> cams(n_src + 1);
  for (int i = 0; i <= n_src; ++i) host_cam_set(&cams[i], K + 9 * i, R + 9 * i, t + 3 * i, 1.f, 2.f);
  RefConst* rc = new RefConst();
  fold_ref(cams[0], W, H, full_w, full_h, 0, rc);
  rc->n_src = n_src;
  std::vector<HostImage> imgs(n_src);
  for (int i = 0; i < n_src; ++i) {
    fold_pair(cams[0], cams[i + 1], W, H, full_w, full_h, &rc->src[i]);
    imgs[i] = HostImage{src_imgs[i], W, H, quant};
    rc->src[i].tex = (unsigned long long)&imgs[i];
  }
  for (int i = 0; i < n_pix; ++i) {
    HostEnv env; env.img = ref_img; env.W = W; env.H = H; env.rcp = rc;
    HostRef ref{ref_img, W, H};
    HostStore st{env.tbl};
    const int x = xy[2 * i], y = xy[2 * i + 1];
    const PatchStats ps = build_patch(ref, x, y, st, getenv("DPE_HOSTSIM_CENTRED") == nullptr, getenv("DPE_HOSTSIM_EXACT") != nullptr);
    const float4 pl = make_float4(planes[4 * i], planes[4 * i + 1], planes[4 * i + 2], planes[4 * i + 3]);
    const float3 m = plane_to_m(*rc, pl);
    for (int v = 0; v < n_src; ++v) out[(size_t)i * n_src + v] = ncc_old(env, ps, v, pl, m, x, y);
  }
  delete rc;
  return 0;
}

// One (view, stage) of the PatchMatch path on the CPU.  images: n_src+1 float images at
// this stage's scale (index 0 = reference); cameras likewise.  src_depths: n_src maps or
// NULL.  prev_*: maps of the previous stage (NULL for FIRST_INIT).  Outputs are W*H.
// Raw state after a step (per-kernel differential tests against oracle/ref_stage_probe.cu).  Steps:
// 0 anchors (GenNeighbours + NeigbourUpdate), 1 init, 2+3i strong sweeps of iteration i, 3+3i fit plane,
// 4+3i weak sweeps, 11 final.
struct HostsimDbg {
  int stop_step;
  float* planes; float* costs; uint32_t* selected; uint8_t* state; float* fit; int32_t* radius;
  int16_t* neighbours;  // P x 9 x 2
  uint8_t* reliable;
};

static int hostsim_stage_impl(int W, int H, int full_w, int full_h, int n_src, const float* const* images, const float* K,
                      const float* R, const float* t, float depth_min, float depth_max,
                      const float* const* src_depths, const float* prev_planes, const uint8_t* prev_state,
                      const uint32_t* prev_selected, int prev_W, int prev_H, const uint8_t* edge,
                      const uint8_t* edge_low, int low_w, int low_h, const int32_t* label,
                      const dpe_stage_params* p, uint64_t seed, int view, uint32_t stage_counter, int quant,
                      float* out_planes, uint8_t* out_state, uint32_t* out_selected, float* out_depth,
                      double* eval_units, const HostsimDbg* dbg) {
  if (n_src > DPE_MAX_SRC) return DPE_ERR_TOO_MANY_IMAGES;
  const size_t P = (size_t)W * H;
  std::vector<HostCam> cams(n_src + 1);
  for (int i = 0; i <= n_src; ++i) host_cam_set(&cams[i], K + 9 * i, R + 9 * i, t + 3 * i, depth_min, depth_max);
  RefConst* rc = new RefConst();
  fold_ref(cams[0], W, H, full_w, full_h, view, rc);
  rc->n_src = n_src;
  std::vector<HostImage> imgs(n_src);
  for (int i = 0; i < n_src; ++i) {
    fold_pair(cams[0], cams[i + 1], W, H, full_w, full_h, &rc->src[i]);
    imgs[i] = HostImage{images[i + 1], W, H, quant};
    rc->src[i].tex = (unsigned long long)&imgs[i];
    rc->src[i].depth = (p->geom_consistency && src_depths) ? src_depths[i] : nullptr;
    rc->src[i].src_view = i;
  }
  std::vector<float4> planes(P), fit(P), outp(P);
  std::vector<float> costs(P), complexity(P), atlas(P);
  std::vector<uint32_t> selected(P), outsel(P);
  std::vector<uint4> vw(P);
  std::vector<uint8_t> state(P), reliable(P), outst(P), zero_edge;
  std::vector<int> radius(P, 0);
  std::vector<int32_t> neg_label;
  std::vector<short2> edge_neigh(P * 8), label_boundary(P * 8), nearest(P), neighbours(P * DPE_NEIGHBOUR_NUM);
  if (!edge) { zero_edge.assign(P, 0); edge = zero_edge.data(); }
  if (!edge_low) { edge_low = edge; low_w = W; low_h = H; }
  if (!label) { neg_label.assign(P, -1); label = neg_label.data(); }

  StageArgs a;
  memset(&a, 0, sizeof(a));
  a.rc = rc; a.ref_img = images[0]; a.W = W; a.H = H;
  a.planes = planes.data(); a.costs = costs.data(); a.selected = selected.data(); a.view_w = vw.data();
  a.state = state.data(); a.fit_planes = fit.data(); a.radius = radius.data(); a.edge = edge; a.edge_low = edge_low;
  a.low_w = low_w; a.low_h = low_h; a.edge_neigh = edge_neigh.data(); a.complexity = complexity.data();
  a.label = label; a.label_boundary = label_boundary.data(); a.weak_reliable = reliable.data();
  a.nearest_strong = nearest.data(); a.neighbours = neighbours.data();
  a.prev_planes = (const float4*)prev_planes; a.prev_state = prev_state; a.prev_selected = prev_selected;
  a.prev_W = prev_planes ? prev_W : W; a.prev_H = prev_planes ? prev_H : H;
  a.out_planes = outp.data(); a.out_state = outst.data(); a.out_selected = outsel.data(); a.atlas_out = atlas.data();
  a.run_state = p->state; a.geom = p->geom_consistency; a.use_apd = p->use_apd; a.top_k = p->top_k;
  a.weak_peak_radius = p->weak_peak_radius; a.rotate_time = p->rotate_time; a.ransac_threshold = p->ransac_threshold;
  a.geom_factor = p->geom_factor;
  a.ref_race = getenv("DPE_HOSTSIM_REF_RACE") ? 1 : 0;
  a.cost_raw = getenv("DPE_HOSTSIM_CENTRED") ? 0 : 1;
  a.exact = getenv("DPE_HOSTSIM_EXACT") ? 1 : 0;
  a.debug_accept = nullptr;
  (void)stage_counter;
  std::vector<Xorwow> rng_states(P);
  xorwow_init_table(seed, W, H, rng_states.data());
  a.rng = rng_states.data();
  double units = 0.0;
  auto stop_at = [&](int step) {
    if (!dbg || dbg->stop_step != step) return false;
    memcpy(dbg->planes, planes.data(), P * sizeof(float4));
    memcpy(dbg->costs, costs.data(), P * sizeof(float));
    memcpy(dbg->selected, selected.data(), P * sizeof(uint32_t));
    memcpy(dbg->state, state.data(), P);
    memcpy(dbg->fit, fit.data(), P * sizeof(float4));
    memcpy(dbg->radius, radius.data(), P * sizeof(int));
    memcpy(dbg->neighbours, neighbours.data(), P * DPE_NEIGHBOUR_NUM * sizeof(short2));
    memcpy(dbg->reliable, reliable.data(), P);
    return true;
  };

  HostRef ref{images[0], W, H};
  auto for_all = [&](auto&& fn) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) fn(x, y);
  };
  auto for_colour = [&](int colour, auto&& fn) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; ++y)
      for (int x = (y + colour) & 1; x < W; x += 2) fn(x, y);
  };
  auto with_patch = [&](int x, int y, auto&& fn) {
    HostEnv env; env.img = images[0]; env.W = W; env.H = H; env.rcp = a.rc;
    HostStore st{env.tbl};
    const PatchStats ps = build_patch(ref, x, y, st, a.cost_raw != 0, a.exact != 0);
    unsigned ev = 0;
    fn(env, ps, ev);
    if (ev) {
#pragma omp atomic
      units += (double)ev;
    }
  };

  if (p->state != DPE_FIRST_INIT) for_all([&](int x, int y) { load_pixel(a, x, y); });
  if (p->use_apd) {
    for_all([&](int x, int y) { edge_info_pixel(a, x, y); label_boundary_pixel(a, x, y); });
    for_all([&](int x, int y) { nearest_strong_pixel(a, x, y); });
    for_all([&](int x, int y) { gen_neighbours_pixel(a, x, y); });
  }
  if (stop_at(0)) { delete rc; return 0; }
  for_all([&](int x, int y) { with_patch(x, y, [&](HostEnv& env, const PatchStats& ps, unsigned& ev) { init_pixel(env, ps, a, x, y, ev); }); });
  if (stop_at(1)) { delete rc; return 0; }
  for (int it = 0; it < p->max_iterations; ++it) {
    a.iter = it;
    for (int colour = 0; colour < 2; ++colour) {
      a.colour = colour;
      for_colour(colour, [&](int x, int y) {
        if (a.state[y * W + x] == DPE_WEAK) return;
        with_patch(x, y, [&](HostEnv& env, const PatchStats& ps, unsigned& ev) {
          float cost_arr[9 * DPE_MAX_IMAGES];
          if (a.use_apd) strong_update_pixel<true>(env, ps, a, x, y, cost_arr, ev);
          else strong_update_pixel<false>(env, ps, a, x, y, cost_arr, ev);
        });
      });
    }
    if (stop_at(2 + 3 * it)) { delete rc; return 0; }
    if (p->use_apd) {
      for_all([&](int x, int y) { fit_plane_pixel(a, x, y); });
      if (stop_at(3 + 3 * it)) { delete rc; return 0; }
      for (int colour = 0; colour < 2; ++colour) {
        a.colour = colour;
        for_colour(colour, [&](int x, int y) {
          if (a.state[y * W + x] != DPE_WEAK) return;
          with_patch(x, y, [&](HostEnv& env, const PatchStats& ps, unsigned& ev) {
            float cost_arr[9 * DPE_MAX_IMAGES];
            weak_update_pixel(env, ps, a, x, y, cost_arr, ev);
          });
        });
      }
    }
    if (stop_at(4 + 3 * it)) { delete rc; return 0; }
  }
  for_all([&](int x, int y) { extract_pixel(a, x, y); });
  for (int colour = 0; colour < 2; ++colour) {
    a.colour = colour;
    for_colour(colour, [&](int x, int y) {
      if (a.state[y * W + x] == DPE_WEAK) return;
      median_pixel(a, x, y);
    });
  }
  for_all([&](int x, int y) { with_patch(x, y, [&](HostEnv& env, const PatchStats& ps, unsigned& ev) { classify_refine_pixel(env, ps, a, x, y, ev); }); });
  for_all([&](int x, int y) { finish_pixel(a, x, y); });
  if (dbg && dbg->stop_step == 11) {
    memcpy(dbg->planes, outp.data(), P * sizeof(float4));
    memcpy(dbg->selected, outsel.data(), P * sizeof(uint32_t));
    memcpy(dbg->state, outst.data(), P);
  }

  memcpy(out_planes, outp.data(), P * sizeof(float4));
  memcpy(out_state, outst.data(), P);
  memcpy(out_selected, outsel.data(), P * sizeof(uint32_t));
  if (out_depth) memcpy(out_depth, atlas.data(), P * sizeof(float));
  if (eval_units) *eval_units = units;
  delete rc;
  return 0;
}

int dpe_hostsim_stage(int W, int H, int full_w, int full_h, int n_src, const float* const* images, const float* K,
                      const float* R, const float* t, float depth_min, float depth_max,
                      const float* const* src_depths, const float* prev_planes, const uint8_t* prev_state,
                      const uint32_t* prev_selected, int prev_W, int prev_H, const uint8_t* edge,
                      const uint8_t* edge_low, int low_w, int low_h, const int32_t* label,
                      const dpe_stage_params* p, uint64_t seed, int view, uint32_t stage_counter, int quant,
                      float* out_planes, uint8_t* out_state, uint32_t* out_selected, float* out_depth,
                      double* eval_units) {
  return hostsim_stage_impl(W, H, full_w, full_h, n_src, images, K, R, t, depth_min, depth_max, src_depths, prev_planes,
                            prev_state, prev_selected, prev_W, prev_H, edge, edge_low, low_w, low_h, label, p, seed, view,
                            stage_counter, quant, out_planes, out_state, out_selected, out_depth, eval_units, nullptr);
}

// same, stopping after `stop_step` and returning the raw state at that point
int dpe_hostsim_stage_dbg(int W, int H, int full_w, int full_h, int n_src, const float* const* images, const float* K,
                          const float* R, const float* t, float depth_min, float depth_max,
                          const float* const* src_depths, const float* prev_planes, const uint8_t* prev_state,
                          const uint32_t* prev_selected, int prev_W, int prev_H, const uint8_t* edge,
                          const uint8_t* edge_low, int low_w, int low_h, const int32_t* label,
                          const dpe_stage_params* p, uint64_t seed, int quant, int stop_step, float* planes, float* costs,
                          uint32_t* selected, uint8_t* state, float* fit, int32_t* radius, int16_t* neighbours,
                          uint8_t* reliable) {
  const size_t P = (size_t)W * H;
  std::vector<float> op(P * 4), od(P);
  std::vector<uint8_t> os(P);
  std::vector<uint32_t> ol(P);
  double u = 0;
  HostsimDbg dbg{stop_step, planes, costs, selected, state, fit, radius, neighbours, reliable};
  return hostsim_stage_impl(W, H, full_w, full_h, n_src, images, K, R, t, depth_min, depth_max, src_depths, prev_planes,
                            prev_state, prev_selected, prev_W, prev_H, edge, edge_low, low_w, low_h, label, p, seed, 0, 0,
                            quant, op.data(), os.data(), ol.data(), od.data(), &u, &dbg);
}

// cv::resize(INTER_LINEAR) on float, host mirror of k_resize_linear (same arithmetic)
void dpe_hostsim_resize_linear(const float* src, int sw, int sh, float* dst, int dw, int dh) {
  const double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
  for (int dy = 0; dy < dh; ++dy) {
    for (int dx = 0; dx < dw; ++dx) {
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
      const float r0 = src[(size_t)sy * sw + sx] * a0 + src[(size_t)sy * sw + sx1] * a1;
      const float r1 = src[(size_t)sy1 * sw + sx] * a0 + src[(size_t)sy1 * sw + sx1] * a1;
      dst[(size_t)dy * dw + dx] = r0 * b0 + r1 * b1;
    }
  }
}

}  // extern "C"
