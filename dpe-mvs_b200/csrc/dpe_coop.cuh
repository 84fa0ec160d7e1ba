// dpe_coop.cuh — warp-cooperative scoring for the two per-pixel functions whose cost loops run over a PER-PIXEL
// set of source views (dpe_core.cuh stays the definition; the CPU simulator runs this file too, with the 32 lanes
// of a warp as fibers, to check that both give the same bits — tests/test_sim_logic.py):
//   * the refinement half of the strong sweep (cost of the current plane + 5 hypotheses over the views the pixel
//     SAMPLED: 2-5 of N), strong_update_coop;
//   * the weak classifier + local refinement (62 depths over the views the pixel SELECTED), classify_refine_coop.
// One thread per pixel walks the union of its warp's view sets with the lanes that did not pick the view idle
// (ncu, round 2: 23.6 of 32 lanes in the classifier, and a TEX instruction costs the texture pipe the same with 8
// lanes as with 32).  Here the warp's (pixel, view) pairs are dealt out densely instead: pairs are numbered
// view-major (all pixels that picked view 0 in lane order, then view 1, ...), a round hands the next <= 32 pairs to
// the lanes — at most COOP_SLOTS different views per round, so neighbouring lanes score neighbouring pixels against
// the same source image — every lane scores ITS PAIR's pixel (position, plane, patch statistics by shuffle from the
// owner, the owner's column of the shared patch table), and the costs go back to the owners by shuffle, who add
// them up in ascending view order: the same float operations in the same order as the per-pixel functions, so the
// results are bit-identical to theirs (tests/test_gpu_parity.py::test_cooperative_scoring_gives_identical_maps).
// A pair is heavy (1, 5 or 62 evaluations of ~800 instructions), so the bookkeeping of a round (~100 instructions)
// and of a cost hand-back (~8 per slot) stays in the per-cent range — the first cooperative classifier (round 2,
// dropped) dealt out single evaluations and paid a third of its time for that.
#pragma once
#include "dpe_core.cuh"

namespace dpe {

constexpr int COOP_SLOTS = 4;  // views a round may span = costs an owner may get back per round

// ---- the warp primitives this file is written in.  Device: the full-mask intrinsics.  Host: the CPU simulator
// (oracle/dpe_hostsim.cu) runs the lanes of a warp as fibers that meet in these three calls.
#ifndef __CUDA_ARCH__
extern "C" int dpe_hostsim_lane();
extern "C" unsigned dpe_hostsim_ballot(int pred);
extern "C" unsigned dpe_hostsim_shfl(unsigned bits, int src_lane);
#endif
DPE_HD int coop_lane() {
#ifdef __CUDA_ARCH__
  return threadIdx.x & 31;
#else
  return dpe_hostsim_lane();
#endif
}
DPE_HD unsigned coop_ballot(const bool pred) {
#ifdef __CUDA_ARCH__
  return __ballot_sync(0xffffffffu, pred);
#else
  return dpe_hostsim_ballot(pred ? 1 : 0);
#endif
}
DPE_HD int coop_shfl(const int v, const int src_lane) {
#ifdef __CUDA_ARCH__
  return __shfl_sync(0xffffffffu, v, src_lane);
#else
  return (int)dpe_hostsim_shfl((unsigned)v, src_lane);
#endif
}
DPE_HD float coop_shfl(const float v, const int src_lane) {
#ifdef __CUDA_ARCH__
  return __shfl_sync(0xffffffffu, v, src_lane);
#else
  union { float f; unsigned u; } a, b;
  a.f = v;
  b.u = dpe_hostsim_shfl(a.u, src_lane);
  return b.f;
#endif
}
DPE_HD int popc32(const unsigned m) {
#ifdef __CUDA_ARCH__
  return __popc(m);
#else
  return __builtin_popcount(m);
#endif
}

// position of the n-th (0-based) set bit of m; m has more than n set bits
DPE_HD int nth_set_bit(const unsigned m, const int n) {
  int pos = 0;
  int left = n;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    const int c = popc32((m >> pos) & ((1u << s) - 1u));
    if (c <= left) { left -= c; pos += s; }
  }
  return pos;
}

// One round of the deal.
struct CoopRound {
  bool has;                // this lane scores a pair in this round
  int p, v;                // its pixel (= owner lane) and source view
  int n_slots;             // uniform: views the round spans
  int src[COOP_SLOTS];     // owner side, per slot in ascending view order: the lane that scores this lane's pair
                           //   with that view, or -1
  int w[COOP_SLOTS];       //   and the owner's weight for the view
};
// where the deal stands (uniform across the warp)
struct CoopCursor {
  int v;     // first view with pairs left
  int done;  // pairs of that view already dealt
};

// views: bit v set = this lane's pixel wants view v scored (0 for a lane without work); vw: the owner's weights.
// Returns false when no pair is left.  Every lane of the warp must call this together.
DPE_HD bool coop_next_round(const uint32_t views, const ViewW& vw, const int N, CoopCursor& cur,
                                                CoopRound& r) {
  const int lane = coop_lane();
  const unsigned lt = (1u << lane) - 1u;
  r.has = false; r.p = 0; r.v = 0; r.n_slots = 0;
#pragma unroll
  for (int s = 0; s < COOP_SLOTS; ++s) { r.src[s] = -1; r.w[s] = 0; }
  int filled = 0;
  int v = cur.v, skip = cur.done;
  while (v < N && filled < 32 && r.n_slots < COOP_SLOTS) {
    const unsigned b = coop_ballot(((views >> v) & 1u) != 0u);
    const int left = popc32(b) - skip;
    if (left <= 0) { ++v; skip = 0; continue; }
    const int take = imin(left, 32 - filled);
    // scoring side: lanes filled .. filled + take - 1 get this view's pairs skip .. skip + take - 1
    if (lane >= filled && lane < filled + take) {
      r.has = true; r.v = v;
      r.p = nth_set_bit(b, skip + lane - filled);
    }
    // owner side
    int src = -1, w = 0;
    if ((b >> lane) & 1u) {
      const int rank = popc32(b & lt);
      if (rank >= skip && rank < skip + take) { src = filled + rank - skip; w = vw.get(v); }
    }
#pragma unroll
    for (int s = 0; s < COOP_SLOTS; ++s)
      if (s == r.n_slots) { r.src[s] = src; r.w[s] = w; }
    ++r.n_slots;
    filled += take;
    if (take < left) { skip += take; break; }  // the view continues in the next round
    ++v; skip = 0;
  }
  cur.v = v; cur.done = skip;
  return filled > 0;
}

DPE_HD float4 shfl_plane(const float4 pl, const int p) {
  return make_float4(coop_shfl(pl.x, p), coop_shfl(pl.y, p), coop_shfl(pl.z, p), coop_shfl(pl.w, p));
}
DPE_HD PatchStats shfl_patch(const PatchStats& ps, const int exact, const int p) {
  PatchStats q;
  q.r0 = coop_shfl(ps.r0, p);
  q.c0 = coop_shfl(ps.c0, p);
  q.exact = exact;  // uniform per launch
  q.inv_sw = coop_shfl(ps.inv_sw, p);
  q.mean_r = coop_shfl(ps.mean_r, p);
  q.var_r = coop_shfl(ps.var_r, p);
  return q;
}

// weighted_cost (dpe_core.cuh) of NH hypotheses of every pixel of the warp over the pixel's sampled views:
// plane_of(h) is the owner's plane for hypothesis h (called by every lane with the same h), out[h] the sum of
// w * cost in ascending view order, still to be divided by the weight norm.
template <int NH, class Env, class PlaneOf>
DPE_HD void coop_weighted_costs(const Env& env, const PatchStats& ps, const RefConst& rc, const int exact,
                                                    const int x, const int y, const uint32_t views, const ViewW& vw,
                                                    const PlaneOf& plane_of, float (&out)[NH], unsigned& evals) {
  const int lane = coop_lane();
#pragma unroll
  for (int h = 0; h < NH; ++h) out[h] = 0.f;
  CoopCursor cur;
  cur.v = 0; cur.done = 0;
  CoopRound r;
  while (coop_next_round(views, vw, rc.n_src, cur, r)) {
    const int px = coop_shfl(x, r.p), py = coop_shfl(y, r.p);
    const PatchStats pps = shfl_patch(ps, exact, r.p);
    Env penv = env;
    penv.tbl = env.tbl + (r.p - lane);
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      const float4 hp = shfl_plane(plane_of(h), r.p);
      float cv = 0.f;
      if (r.has) {
        const float3 m = plane_to_m(rc, hp);
        cv = ncc_old(penv, pps, r.v, hp, m, px, py);
        evals += cv < 2.0f;
      }
      for (int s = 0; s < r.n_slots; ++s) {
        int src = -1, w = 0;
#pragma unroll
        for (int t = 0; t < COOP_SLOTS; ++t)
          if (t == s) { src = r.src[t]; w = r.w[t]; }
        const float c = coop_shfl(cv, src < 0 ? 0 : src);
        if (src >= 0) out[h] += w * c;  // weighted_cost: c += w * cv
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// strong_update_pixel (dpe_core.cuh) for a whole warp; act = this lane has a pixel to update
// ------------------------------------------------------------------------------------
template <bool EDGE, class Env>
DPE_HDN void strong_update_coop(const Env& env, const PatchStats& ps, const StageArgs& a, const int x, const int y,
                                   const bool act, float* cost_arr, unsigned& evals) {
  const RefConst& rc = env.rc();
  const int center = y * a.W + x;
  StrongPick s;
  s.vw.clear(); s.weight_norm = 0.f; s.sel_bits = 0u; s.min_idx = 0; s.fl = false; s.pos = 0; s.fc = 0.f;
  float4 plane_now = make_float4(0.f, 0.f, 1.f, 1.f);
  if (act) {
    bool flag[8];
    int positions[8];
    strong_candidates<EDGE>(env, ps, a, x, y, cost_arr, evals, flag, positions);
    strong_select(a, rc.n_src, center, flag, positions, cost_arr, s);
    plane_now = a.planes[center];
  }
  // the views weighted_cost walks: sampled weight > 0
  const uint32_t views = act ? s.sel_bits : 0u;

  float c1[1];
  coop_weighted_costs<1>(env, ps, rc, a.exact, x, y, views, s.vw, [&](int) { return plane_now; }, c1, evals);
  float cost_now = c1[0] / s.weight_norm;
  const float cost_before = cost_now;
  int accepted = 0;
  float depth_now = 0.f;
  RefineDraws rd;
  rd.depth_rand = rd.depth_pert = rd.depth_in = 1.f;
  rd.n_rand = rd.n_pert = rd.plane_in = plane_now;
  if (act) {
    depth_now = depth_from_plane(rc, plane_now, x, y);
    strong_accept<EDGE>(a, rc, s, x, y, center, plane_now, depth_now, cost_now, accepted);
    rd = refine_draws(rc, plane_now, depth_now, s.rng, x, y);
  }
  float c5[5];
  coop_weighted_costs<5>(env, ps, rc, a.exact, x, y, views, s.vw, [&](int h) { return refine_hypothesis(rc, rd, h, x, y); },
                         c5, evals);
  if (act) {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float4 n = refine_hypothesis(rc, rd, i, x, y);
      refine_take(rc, n, c5[i] / s.weight_norm, i, plane_now, depth_now, cost_now, x, y, &accepted);
    }
    strong_store(a, center, s.rng, plane_now, cost_now, cost_before, accepted);
  }
}

// ------------------------------------------------------------------------------------
// classify_refine_pixel (dpe_core.cuh) for a whole warp; act = this lane has a pixel
// ------------------------------------------------------------------------------------
template <class Env>
DPE_HDN void classify_refine_coop(const Env& env, const PatchStats& ps, const StageArgs& a, const int x, const int y,
                                     const bool act, unsigned& evals) {
  const RefConst& rc = env.rc();
  const int lane = coop_lane();
  const int W = a.W, H = a.H, N = rc.n_src;
  const int center = y * W + x;
  // ---- owner: what classify_refine_pixel does before its loop
  bool live = false, classify = false, refine = false;
  uint8_t new_state = DPE_UNKNOWN;
  float4 pl = make_float4(0.f, 0.f, 1.f, 0.f);
  float origin_depth = 1.f, base_line = 0.f, weight_normal = 0.f, disp = 0.f;
  uint32_t sel = 0u;
  ViewW vw;
  vw.clear();
  if (act) {
    const bool border = (x < 6 || y < 6 || x >= W - 6 || y >= H - 6);
    new_state = a.state[center];
    classify = true;
    if (border) { new_state = DPE_UNKNOWN; classify = false; }
    const float4 pw = a.planes[center];
    origin_depth = pw.w;
    pl = world_to_cam_normal(rc, pw);
    if (origin_depth == 0.f) {
      if (classify) new_state = DPE_UNKNOWN;
      a.state[center] = new_state;  // LocalRefine also returns (DPE.cu:2767)
    } else {
      sel = a.selected[center];
      vw = ViewW::unpack(a.view_w[center]);
      int valid = 0;
      for (int v = 0; v < N; ++v) {
        if ((sel >> v) & 1u) { base_line += rc.src[v].baseline; weight_normal += (float)vw.get(v); valid++; }
      }
      if (valid == 0) {
        if (classify) new_state = DPE_UNKNOWN;
        a.state[center] = new_state;
      } else {
        base_line /= valid;
        disp = rc.fx * base_line / origin_depth;
        refine = !(weight_normal == 0.f);
        live = true;
      }
    }
  }
  const uint32_t views = (live && N > 0) ? (sel & (0xffffffffu >> (32 - N))) : 0u;  // bits of the N <= 31 sources only

  // per-view sums of the owner's profile: acc[k + 30] for disparity step k = -30 .. 30, acc[61] for the stored depth
  // (cost_now of LocalRefine), lr[k + 5] LocalRefine's own accumulation of steps -5 .. 5 (see classify_refine_pixel)
  float acc[62], lr[11];
#pragma unroll 1
  for (int i = 0; i < 62; ++i) acc[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 11; ++i) lr[i] = 0.f;

  CoopCursor cur;
  cur.v = 0; cur.done = 0;
  CoopRound r;
  while (coop_next_round(views, vw, N, cur, r)) {
    // the pair's pixel
    const int px = coop_shfl(x, r.p), py = coop_shfl(y, r.p);
    const float4 ppl = shfl_plane(pl, r.p);
    const float p_origin = coop_shfl(origin_depth, r.p);
    const float p_base = coop_shfl(base_line, r.p);
    const float p_disp = coop_shfl(disp, r.p);
    const int p_flags = coop_shfl((classify ? 1 : 0) | (refine ? 2 : 0), r.p);
    const int p_klo = (p_flags & 1) ? -30 : -5, p_khi = (p_flags & 1) ? 30 : 5;
    const PatchStats pps = shfl_patch(ps, a.exact, r.p);
    Env penv = env;
    penv.tbl = env.tbl + (r.p - lane);
#pragma unroll 1
    for (int k = -30; k <= 31; ++k) {
      const bool extra = (k == 31);  // the stored depth itself
      float nc = -1.0f, g = 0.f;     // -1: not evaluated
      if (r.has && (extra ? (p_flags & 2) != 0 : (k >= p_klo && k <= p_khi))) {
        const float p_depth = extra ? p_origin : rc.fx * p_base / (p_disp + k);
        const bool in_range = extra || !(p_depth < rc.depth_min || p_depth > rc.depth_max);
        if (in_range) {
          float4 hp = ppl;
          hp.w = dist2origin(rc, px, py, p_depth, hp);
          const float3 m = plane_to_m(rc, hp);
          g = a.geom ? a.geom_factor * geom_cost(a, rc, rc.src[r.v], hp, px, py) : 0.f;
          nc = ncc_old(penv, pps, r.v, hp, m, px, py);
          evals += nc < 2.0f;
        }
      }
      for (int s = 0; s < r.n_slots; ++s) {
        int src = -1, w = 0;
#pragma unroll
        for (int t = 0; t < COOP_SLOTS; ++t)
          if (t == s) { src = r.src[t]; w = r.w[t]; }
        const float nc_s = coop_shfl(nc, src < 0 ? 0 : src);
        const float g_s = coop_shfl(g, src < 0 ? 0 : src);
        if (src >= 0 && nc_s >= 0.f) {
          float c = nc_s;
          if (a.geom) c += g_s;
          acc[k + 30] += c * w;
          if (k >= -5 && k <= 5) {
            const float wf = (float)w;
            lr[k + 5] = fmaf(nc_s, wf, lr[k + 5]);
            if (a.geom) lr[k + 5] = fmaf(g_s, wf, lr[k + 5]);
          }
        }
      }
    }
  }

  // ---- owner: profile, LocalRefine's minimum, decision
  if (live) {
    const int k_lo = classify ? -30 : -5, k_hi = classify ? 30 : 5;
    float lr_min = 2.0f, lr_best_depth = origin_depth;
#pragma unroll 1
    for (int k = k_lo; k <= k_hi; ++k) {
      const float p_depth = rc.fx * base_line / (disp + k);
      const bool in_range = !(p_depth < rc.depth_min || p_depth > rc.depth_max);
      float pc = 2.0f;
      if (in_range) {
        pc = acc[k + 30] / weight_normal;
        if (k >= -5 && k <= 5 && refine) {
          const float lr_pc = lr[k + 5] / weight_normal;
          if (lr_pc < lr_min) { lr_min = lr_pc; lr_best_depth = p_depth; }
        }
      }
      if (classify) acc[k + 30] = (2.0f > pc) ? pc : 2.0f;  // MIN(2.0f, pc); NaN -> 2.0 as in OpenCV's MIN
    }
    const float lr_now = refine ? acc[61] / weight_normal : 0.f;
    if (classify) new_state = classify_profile(acc, a.weak_peak_radius);
    a.state[center] = new_state;
    if (refine && (lr_now - lr_min > 0.1)) a.planes[center].w = lr_best_depth;  // double comparison, DPE.cu:2832
  }
}

}  // namespace dpe
