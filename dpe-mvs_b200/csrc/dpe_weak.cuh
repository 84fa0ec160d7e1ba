// dpe_weak.cuh — the DPE weak-texture / edge path: edge information, nearest strong pixel,
// anchor search + RANSAC, fit plane + adaptive radius, deformable NCC and the weak
// checkerboard propagation.  Restates (csrc/DPE-MVS/):
//   BresenhamLine, PointinTriangle          DPE.cu:135-250
//   ComputeBilateralNCCNew                  DPE.cu:557-690
//   PlaneHypothesisRefinementWeak           DPE.cu:1120-1212
//   CheckerboardPropagationWeak             DPE.cu:1668-1862
//   GenNeighbours / NeigbourUpdate          DPE.cu:2103-2481
//   GenEdgeInform                           DPE.cu:2483-2591
//   FindNearestStrongPoint                  DPE.cu:2855-2889
//   RANSACToGetFitPlane                     DPE.cu:2891-3124
// Anchors are stored 9 per pixel (the reference compacts them with a host-built prefix
// index, DPE.cpp:859-870; the values are the same).
#pragma once
#include "dpe_core.cuh"

namespace dpe {

DPE_HD float cross2(float ax, float ay, float bx, float by) { return ax * by - ay * bx; }

DPE_HD bool point_in_triangle(short2 A, short2 B, short2 C, int px, int py) {  // DPE.cu:135-156
  const float abx = B.x - A.x, aby = B.y - A.y, bcx = C.x - B.x, bcy = C.y - B.y, cax = A.x - C.x, cay = A.y - C.y;
  const float AB = sqrtf(abx * abx + aby * aby), BC = sqrtf(bcx * bcx + bcy * bcy), CA = sqrtf(cax * cax + cay * cay);
  if (AB <= 2 || BC <= 2 || CA <= 2) return false;
  if (!(AB + BC > CA && BC + CA > AB && AB + CA > BC)) return false;
  const float pax = A.x - px, pay = A.y - py, pbx = B.x - px, pby = B.y - py, pcx = C.x - px, pcy = C.y - py;
  const float t1 = cross2(pax, pay, pbx, pby), t2 = cross2(pbx, pby, pcx, pcy), t3 = cross2(pcx, pcy, pax, pay);
  return t1 * t2 >= 0 && t1 * t3 >= 0;
}

// coarse edge map lookup: the bit-packed copy when there is one (48 KB instead of 381 KB for a 756x504 map,
// so the Bresenham walks of the anchor search stay in L1)
DPE_HD bool edge_low_at(const StageArgs& a, const int x, const int y) {
  if (a.edge_low_bits) return (a.edge_low_bits[y * a.low_words + (x >> 5)] >> (x & 31)) & 1u;
  return a.edge_low[x + y * a.low_w] != 0;
}

DPE_HD bool bresenham_walk(const StageArgs& a, int lw, int lh, int x0, int y0, int x1, int y1, int max_step) {
  const int dx = x1 > x0 ? x1 - x0 : x0 - x1, sx = x0 < x1 ? 1 : -1;
  const int dy = y1 > y0 ? y1 - y0 : y0 - y1, sy = y0 < y1 ? 1 : -1;
  int err = (dx > dy ? dx : dy) / 2, step = 0;
  bool tagx = true, tagy = true;
  while (tagx || tagy) {
    if (x0 == x1) tagx = false;
    if (y0 == y1) tagy = false;
    const int e2 = err;
    if (e2 > -dx) { err -= dy; x0 += sx; }
    if (e2 < dy) { err += dx; y0 += sy; }
    // the reference indexes the coarse edge map unchecked; the walk can step one pixel past
    // the end point, so guard the read
    if (x0 >= 0 && y0 >= 0 && x0 < lw && y0 < lh && edge_low_at(a, x0, y0)) return true;
    if (++step >= max_step) break;
  }
  return false;
}

// true if the segment A-B crosses an edge of the coarsest edge map (DPE.cu:158-244)
DPE_HD bool bresenham_line(const StageArgs& a, int ax, int ay, int bx, int by) {
  if (a.edge[ax + ay * a.W] || a.edge[bx + by * a.W]) return false;
  const float scale_x = 1.0f * a.low_w / (float)a.W, scale_y = 1.0f * a.low_h / (float)a.H;
  const int lw = a.low_w, lh = a.low_h;
  const int max_step = (int)round(imax(lh, lw) / 60.0);
  const int bx0 = (int)fminf(roundf(bx * scale_x), (float)(lw - 1)), by0 = (int)fminf(roundf(by * scale_y), (float)(lh - 1));
  const int ax0 = (int)fminf(roundf(ax * scale_x), (float)(lw - 1)), ay0 = (int)fminf(roundf(ay * scale_y), (float)(lh - 1));
  if (bresenham_walk(a, lw, lh, bx0, by0, ax0, ay0, max_step)) return true;
  if (bresenham_walk(a, lw, lh, ax0, ay0, bx0, by0, max_step)) return true;
  return false;
}

// "does the segment between anchors i and j cross an edge" cache of the RANSAC loops: one tested bit and
// one result bit per pair (the reference keeps a byte per pair, DPE.cu:2289, 2957)
template <int N>
struct PairCache {
  unsigned long long tested[N], crossed[N];
  DPE_HD void clear() { for (int i = 0; i < N; ++i) { tested[i] = 0ull; crossed[i] = 0ull; } }
  DPE_HD bool crosses(const StageArgs& a, const short2* pts, const int i, const int j) {
    if (!((tested[i] >> j) & 1ull)) {
      const bool c = bresenham_line(a, pts[i].x, pts[i].y, pts[j].x, pts[j].y);
      tested[i] |= 1ull << j; tested[j] |= 1ull << i;
      if (c) { crossed[i] |= 1ull << j; crossed[j] |= 1ull << i; }
    }
    return (crossed[i] >> j) & 1ull;
  }
};

// ---- GenEdgeInform, DPE.cu:2483-2591 -------------------------------------------------------
DPE_HDN void edge_info_pixel(const StageArgs& a, const int x, const int y) {
  const int W = a.W, H = a.H, center = y * W + x;
  const int dirx[8] = {0, 0, -1, 1, -1, 1, -1, 1};
  const int diry[8] = {-1, 1, 0, 0, -1, 1, 1, -1};
  if (a.use_apd) {  // use_edge
    short2* en = a.edge_neigh + (size_t)center * 8;
    for (int i = 0; i < 8; ++i) {
      short2 r = make_short2(-1, -1);
      int nx = x + dirx[i], ny = y + diry[i];
      while (!(nx < 0 || nx >= W || ny < 0 || ny >= H)) {
        if (a.edge[nx + ny * W]) { r.x = (short)nx; r.y = (short)ny; break; }
        nx += dirx[i]; ny += diry[i];
      }
      en[i] = r;
    }
    int edge_pix = 0, tot_pix = 0, bound_pix = 0;
    for (int j = -5; j <= 5; ++j) {
      for (int i = -5; i <= 5; ++i) {
        const int nx = x + i, ny = y + j;
        if (nx < 0 || nx >= W || ny < 0 || ny >= H) continue;
        if (a.edge[ny * W + nx]) edge_pix++;
        if (a.label[ny * W + nx] == 0) bound_pix++;
        tot_pix++;
      }
    }
    float density = 1.0f * edge_pix / tot_pix;
    density = fmaxf(density, (float)(bound_pix / tot_pix));  // integer division
    a.complexity[center] = (float)(1.0f / (1.0f + exp(-25.0 * (density - 0.35))));  // double arithmetic, DPE.cu:2554
  }
}

// the use_label part of GenEdgeInform (DPE.cu:2560-2590): only WEAK pixels with a positive label take part
DPE_HDN void label_boundary_pixel(const StageArgs& a, const int x, const int y) {
  const int W = a.W, H = a.H, center = y * W + x;
  const int dirx[8] = {0, 0, -1, 1, -1, 1, -1, 1};
  const int diry[8] = {-1, 1, 0, 0, -1, 1, 1, -1};
  if (a.state[center] == DPE_WEAK) {  // use_label
    const int center_label = a.label[center];
    if (center_label > 0) {
      short2* lb = a.label_boundary + (size_t)center * 8;
      for (int i = 0; i < 8; ++i) {
        int nx = x + dirx[i], ny = y + diry[i], lx = -1, ly = -1;
        while (!(nx < 0 || nx >= W || ny < 0 || ny >= H)) {
          const int nl = a.label[nx + ny * W];
          if (nl == center_label) { lx = nx; ly = ny; }
          else if (nl == -1) break;
          nx += dirx[i]; ny += diry[i];
        }
        lb[i] = make_short2((short)lx, (short)ly);
      }
    }
  }
}

// ---- FindNearestStrongPoint, DPE.cu:2855-2889 ----------------------------------------------
DPE_HDN void nearest_strong_pixel(const StageArgs& a, const int x, const int y) {
  const int W = a.W, H = a.H, center = y * W + x;
  short2 r = make_short2(-1, -1);
  if (a.state[center] == DPE_WEAK) {
    bool found = false;
    for (int radius = 0; radius <= 100 && !found; ++radius) {
      for (int dx = -radius; dx <= radius && !found; ++dx) {
        const bool col_edge = (dx == -radius || dx == radius);
        const int step = (col_edge || radius == 0) ? 1 : 2 * radius;
        for (int dy = -radius; dy <= radius; dy += step) {
          const int nx = x + dx, ny = y + dy;
          if (nx < 0 || ny < 0 || nx >= W || ny >= H) continue;
          if (a.state[nx + ny * W] == DPE_STRONG) { r = make_short2((short)nx, (short)ny); found = true; break; }
        }
      }
    }
  }
  a.nearest_strong[center] = r;
}

// ---- GenNeighbours + NeigbourUpdate, DPE.cu:2103-2481 -------------------------------------
// planes[] holds (world normal, depth) here (the kernel runs before RandomInitialization).
DPE_HDN void gen_neighbours_pixel(const StageArgs& a, const int x, const int y) {
  const RefConst& rc = *a.rc;
  const int W = a.W, H = a.H, center = y * W + x;
  if (a.state[center] != DPE_WEAK) return;
  const int MAXP = 64, min_margin = 6;
  const float depth_diff = rc.depth_max - rc.depth_min;
  short2* neighbours = a.neighbours + (size_t)center * DPE_NEIGHBOUR_NUM;
  for (int i = 0; i < DPE_NEIGHBOUR_NUM; ++i) neighbours[i] = make_short2(-1, -1);
  neighbours[0] = make_short2((short)x, (short)y);
  short2 strong_points[MAXP];
  bool dir_valid[MAXP];
  for (int i = 0; i < MAXP; ++i) { strong_points[i] = make_short2(-1, -1); dir_valid[i] = false; }
  int strong_point_size = 0;
  Rng rng;
  rng.load(a.rng + center);

  const int rotate_time = a.rotate_time;
  const float angle = 45.0f / rotate_time;
  // in double like the reference (angle * M_PI is a double expression there, DPE.cu:2149-2152), so that
  // --use_fast_math does not turn these into the approximate single-precision intrinsics
  const double kPi = 3.14159265358979323846;
  const float cos_a = (float)cos(angle * kPi / 180.f), sin_a = (float)sin(angle * kPi / 180.f);
  const float threshold = (float)cos((angle / 2.0f) * kPi / 180.0f);
  const int shift_range = imax((int)(tan((angle / 2.0f) * kPi / 180.0f) * 20), 1);
  const float ransac_threshold = a.ransac_threshold * depth_diff;

  bool edge_limit = true;  // use_limit
  if (a.use_apd) {         // use_edge
    const float cv = a.complexity[center];
    const float rp = rng.uniform() - FLT_EPSILON;
    if (rp < cv) edge_limit = false;
    else a.complexity[center] = fmaxf(0.99f, cv);
  }

  int odi = -1;
  for (int odx = -1; odx <= 1; ++odx) {
    for (int ody = -1; ody <= 1; ++ody) {
      if (odx == 0 && ody == 0) continue;
      float ox = (float)odx, oy = (float)ody;
      { const float inv = fast_rsqrt(ox * ox + oy * oy); ox *= inv; oy *= inv; }
      odi++;
      for (int rot = 0; rot < rotate_time; ++rot) {
        const int dir_index = odi * 4 + rot;
        for (int radius = 2; radius <= 4096; radius = imin(radius * 2, radius + 25)) {
          const float tx = x + ox * radius, ty = y + oy * radius;
          if (tx < 0 || ty < 0 || tx >= W || ty >= H) break;
          for (int ri = 0; ri < 4; ++ri) {
            // unsigned arithmetic: the shift is always in [0, shift_range)
            const uint32_t s1 = (rng.next() % 2 == 0) ? 1u : 0xFFFFFFFFu;
            const int rxs = (int)((s1 * rng.next()) % (uint32_t)shift_range);
            const uint32_t s2 = (rng.next() % 2 == 0) ? 1u : 0xFFFFFFFFu;
            const int rys = (int)((s2 * rng.next()) % (uint32_t)shift_range);
            float dx = ox * 20 + rxs, dy = oy * 20 + rys;
            { const float inv = fast_rsqrt(dx * dx + dy * dy); dx *= inv; dy *= inv; }
            short2 np = make_short2((short)(x + dx * radius), (short)(y + dy * radius));
            if (np.x < min_margin || np.y < min_margin || np.x >= W - min_margin || np.y >= H - min_margin) continue;
            int npc = np.x + np.y * W;
            if (a.state[npc] != DPE_STRONG) {
              np = a.nearest_strong[npc];
              if (np.x == -1 || np.y == -1) continue;
              npc = np.x + np.y * W;
            }
            float tdx = (float)(np.x - x), tdy = (float)(np.y - y);
            { const float inv = fast_rsqrt(tdx * tdx + tdy * tdy); tdx *= inv; tdy *= inv; }
            const float ca = tdx * ox + tdy * oy;
            if (ca > threshold && (!edge_limit || !bresenham_line(a, x, y, np.x, np.y))) {
              strong_points[dir_index] = np;
              dir_valid[dir_index] = true;
              strong_point_size++;
              break;
            }
          }
          if (dir_valid[dir_index]) break;
        }
        const float rx = ox * cos_a - oy * sin_a, ry = ox * sin_a + oy * cos_a;
        const float inv = fast_rsqrt(rx * rx + ry * ry);
        ox = rx * inv; oy = ry * inv;
      }
    }
  }

  // label-guided extension anchors (DPE.cu:2224-2272)
  int extend_index = 31;
  const int center_label = a.label[center];
  if (center_label > 0) {
    const int dirx[8] = {0, 0, -1, 1, -1, 1, -1, 1};
    const int diry[8] = {-1, 1, 0, 0, -1, 1, 1, -1};
    const short2* lb = a.label_boundary + (size_t)center * 8;
    float bound_dist[8];
    int dir_step[8];
    for (int i = 0; i < 8; ++i) {
      const short2 bp = lb[i];
      float dist = 0.0f;
      if (bp.x != -1 && bp.y != -1) {
        dist = (float)sqrt((double)((x - bp.x) * (x - bp.x) + (y - bp.y) * (y - bp.y)));  // double in the reference (DPE.cu:2235)
        if (i >= 4) dist = (float)(dist / 1.4142135623730951);  // dist /= std::sqrt(2.0): a double division
      }
      bound_dist[i] = dist;
      if (i % 2 == 1) {
        // step = MIN(1, MAX(2 rt - 1, ...)) is always 1
        dir_step[i - 1] = 2 * rotate_time - 1;
        dir_step[i] = 1;
      }
    }
    for (int i = 0; i < 8; ++i) {
      const float dist = bound_dist[i];
      const int gap_num = dir_step[i] + 1;
      const int step_len = imax(1, (int)floor(1.0 * dist / gap_num));
      for (int step = 1; step <= dir_step[i]; ++step) {
        short2 np = make_short2((short)(x + step * step_len * dirx[i]), (short)(y + step * step_len * diry[i]));
        if (np.x < min_margin || np.y < min_margin || np.x >= W - min_margin || np.y >= H - min_margin) continue;
        int npc = np.x + np.y * W;
        if (a.state[npc] != DPE_STRONG) {
          np = a.nearest_strong[npc];
          if (np.x == -1 || np.y == -1) continue;
          npc = np.x + np.y * W;
        }
        if (a.label[npc] != 0 && a.label[npc] != center_label) continue;
        extend_index++;
        strong_points[extend_index] = np;
        dir_valid[extend_index] = true;
        strong_point_size++;
      }
    }
  }

  if (strong_point_size <= 3) { rng.store(a.rng + center); a.weak_reliable[center] = 0; a.state[center] = DPE_UNKNOWN; return; }

  float4 best_plane = make_float4(0.f, 0.f, 0.f, 0.f);
  bool has_valid_plane = false;
  // per-thread arrays live in local memory (~3 KB here): what can be recomputed from the anchor positions is
  // (the anchors' normals, their normalised image coordinates) — the stacks of the resident warps should stay in L2
  short2 spv[MAXP];
  float3 spv3d[MAXP];
  int valid_count = 0;
  float X[3];
  point3d(rc, x, y, a.planes[center].w, X);
  const float center_z = X[2];
  for (int i = 0; i < MAXP; ++i) {
    spv[i] = make_short2(-1, -1);
    if (dir_valid[i]) {
      const short2 sp = strong_points[i];
      const int spc = sp.x + sp.y * W;
      spv[valid_count] = sp;
      const float4 spl = a.planes[spc];
      point3d(rc, sp.x, sp.y, spl.w, X);
      spv3d[valid_count] = make_float3(X[0], X[1], X[2]);
      valid_count++;
    }
  }

  {
    int iteration = 50, max_iter = 200, max_count = 3;
    float min_cost = FLT_MAX;
    float residuals[MAXP];
    for (int i = 0; i < MAXP; ++i) residuals[i] = 0.f;
    float temp_thr = ransac_threshold;
    PairCache<MAXP> edge_test;
    edge_test.clear();
    const float fx_c = fast_div(x - rc.cx, rc.fx), fy_c = fast_div(y - rc.cy, rc.fy);
    auto anchor_normal = [&](const int i) {
      const float4 n4 = world_to_cam_normal(rc, a.planes[spv[i].x + spv[i].y * W]);
      return make_float3(n4.x, n4.y, n4.z);
    };
    bool has_consist_normal_plane = false;
    bool must_in_triangle = !(center_label > 0 && edge_limit);
    while (iteration > 0 && max_iter > 0) {
      max_iter--;
      const int ai = (int)(rng.next() % (uint32_t)valid_count);
      const int bi = (int)(rng.next() % (uint32_t)valid_count);
      const int ci = (int)(rng.next() % (uint32_t)valid_count);
      if (ai == bi || bi == ci || ai == ci) continue;
      if (must_in_triangle && !point_in_triangle(spv[ai], spv[bi], spv[ci], x, y)) continue;
      if (edge_limit) {
        const bool c0 = edge_test.crosses(a, spv, ai, bi), c1 = edge_test.crosses(a, spv, bi, ci), c2 = edge_test.crosses(a, spv, ci, ai);
        if (c0 || c1 || c2) continue;
      }
      bool normal_consistency = false;
      if (a.geom && edge_limit) {
        const float3 AN = anchor_normal(ai), BN = anchor_normal(bi), CN = anchor_normal(ci);
        normal_consistency = true;
        // the threshold is a double literal in the reference (DPE.cu:2347)
        if (AN.x * BN.x + AN.y * BN.y + AN.z * BN.z < 0.8660254 || AN.x * CN.x + AN.y * CN.y + AN.z * CN.z < 0.8660254 ||
            BN.x * CN.x + BN.y * CN.y + BN.z * CN.z < 0.8660254)
          normal_consistency = false;
        if (has_consist_normal_plane && !normal_consistency) continue;
      }
      iteration--;
      const float3 A = spv3d[ai], B = spv3d[bi], C = spv3d[ci];
      const float3 AC = make_float3(A.x - C.x, A.y - C.y, A.z - C.z), BC = make_float3(B.x - C.x, B.y - C.y, B.z - C.z);
      float4 cv;
      cv.x = AC.y * BC.z - BC.y * AC.z;
      cv.y = -(AC.x * BC.z - BC.x * AC.z);
      cv.z = AC.x * BC.y - BC.x * AC.y;
      cv.w = 0.f;
      if ((cv.x == 0 && cv.y == 0 && cv.z == 0) || cv.x != cv.x || cv.y != cv.y || cv.z != cv.z) continue;
      normalize3(cv);
      cv.w = -(cv.x * A.x + cv.y * A.y + cv.z * A.z);
      int temp_count = 0;
      for (int si = 0; si < valid_count; ++si) {
        const float fxs = fast_div(spv[si].x - rc.cx, rc.fx), fys = fast_div(spv[si].y - rc.cy, rc.fy);  // DPE.cu:2383-2385
        const float fit_depth = fast_div(-cv.w, cv.x * fxs + cv.y * fys + cv.z);
        const float distance = fabsf(fit_depth - spv3d[si].z);
        residuals[si] = distance;
        if (distance < temp_thr) temp_count++;
      }
      if (temp_count < 6) continue;
      if (temp_count > max_count) {
        if (!must_in_triangle && point_in_triangle(spv[ai], spv[bi], spv[ci], x, y)) must_in_triangle = true;
        if (!has_consist_normal_plane && normal_consistency) has_consist_normal_plane = true;
        const float fit_depth = fast_div(-cv.w, cv.x * fx_c + cv.y * fy_c + cv.z);
        best_plane = cv;
        max_count = temp_count;
        min_cost = fabsf(fit_depth - center_z);
        has_valid_plane = true;
        if (temp_thr > 0.05) {  // high_res_img (main.h:97); double literals, DPE.cu:2403-2406
          sort_small(residuals, valid_count);
          if (temp_thr < residuals[DPE_NEIGHBOUR_NUM]) continue;
          temp_thr = (float)(residuals[DPE_NEIGHBOUR_NUM] - 1e-6);
          temp_count = 0;
          for (int i = 0; i < valid_count; ++i) {
            if (residuals[i] < temp_thr) temp_count++;
            else break;
          }
          max_count = temp_count;
        }
      } else if (temp_count == max_count) {
        if (!must_in_triangle && point_in_triangle(spv[ai], spv[bi], spv[ci], x, y)) must_in_triangle = true;
        const float fit_depth = fast_div(-cv.w, cv.x * fx_c + cv.y * fy_c + cv.z);
        const float cd = fabsf(fit_depth - center_z);
        if (cd < min_cost) { best_plane = cv; max_count = temp_count; min_cost = cd; }
      }
    }
  }

  rng.store(a.rng + center);
  if (!has_valid_plane) { a.weak_reliable[center] = 0; a.state[center] = DPE_UNKNOWN; return; }

  float weight[MAXP];
  for (int i = 0; i < valid_count; ++i) {
    const float fxx = fast_div(spv[i].x - rc.cx, rc.fx), fyy = fast_div(spv[i].y - rc.cy, rc.fy);
    const float fit_depth = fast_div(-best_plane.w, best_plane.x * fxx + best_plane.y * fyy + best_plane.z);
    const float distance = fabsf(fit_depth - spv3d[i].z);
    if (distance >= ransac_threshold) { spv[i] = make_short2(-1, -1); weight[i] = FLT_MAX; continue; }
    weight[i] = distance;
  }
  // sort_small_weighted, DPE.cu:16-29
  for (int i = 1; i < valid_count; i++) {
    const short2 tp = spv[i];
    const float tw = weight[i];
    int j = i;
    for (; j >= 1 && tw < weight[j - 1]; j--) { spv[j] = spv[j - 1]; weight[j] = weight[j - 1]; }
    spv[j] = tp; weight[j] = tw;
  }
  for (int i = 1; i < DPE_NEIGHBOUR_NUM; ++i) neighbours[i] = spv[i - 1];
  a.weak_reliable[center] = 1;
}

// ---- RANSACToGetFitPlane, DPE.cu:2891-3124 (planes[] in camera coordinates) ---------------
DPE_HDN void fit_plane_pixel(const StageArgs& a, const int x, const int y) {
  const RefConst& rc = *a.rc;
  const int W = a.W, center = y * W + x;
  if (a.state[center] != DPE_WEAK) { a.fit_planes[center] = a.planes[center]; return; }
  Rng rng;
  rng.load(a.rng + center);
  bool edge_limit = true;
  if (a.use_apd) {
    const float cv = a.complexity[center];
    const float rp = rng.uniform() - FLT_EPSILON;
    if (rp < cv) edge_limit = false;
  }
  const int NB = DPE_NEIGHBOUR_NUM - 1;
  short2 sp[NB];
  float3 sp3[NB], spn[NB];
  int sc = 0;
  float X[3];
  const short2* nbrs = a.neighbours + (size_t)center * DPE_NEIGHBOUR_NUM;
  for (int i = 1; i < DPE_NEIGHBOUR_NUM; ++i) {
    const short2 tp = nbrs[i];
    if (tp.x == -1 || tp.y == -1) continue;
    sp[sc] = tp;
    const float4 pl = a.planes[tp.x + tp.y * W];
    const float d = depth_from_plane(rc, pl, tp.x, tp.y);
    point3d(rc, tp.x, tp.y, d, X);
    sp3[sc] = make_float3(X[0], X[1], X[2]);
    spn[sc] = make_float3(pl.x, pl.y, pl.z);
    sc++;
  }
  if (sc < 3) { rng.store(a.rng + center); a.fit_planes[center] = a.planes[center]; return; }

  int iteration = 50, ua = -1, ub = -1, uc = -1;
  float min_cost = FLT_MAX;
  float4 best_plane = make_float4(0.f, 0.f, 0.f, 0.f);
  bool has_best = false, has_strong_plane = false;
  const int center_label = a.label[center];
  bool must_in_triangle = !(center_label > 0 && edge_limit);
  PairCache<NB> edge_test;
  edge_test.clear();
  float fxs[NB], fys[NB];
  for (int si = 0; si < sc; ++si) {
    fxs[si] = fast_div(sp[si].x - rc.cx, rc.fx);
    fys[si] = fast_div(sp[si].y - rc.cy, rc.fy);
  }
  while (iteration--) {
    const int ai = (int)(rng.next() % (uint32_t)sc), bi = (int)(rng.next() % (uint32_t)sc), ci = (int)(rng.next() % (uint32_t)sc);
    if (ai == bi || bi == ci || ai == ci) continue;
    bool is_strong_plane = false;
    if (a.geom && edge_limit) {
      const float3 AN = spn[ai], BN = spn[bi], CN = spn[ci];
      is_strong_plane = true;
      if (AN.x * BN.x + AN.y * BN.y + AN.z * BN.z < 0.8660254 || AN.x * CN.x + AN.y * CN.y + AN.z * CN.z < 0.8660254 ||
          BN.x * CN.x + BN.y * CN.y + BN.z * CN.z < 0.8660254)
        is_strong_plane = false;
      if (has_strong_plane && !is_strong_plane) continue;
    }
    if (must_in_triangle && !point_in_triangle(sp[ai], sp[bi], sp[ci], x, y)) continue;
    if (edge_limit) {
      const bool c0 = edge_test.crosses(a, sp, ai, bi), c1 = edge_test.crosses(a, sp, bi, ci), c2 = edge_test.crosses(a, sp, ci, ai);
      if (c0 || c1 || c2) continue;
    }
    const float3 A = sp3[ai], B = sp3[bi], C = sp3[ci];
    const float3 AC = make_float3(A.x - C.x, A.y - C.y, A.z - C.z), BC = make_float3(B.x - C.x, B.y - C.y, B.z - C.z);
    float4 cv;
    cv.x = AC.y * BC.z - BC.y * AC.z;
    cv.y = -(AC.x * BC.z - BC.x * AC.z);
    cv.z = AC.x * BC.y - BC.x * AC.y;
    cv.w = 0.f;
    if ((cv.x == 0 && cv.y == 0 && cv.z == 0) || cv.x != cv.x || cv.y != cv.y || cv.z != cv.z) continue;
    normalize3(cv);
    cv.w = -(cv.x * A.x + cv.y * A.y + cv.z * A.z);
    if (!has_strong_plane && is_strong_plane) has_strong_plane = true;
    float temp_cost = 0.f;
    for (int si = 0; si < sc; ++si) {
      if (si == ai || si == bi || si == ci) continue;
      const float fit_depth = fast_div(-cv.w, cv.x * fxs[si] + cv.y * fys[si] + cv.z);
      temp_cost += fabsf(fit_depth - sp3[si].z);
    }
    if (temp_cost < min_cost) {
      if (!must_in_triangle && point_in_triangle(sp[ai], sp[bi], sp[ci], x, y)) must_in_triangle = true;
      min_cost = temp_cost; best_plane = cv; has_best = true; ua = ai; ub = bi; uc = ci;
    }
  }

  rng.store(a.rng + center);
  if (!has_best) {
    a.fit_planes[center] = make_float4(0.f, 0.f, 0.f, 0.f);
    a.radius[center] = 5;
    return;
  }
  const float depth = depth_from_plane(rc, a.planes[center], x, y);
  const float4 vd = view_direction(rc, x, y, depth);
  if (best_plane.x * vd.x + best_plane.y * vd.y + best_plane.z * vd.z > 0) {
    best_plane.x = -best_plane.x; best_plane.y = -best_plane.y; best_plane.z = -best_plane.z; best_plane.w = -best_plane.w;
  }
  a.fit_planes[center] = best_plane;
  // adaptive patch radius (use_radius), DPE.cu:3060-3114
  if (!must_in_triangle) { a.radius[center] = 5; return; }
  const short2 A = sp[ua], B = sp[ub], C = sp[uc];
  const float la = sqrtf((float)((A.x - B.x) * (A.x - B.x) + (A.y - B.y) * (A.y - B.y)));
  const float lb = sqrtf((float)((B.x - C.x) * (B.x - C.x) + (B.y - C.y) * (B.y - C.y)));
  const float lc = sqrtf((float)((C.x - A.x) * (C.x - A.x) + (C.y - A.y) * (C.y - A.y)));
  const float p = (float)((la + lb + lc) / 2.0);
  const float S = sqrtf(p * (p - la) * (p - lb) * (p - lc));
  int radius = (int)floor(sqrtf(S) / 2.0);
  const float Ad = sqrtf((float)((A.x - x) * (A.x - x) + (A.y - y) * (A.y - y)));
  const float Bd = sqrtf((float)((B.x - x) * (B.x - x) + (B.y - y) * (B.y - y)));
  const float Cd = sqrtf((float)((C.x - x) * (C.x - x) + (C.y - y) * (C.y - y)));
  const float min_dis = fminf(fminf(Ad, Bd), Cd);
  if (2.5 * min_dis < radius) radius = (int)min_dis;
  if (edge_limit) {
    if (a.use_apd) {
      float med = FLT_MAX;
      const short2* en = a.edge_neigh + (size_t)center * 8;
      for (int d = 0; d < 8; ++d) {
        const short2 e = en[d];
        if (e.x == -1 || e.y == -1) continue;
        med = fminf(med, sqrtf((float)((e.x - x) * (e.x - x) + (e.y - y) * (e.y - y))));
      }
      if (med < radius) radius = (int)med;
    }
    if (center_label > 0) {
      float mbd = FLT_MAX;
      const short2* lbp = a.label_boundary + (size_t)center * 8;
      for (int d = 0; d < 8; ++d) {
        const short2 b = lbp[d];
        if (b.x == -1 || b.y == -1) continue;
        mbd = fminf(mbd, (float)sqrt((double)((x - b.x) * (x - b.x) + (y - b.y) * (y - b.y))));  // double in the reference (DPE.cu:3097)
      }
      if (mbd < radius) radius = (int)mbd;
    }
  }
  // NaN area (degenerate triangle) gives INT_MIN in the reference's float->int cast and an
  // endless loop guard is needed here; clamp to 0 first
  if (radius < 0) radius = 0;
  while ((radius << 1) % 5 != 0) radius--;
  if (!edge_limit) a.radius[center] = radius > 5 ? 0 : 5;
  else a.radius[center] = radius > 5 ? radius : 5;
}

// ---- deformable NCC, ComputeBilateralNCCNew DPE.cu:557-690 ---------------------------------
// The reference recomputes, for every (hypothesis, view) it scores at a WEAK pixel, the reference
// intensities, bilateral weights and reference moments of all nine patches (centre patch with
// the adaptive radius + up to 8 anchor patches of 3 x 3 taps; weights relative to the centre
// pixel's intensity, DPE.cu:619-663).  None of that depends on the hypothesis or the view, so it
// is built once per pixel and sweep into a WeakTab and every evaluation only does the source
// fetches and the source-side sums.  The adaptive radius is 0 or a multiple of 5 with
// inc = max(2, 2r/5): the centre patch has 1 or 6 taps per axis.
struct WeakTab {
  float2 ww[108];     // (w, w * (r - r0)) per tap; patch k starts at tab_offset(k)
  short2 xy[108];     // reference pixel of the tap
  float inv_sw[9], mean_r[9], var_r[9];
  short2 anchor[9];   // anchor[0] = the pixel itself; (-1,-1) = absent
  uint32_t asel[9];   // selected-view bits of the anchors (constant during a weak sweep: anchors are STRONG)
  uint8_t ntap[9];    // taps per axis: 0 = absent, 1, 3 or 6
};
DPE_HD int tab_offset(const int k) { return k == 0 ? 0 : 36 + (k - 1) * 9; }

// reference side of patch k (tap order: x offset outer, y offset inner, as DPE.cu:619-621)
template <class RefFetch>
DPE_HD void build_weak_patch(const RefFetch& ref, const float r0, const float c0, WeakTab& T, const int k, const int first,
                             const int inc) {
  const int n = T.ntap[k];
  if (n == 0) return;
  const short2 np = T.anchor[k];
  const int off = tab_offset(k);
  float sw = 0.f, sr = 0.f, srr = 0.f;
  for (int ti = 0; ti < n; ++ti) {
    float sw_c = 0.f, sr_c = 0.f, srr_c = 0.f;  // one partial sum per tap column, as the reference
    for (int tj = 0; tj < n; ++tj) {
      const int i = first + ti * inc, j = first + tj * inc;
      const int rx = np.x + i, ry = np.y + j;
      const float r = ref(rx, ry);
      const float w = bilateral_weight(i, j, r, r0);
      const float rp = r - c0;
      const float wr = mul_rn(w, rp);
      T.ww[off + ti * n + tj] = make_float2(w, wr);
      T.xy[off + ti * n + tj] = make_short2((short)rx, (short)ry);
      sw_c += w; sr_c += wr; srr_c = fmaf(wr, rp, srr_c);
    }
    sw += sw_c; sr += sr_c; srr += srr_c;
  }
  const float inv = fast_rcp(sw);
  const float mr = inv * sr;
  T.inv_sw[k] = inv; T.mean_r[k] = mr; T.var_r[k] = fmaf(inv, srr, -mul_rn(mr, mr));
}

// anchors, their selected views, tap counts: everything of the WeakTab but the patches
DPE_HD void init_weak_tab_entry(const StageArgs& a, const int center, WeakTab& T, const int k, int& first0, int& inc0) {
  const short2 np = a.neighbours[(size_t)center * DPE_NEIGHBOUR_NUM + k];
  T.anchor[k] = np;
  const bool present = !(np.x == -1 || np.y == -1);
  T.asel[k] = present ? a.selected[np.x + np.y * a.W] : 0u;
  if (k == 0) {
    const int radius = a.radius[center];
    const int inc = imax(2, (int)(2.0 * radius / 5.0));
    T.ntap[0] = present ? (radius < inc ? 1 : 6) : 0;
    first0 = -radius; inc0 = inc;
  } else {
    T.ntap[k] = present ? 3 : 0;
  }
}

// source side of one patch: NTAP x NTAP independent fetches
template <int NTAP, bool EXACT, class Env>
DPE_HD float patch_cost_tab(const Env& env, const float c0, const WeakTab& T, const int k, const SrcConst& sc, const float* h) {
  const int off = tab_offset(k);
  float ss = 0.f, sss = 0.f, srs = 0.f;
  // one column of taps (NTAP independent fetches) per iteration: the code stays small enough for the
  // instruction cache when every warp of an SM is at a different place of the weak sweep
#pragma unroll 1
  for (int tr = 0; tr < NTAP; ++tr) {
    float ss_c = 0.f, sss_c = 0.f, srs_c = 0.f;
#pragma unroll
    for (int tc = 0; tc < NTAP; ++tc) {
      const int t = tr * NTAP + tc;
      const short2 q = T.xy[off + t];
      float u, v;
      if (EXACT) {  // the reference's association (ncc_old_exact, dpe_core.cuh)
        const float qx = (float)q.x, qy = (float)q.y;
        const float iz = fast_rcp(add_rn(h[8], fmaf(h[7], qy, mul_rn(h[6], qx))));
        u = fmaf(add_rn(h[2], fmaf(h[1], qy, mul_rn(h[0], qx))), iz, 0.5f);
        v = fmaf(add_rn(h[5], fmaf(h[4], qy, mul_rn(h[3], qx))), iz, 0.5f);
      } else {
        const float iz = fast_rcp(h[6] * q.x + h[7] * q.y + h[8]);
        u = (h[0] * q.x + h[1] * q.y + h[2]) * iz + 0.5f;
        v = (h[3] * q.x + h[4] * q.y + h[5]) * iz + 0.5f;
      }
      const float s = env.tex(sc, u, v) - c0;
      const float2 ww = T.ww[off + t];
      const float ws = mul_rn(ww.x, s);
      ss_c += ws; sss_c = fmaf(ws, s, sss_c); srs_c = fmaf(ww.y, s, srs_c);
    }
    ss += ss_c; sss += sss_c; srs += srs_c;
  }
  return ncc_finish(T.inv_sw[k], T.mean_r[k], T.var_r[k], ss, sss, srs);
}

// `taps` accumulates evaluated source taps
// EXACT is a template flag: a run-time test per tap would put both variants into the tap loops, and the weak
// sweep is sensitive to its code size (instruction cache, see patch_cost_tab)
template <bool EXACT, class Env>
__noinline__ DPE_HDN float ncc_new_t(const Env& env, const float c0, const WeakTab& T, const SrcConst& sc, const int v, const float3 m,
                                     const int x, const int y, const int W, const int H, int& taps,
                                     const RefConst* rc_exact, const float4 pl) {
  float h[9];
  if (EXACT) {
    homography_ref(*rc_exact, sc, pl, h);  // StageArgs::exact: the reference's operation order (dpe_core.cuh)
  } else {
    h[0] = sc.A[0] - sc.b[0] * m.x; h[1] = sc.A[1] - sc.b[0] * m.y; h[2] = sc.A[2] - sc.b[0] * m.z;
    h[3] = sc.A[3] - sc.b[1] * m.x; h[4] = sc.A[4] - sc.b[1] * m.y; h[5] = sc.A[5] - sc.b[1] * m.z;
    h[6] = sc.A[6] - sc.b[2] * m.x; h[7] = sc.A[7] - sc.b[2] * m.y; h[8] = sc.A[8] - sc.b[2] * m.z;
  }
  // H (x, y, 1) dehomogenised: ComputeCorrespondingPoint (DPE.cu:515-522)
  auto corr = [&](const int qx_, const int qy_, float& ox, float& oy) {
    if (EXACT) {
      const float qx = (float)qx_, qy = (float)qy_;
      const float iz = fast_rcp(add_rn(h[8], fmaf(h[7], qy, mul_rn(h[6], qx))));
      ox = mul_rn(add_rn(h[2], fmaf(h[1], qy, mul_rn(h[0], qx))), iz);
      oy = mul_rn(add_rn(h[5], fmaf(h[4], qy, mul_rn(h[3], qx))), iz);
    } else {
      const float Z = h[6] * qx_ + h[7] * qy_ + h[8];
      ox = fast_div(h[0] * qx_ + h[1] * qy_ + h[2], Z); oy = fast_div(h[3] * qx_ + h[4] * qy_ + h[5], Z);
    }
  };
  {
    float px, py;
    corr(x, y, px, py);
    if (px >= sc.width || px < 0.0f || py >= sc.height || py < 0.0f) return 2.0f;
  }
  float center_cost = 0.f, strong_cost = 0.f;
  int strong_count = 0;
#pragma unroll 1
  for (int k = 0; k < DPE_NEIGHBOUR_NUM; ++k) {
    const int n = T.ntap[k];
    if (n == 0) continue;
    const short2 np = T.anchor[k];
    {
      float qx, qy;
      corr(np.x, np.y, qx, qy);
      if (qx < 0 || qy < 0 || qx >= W || qy >= H) {  // sic: reference-image size (DPE.cu:596)
        if (k != 0) {
          if ((T.asel[k] >> v) & 1u) { strong_cost += 2.0f; strong_count++; }
          continue;
        }
        return 2.0f;
      }
    }
    if (k == 0) {
      center_cost = (n == 1) ? patch_cost_tab<1, EXACT>(env, c0, T, 0, sc, h) : patch_cost_tab<6, EXACT>(env, c0, T, 0, sc, h);
      taps += n * n;
    } else {
      strong_cost += patch_cost_tab<3, EXACT>(env, c0, T, k, sc, h);
      strong_count++;
      taps += 9;
    }
  }
  if (strong_count == 0) return center_cost;
  strong_cost /= strong_count;
  strong_cost = fminf(strong_cost, 2.0f);
  return (float)(0.25 * center_cost + 0.75 * strong_cost);
}

template <class Env>
DPE_HD float ncc_new(const Env& env, const float c0, const WeakTab& T, const SrcConst& sc, const int v, const float3 m,
                     const int x, const int y, const int W, const int H, int& taps,
                     const RefConst* rc_exact = nullptr, const float4 pl = float4{0.f, 0.f, 0.f, 1.f}) {
  return rc_exact ? ncc_new_t<true>(env, c0, T, sc, v, m, x, y, W, H, taps, rc_exact, pl)
                  : ncc_new_t<false>(env, c0, T, sc, v, m, x, y, W, H, taps, nullptr, pl);
}

// weighted (photometric + geometric) cost of one hypothesis over the sampled views
template <class Env>
DPE_HD float weighted_cost_weak(const Env& env, const PatchStats& ps, const WeakTab& T, const StageArgs& a, const float4 pl,
                                const int x, const int y, const ViewW& vw, const float weight_norm, int& taps) {
  const RefConst& rc = *a.rc;
  const float3 m = plane_to_m(rc, pl);
  float c = 0.f;
  for (int v = 0; v < rc.n_src; ++v) {
    const int w = vw.get(v);
    if (w > 0) {
      float cv = ncc_new(env, ps.c0, T, rc.src[v], v, m, x, y, a.W, a.H, taps, a.exact ? &rc : nullptr, pl);
      if (a.geom) cv += a.geom_factor * geom_cost(a, rc, rc.src[v], pl, x, y);
      c += w * cv;
    }
  }
  return c / weight_norm;
}

// ---- CheckerboardPropagationWeak, DPE.cu:1668-1862 -----------------------------------------
template <class Env>
DPE_HDN void weak_update_pixel(const Env& env, const PatchStats& ps, const StageArgs& a, const int x, const int y,
                               float* cost_arr, unsigned& evals) {
  const RefConst& rc = *a.rc;
  const int W = a.W, N = rc.n_src, center = y * W + x;
  const int iter = a.iter;
  int taps = 0;
  WeakTab T;
  {
    int first0 = 0, inc0 = 2;
    for (int k = 0; k < DPE_NEIGHBOUR_NUM; ++k) init_weak_tab_entry(a, center, T, k, first0, inc0);
    auto ref = [&](int rx, int ry) { return env.ref(rx, ry); };
    for (int k = 0; k < DPE_NEIGHBOUR_NUM; ++k) build_weak_patch(ref, ps.r0, ps.c0, T, k, k == 0 ? first0 : -5, k == 0 ? inc0 : 5);
  }
  for (int j = 0; j < 8; ++j)
    for (int v = 0; v < N; ++v) cost_arr[j * DPE_MAX_IMAGES + v] = 0.f;
  cost_arr[0] = 2.0f;
  bool flag[8];
  int positions[8];
  const short2* nbrs = a.neighbours + (size_t)center * DPE_NEIGHBOUR_NUM;
  float priors[DPE_MAX_IMAGES];
  for (int v = 0; v < N; ++v) priors[v] = 0.f;
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    flag[i] = false; positions[i] = 0;
    const short2 np = nbrs[i + 1];
    if (np.x == -1 || np.y == -1) continue;
    const int npc = np.x + np.y * W;
    const uint32_t sv = a.selected[npc];
    for (int v = 0; v < N; ++v) priors[v] += ((sv >> v) & 1u) ? 0.9f : 0.1f;
    if (a.state[npc] != DPE_STRONG) continue;
    positions[i] = npc; flag[i] = true;
    const float4 cpl = a.planes[npc];
    const float3 m = plane_to_m(rc, cpl);
    for (int v = 0; v < N; ++v)
      cost_arr[i * DPE_MAX_IMAGES + v] = ncc_new(env, ps.c0, T, rc.src[v], v, m, x, y, a.W, a.H, taps, a.exact ? &rc : nullptr, cpl);
  }
  Rng rng;
  rng.load(a.rng + center);
  ViewW vw;
  float weight_norm;
  uint32_t sel_bits;
  sample_views(cost_arr, priors, N, iter, rng, vw, weight_norm, sel_bits);
  a.view_w[center] = vw.pack();

  float final_costs[8];
#pragma unroll 1
  for (int j = 0; j < 8; ++j) {
    float f = 0.f;
    float4 cand = make_float4(0.f, 0.f, 1.f, 1.f);
    if (flag[j]) cand = a.planes[positions[j]];
    for (int v = 0; v < N; ++v) {
      const int w = vw.get(v);
      if (w > 0) {
        float c = cost_arr[j * DPE_MAX_IMAGES + v];
        if (a.geom) c += a.geom_factor * (flag[j] ? geom_cost(a, rc, rc.src[v], cand, x, y) : 3.0f);
        f += w * c;
      }
    }
    final_costs[j] = f / weight_norm;
  }
  int min_idx = 0;
  {
    float mc = final_costs[0];
    for (int j = 1; j < 8; ++j)
      if (final_costs[j] <= mc) { mc = final_costs[j]; min_idx = j; }
  }
  float4 plane_now = a.planes[center];
  float cost_now = weighted_cost_weak(env, ps, T, a, plane_now, x, y, vw, weight_norm, taps);
  const float cost_before = cost_now;
  float depth_now = depth_from_plane(rc, plane_now, x, y);
  if (flag[min_idx]) {
    const float4 cand = a.planes[positions[min_idx]];
    const float db = depth_from_plane(rc, cand, x, y);
    if (db >= rc.depth_min && db <= rc.depth_max && final_costs[min_idx] < cost_now) {
      depth_now = db; plane_now = cand; cost_now = final_costs[min_idx];
      a.selected[center] = sel_bits;
    }
  }
  // PlaneHypothesisRefinementWeak, DPE.cu:1120-1212
  {
    const float dmin = rc.depth_min, dmax = rc.depth_max;
    const float4 fit = a.fit_planes[center];
    const bool has_fit = !(fit.x == 0 && fit.y == 0 && fit.z == 0);
    if (has_fit) {  // without a fit plane the reference returns before the random refinement
      {
        const float c = weighted_cost_weak(env, ps, T, a, fit, x, y, vw, weight_norm, taps);
        const float db = depth_from_plane(rc, fit, x, y);
        if (db >= dmin && db <= dmax && c < cost_now) { depth_now = db; plane_now = fit; cost_now = c; }
      }
      const float depth_rand = rng.uniform() * (dmax - dmin) + dmin;
      const float4 n_rand = random_normal(rc, x, y, rng, depth_now);
      const float lo = (1 - 0.02f) * depth_now, hi = (1 + 0.02f) * depth_now;
      const float depth_pert = rng.uniform() * (hi - lo) + lo;
      const float4 n_pert = perturbed_normal(rc, x, y, plane_now, rng, (float)(0.02f * 3.14159265358979323846));
      const float4 plane_in = plane_now;
      const float depth_in = depth_now;
#pragma unroll 1
      for (int i = 0; i < 5; ++i) {
        const float d = (i == 0 || i == 2) ? depth_rand : (i == 4 ? depth_pert : depth_in);
        float4 n = (i == 1 || i == 2) ? n_rand : (i == 3 ? n_pert : plane_in);
        n.w = dist2origin(rc, x, y, d, n);
        const float c = weighted_cost_weak(env, ps, T, a, n, x, y, vw, weight_norm, taps);
        const float db = depth_from_plane(rc, n, x, y);
        if (db >= dmin && db <= dmax && c < cost_now) { depth_now = db; plane_now = n; cost_now = c; }
      }
    }
  }
  rng.store(a.rng + center);
  float4 final_plane = a.planes[center];
  if (a.run_state == DPE_REFINE_INIT) {
    if (cost_now < cost_before - 0.1) { final_plane = plane_now; a.planes[center] = plane_now; }  // double, DPE.cu:1835
  } else {
    final_plane = plane_now;
    a.planes[center] = plane_now;
  }
  // costs[] is re-scored with the plain 6x6 NCC (DPE.cu:1845-1861)
  {
    const float3 m = plane_to_m(rc, final_plane);
    float c = 0.f;
    for (int v = 0; v < N; ++v) {
      const int w = vw.get(v);
      if (w > 0) {
        const float cv = ncc_old(env, ps, v, final_plane, m, x, y);
        c += w * cv;
        if (cv < 2.0f) taps += 36;
      }
    }
    a.costs[center] = c / weight_norm;
  }
  evals += (unsigned)((taps + 18) / 36);
}

}  // namespace dpe
