// TEST INFRASTRUCTURE — CPU restatement of the reference's depth-map fusion (RunFusion, DPE.cpp:1220-1370):
// sequential over views and, inside a view, over pixels in raster order, because a source pixel that was
// fused into a point is masked out for everything that comes later (DPE.cpp:1299, 1353).  The product
// fuses on the device (dpe-mvs_b200/csrc/dpe_fusion.cu); tests/test_gpu_parity.py compares the two clouds.
// Built by oracle/Makefile into _ref/libfusion_oracle.so.
#include <stdint.h>
#include <string>
#include <vector>

#define DPE_WEAK 0

namespace dpe_host {
struct CamFile { float K[9], R[9], t[3]; float depth_min, depth_max; };
struct FusionInput {
  int width = 0, height = 0, n_views = 0;
  const std::vector<std::vector<float>>* depth = nullptr;
  const std::vector<std::vector<float>>* normal = nullptr;
  const std::vector<std::vector<uint8_t>>* state = nullptr;
  const std::vector<std::vector<uint8_t>>* bgr = nullptr;
  const std::vector<CamFile>* cams = nullptr;
  const std::vector<std::vector<uint8_t>>* blocks = nullptr;  // optional <dense>/blocks/mask_<id>.jpg, grey (DPE.cpp:1242-1268)
  std::vector<std::vector<int>> src;
};
struct FusedPoint { float x, y, z; uint8_t b, g, r; };
void fuse_views(const FusionInput& in, std::vector<FusedPoint>* cloud);
}  // namespace dpe_host

#include <math.h>


namespace dpe_host {

namespace {
struct Cam { float K[9], R[9], t[3], C[3]; };

inline void world_point(const Cam& c, float x, float y, float depth, float X[3]) {  // DPE.cpp:1170-1194
  const float px = depth * (x - c.K[2]) / c.K[0], py = depth * (y - c.K[5]) / c.K[4], pz = depth;
  X[0] = c.R[0] * px + c.R[3] * py + c.R[6] * pz + c.C[0];
  X[1] = c.R[1] * px + c.R[4] * py + c.R[7] * pz + c.C[1];
  X[2] = c.R[2] * px + c.R[5] * py + c.R[8] * pz + c.C[2];
}
inline void project(const Cam& c, const float X[3], float* u, float* v, float* d) {  // DPE.cpp:1196-1206
  const float tx = c.R[0] * X[0] + c.R[1] * X[1] + c.R[2] * X[2] + c.t[0];
  const float ty = c.R[3] * X[0] + c.R[4] * X[1] + c.R[5] * X[2] + c.t[1];
  const float tz = c.R[6] * X[0] + c.R[7] * X[1] + c.R[8] * X[2] + c.t[2];
  *d = c.K[6] * tx + c.K[7] * ty + c.K[8] * tz;
  *u = (c.K[0] * tx + c.K[1] * ty + c.K[2] * tz) / *d;
  *v = (c.K[3] * tx + c.K[4] * ty + c.K[5] * tz) / *d;
}
}  // namespace

void fuse_views(const FusionInput& in, std::vector<FusedPoint>* cloud) {
  const int W = in.width, H = in.height, V = in.n_views;
  std::vector<Cam> cams(V);
  for (int v = 0; v < V; ++v) {
    const CamFile& f = (*in.cams)[v];
    for (int i = 0; i < 9; ++i) { cams[v].K[i] = f.K[i]; cams[v].R[i] = f.R[i]; }
    for (int i = 0; i < 3; ++i) cams[v].t[i] = f.t[i];
    for (int j = 0; j < 3; ++j) cams[v].C[j] = -(f.R[0 + j] * f.t[0] + f.R[3 + j] * f.t[1] + f.R[6 + j] * f.t[2]);
  }
  std::vector<std::vector<uint8_t>> masks(V, std::vector<uint8_t>((size_t)W * H, 0));
  cloud->clear();
  std::vector<int> used_x, used_y;
  for (int i = 0; i < V; ++i) {
    const std::vector<int>& src = in.src[i];
    const int num_ngb = (int)src.size();
    used_x.assign(num_ngb, -1); used_y.assign(num_ngb, -1);
    const float* depth = (*in.depth)[i].data();
    const float* normal = (*in.normal)[i].data();
    for (int r = 0; r < H; ++r) {
      for (int c = 0; c < W; ++c) {
        const size_t idx = (size_t)r * W + c;
        if (in.blocks && (*in.blocks)[i][idx] < 128) continue;  // DPE.cpp:1296-1298: only the reference pixel is gated
        if (masks[i][idx] == 1) continue;
        const float ref_depth = depth[idx];
        if (ref_depth <= 0.0f) continue;
        const float* rn = normal + 3 * idx;
        float X[3];
        world_point(cams[i], (float)c, (float)r, ref_depth, X);
        int num_consistent = 0;
        float dynamic_consistency = 0.0f;
        for (int j = 0; j < num_ngb; ++j) { used_x[j] = -1; used_y[j] = -1; }
        for (int j = 0; j < num_ngb; ++j) {
          const int s = src[j];
          if (s < 0) continue;
          float u, v, pd;
          project(cams[s], X, &u, &v, &pd);
          const int sr = (int)(v + 0.5f), sc = (int)(u + 0.5f);
          if (!(sc >= 0 && sc < W && sr >= 0 && sr < H)) continue;
          const size_t sidx = (size_t)sr * W + sc;
          if (masks[s][sidx] == 1) continue;
          const float sd = (*in.depth)[s][sidx];
          if (sd <= 0.0f) continue;
          const float* sn = (*in.normal)[s].data() + 3 * sidx;
          float Y[3], bu, bv;
          world_point(cams[s], (float)sc, (float)sr, sd, Y);
          project(cams[i], Y, &bu, &bv, &pd);
          const float reproj = sqrtf((c - bu) * (c - bu) + (r - bv) * (r - bv));
          const float rel = fabsf(pd - ref_depth) / ref_depth;
          float angle = acosf(rn[0] * sn[0] + rn[1] * sn[1] + rn[2] * sn[2]);
          if (angle != angle) angle = 0.0f;
          if (reproj < 2.0f && rel < 0.01f && angle < 0.174533f) {
            used_x[j] = sc; used_y[j] = sr;
            dynamic_consistency += expf(-(reproj + 200 * rel + angle * 10));
            num_consistent++;
          }
        }
        const float factor = ((*in.state)[i][idx] == DPE_WEAK ? 0.45f : 0.3f);
        if (num_consistent >= 1 && dynamic_consistency > factor * num_consistent) {
          const uint8_t* px = (*in.bgr)[i].data() + 3 * idx;
          float col[3] = {(float)px[0], (float)px[1], (float)px[2]};
          for (int j = 0; j < num_ngb; ++j) {
            if (used_x[j] == -1) continue;
            const size_t sidx = (size_t)used_y[j] * W + used_x[j];
            masks[src[j]][sidx] = 1;
            const uint8_t* sp = (*in.bgr)[src[j]].data() + 3 * sidx;
            col[0] += sp[0]; col[1] += sp[1]; col[2] += sp[2];
          }
          FusedPoint p;
          p.x = X[0]; p.y = X[1]; p.z = X[2];
          p.b = (uint8_t)(col[0] / (num_consistent + 1)); p.g = (uint8_t)(col[1] / (num_consistent + 1));
          p.r = (uint8_t)(col[2] / (num_consistent + 1));
          cloud->push_back(p);
        }
      }
    }
  }
}

}  // namespace dpe_host

// C entry for the tests: arrays are n_views x (H*W [x3]); src: n_views x max_src (-1 padded); K,R: n_views x 9; t: n_views x 3.
// Returns the number of points; writes at most `cap` of them to xyz (x3) / bgr (x3).
extern "C" long fusion_oracle_run(int n_views, int width, int height, const float* depth, const float* normal, const uint8_t* state,
                                  const uint8_t* bgr, const float* K, const float* R, const float* t, const int* src, int max_src,
                                  float* xyz, uint8_t* out_bgr, long cap) {
  using namespace dpe_host;
  const size_t P = (size_t)width * height;
  std::vector<std::vector<float>> d(n_views), n(n_views);
  std::vector<std::vector<uint8_t>> st(n_views), c(n_views);
  std::vector<CamFile> cams(n_views);
  FusionInput in;
  in.width = width; in.height = height; in.n_views = n_views;
  for (int v = 0; v < n_views; ++v) {
    d[v].assign(depth + v * P, depth + (v + 1) * P);
    n[v].assign(normal + v * P * 3, normal + (v + 1) * P * 3);
    st[v].assign(state + v * P, state + (v + 1) * P);
    c[v].assign(bgr + v * P * 3, bgr + (v + 1) * P * 3);
    for (int i = 0; i < 9; ++i) { cams[v].K[i] = K[v * 9 + i]; cams[v].R[i] = R[v * 9 + i]; }
    for (int i = 0; i < 3; ++i) cams[v].t[i] = t[v * 3 + i];
    std::vector<int> s;
    for (int j = 0; j < max_src; ++j) if (src[v * max_src + j] >= 0) s.push_back(src[v * max_src + j]);
    in.src.push_back(s);
  }
  in.depth = &d; in.normal = &n; in.state = &st; in.bgr = &c; in.cams = &cams;
  std::vector<FusedPoint> cloud;
  fuse_views(in, &cloud);
  for (long i = 0; i < (long)cloud.size() && i < cap; ++i) {
    xyz[3 * i] = cloud[i].x; xyz[3 * i + 1] = cloud[i].y; xyz[3 * i + 2] = cloud[i].z;
    out_bgr[3 * i] = cloud[i].b; out_bgr[3 * i + 1] = cloud[i].g; out_bgr[3 * i + 2] = cloud[i].r;
  }
  return (long)cloud.size();
}

// the same with block masks: blocks = n_views x (H*W) grey values (a reference pixel below 128 is skipped)
extern "C" long fusion_oracle_run_blocks(int n_views, int width, int height, const float* depth, const float* normal, const uint8_t* state,
                                         const uint8_t* bgr, const uint8_t* blocks, const float* K, const float* R, const float* t,
                                         const int* src, int max_src, float* xyz, uint8_t* out_bgr, long cap) {
  using namespace dpe_host;
  const size_t P = (size_t)width * height;
  std::vector<std::vector<float>> d(n_views), n(n_views);
  std::vector<std::vector<uint8_t>> st(n_views), c(n_views), bl(n_views);
  std::vector<CamFile> cams(n_views);
  FusionInput in;
  in.width = width; in.height = height; in.n_views = n_views;
  for (int v = 0; v < n_views; ++v) {
    d[v].assign(depth + v * P, depth + (v + 1) * P);
    n[v].assign(normal + v * P * 3, normal + (v + 1) * P * 3);
    st[v].assign(state + v * P, state + (v + 1) * P);
    c[v].assign(bgr + v * P * 3, bgr + (v + 1) * P * 3);
    bl[v].assign(blocks + v * P, blocks + (v + 1) * P);
    for (int i = 0; i < 9; ++i) { cams[v].K[i] = K[v * 9 + i]; cams[v].R[i] = R[v * 9 + i]; }
    for (int i = 0; i < 3; ++i) cams[v].t[i] = t[v * 3 + i];
    std::vector<int> s;
    for (int j = 0; j < max_src; ++j) if (src[v * max_src + j] >= 0) s.push_back(src[v * max_src + j]);
    in.src.push_back(s);
  }
  in.depth = &d; in.normal = &n; in.state = &st; in.bgr = &c; in.cams = &cams; in.blocks = &bl;
  std::vector<FusedPoint> cloud;
  fuse_views(in, &cloud);
  for (long i = 0; i < (long)cloud.size() && i < cap; ++i) {
    xyz[3 * i] = cloud[i].x; xyz[3 * i + 1] = cloud[i].y; xyz[3 * i + 2] = cloud[i].z;
    out_bgr[3 * i] = cloud[i].b; out_bgr[3 * i + 1] = cloud[i].g; out_bgr[3 * i + 2] = cloud[i].r;
  }
  return (long)cloud.size();
}
