// dpe_consts.h — host-side folding of cameras into the per-(reference, source, scale)
// constants of dpe_types.h, in double precision.  Shared by the C-ABI layer and by the CPU
// logic simulator used in tests.
//   Rrel = Rs Rr^T, trel = Rs (Cr - Cs)                       (DPE.cu:455-481)
//   A  = Ks Rrel Kr^-1, b  = Ks trel                           (DPE.cu:483-512)
//   Ai = Kr Rr Rs^T Ks^-1, bi = Kr Rr (Cs - Cr)                (DPE.cu:881-913 composed)
#pragma once
#include <math.h>
#include <string.h>
#include "dpe_types.h"

namespace dpe {

struct HostCam {
  double K[9], R[9], t[3], C[3];
  float depth_min, depth_max;
};

// camera centre C = -R^T t in double, rounded to float like ReadCamera (DPE.cpp:362-367)
inline void host_cam_set(HostCam* c, const float K[9], const float R[9], const float t[3], float dmin, float dmax) {
  for (int i = 0; i < 9; ++i) { c->K[i] = K[i]; c->R[i] = R[i]; }
  for (int i = 0; i < 3; ++i) c->t[i] = t[i];
  for (int j = 0; j < 3; ++j)
    c->C[j] = (double)(-(float)((double)R[0 + j] * (double)t[0] + (double)R[3 + j] * (double)t[1] + (double)R[6 + j] * (double)t[2]));
  c->depth_min = dmin; c->depth_max = dmax;
}

inline void mat3_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
inline void mat3_mulT(const double* A, const double* B, double* C) {  // A * B^T
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j * 3] + A[i * 3 + 1] * B[j * 3 + 1] + A[i * 3 + 2] * B[j * 3 + 2];
}
inline void mat3_vec(const double* A, const double* v, double* r) {
  for (int i = 0; i < 3; ++i) r[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}

// K[0],K[2] *= new_cols/cols; K[4],K[5] *= new_rows/rows, in float like DPE.cpp:804-817
inline void scaled_K(const HostCam& c, int w, int h, int full_w, int full_h, double* fx, double* cx, double* fy,
                     double* cy, double* k8) {
  float f0 = (float)c.K[0], f2 = (float)c.K[2], f4 = (float)c.K[4], f5 = (float)c.K[5];
  if (w != full_w || h != full_h) {
    const float sx = w / (float)full_w, sy = h / (float)full_h;
    f0 *= sx; f2 *= sx; f4 *= sy; f5 *= sy;
  }
  *fx = f0; *cx = f2; *fy = f4; *cy = f5; *k8 = c.K[8];
}

// Relative pose in fp32 exactly as the reference's ComputeHomography forms it for every evaluation
// (DPE.cu:455-481): camera centres from R, t, then R_rel = Rs Rr^T and t_rel = Rs (Cr - Cs).  Host version;
// the CUDA side recomputes it on the device with the same expressions (k_relative_pose) so that the values
// carry the device compiler's fused multiply-adds.
#if defined(__CUDACC__)
__host__ __device__
#endif
inline void relative_pose_ref(const float* rR, const float* rt, const float* sR, const float* st, float* Rrel, float* trel) {
  float ref_C[3], src_C[3];
  ref_C[0] = -(rR[0] * rt[0] + rR[3] * rt[1] + rR[6] * rt[2]);
  ref_C[1] = -(rR[1] * rt[0] + rR[4] * rt[1] + rR[7] * rt[2]);
  ref_C[2] = -(rR[2] * rt[0] + rR[5] * rt[1] + rR[8] * rt[2]);
  src_C[0] = -(sR[0] * st[0] + sR[3] * st[1] + sR[6] * st[2]);
  src_C[1] = -(sR[1] * st[0] + sR[4] * st[1] + sR[7] * st[2]);
  src_C[2] = -(sR[2] * st[0] + sR[5] * st[1] + sR[8] * st[2]);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      Rrel[i * 3 + j] = sR[i * 3 + 0] * rR[j * 3 + 0] + sR[i * 3 + 1] * rR[j * 3 + 1] + sR[i * 3 + 2] * rR[j * 3 + 2];
  const float C_relative[3] = {ref_C[0] - src_C[0], ref_C[1] - src_C[1], ref_C[2] - src_C[2]};
  for (int i = 0; i < 3; ++i) trel[i] = sR[i * 3 + 0] * C_relative[0] + sR[i * 3 + 1] * C_relative[1] + sR[i * 3 + 2] * C_relative[2];
}

inline void fold_ref(const HostCam& r, int w, int h, int full_w, int full_h, int view, RefConst* rc) {
  memset(rc, 0, sizeof(*rc));
  double fx, cx, fy, cy, k8;
  scaled_K(r, w, h, full_w, full_h, &fx, &cx, &fy, &cy, &k8);
  rc->W = w; rc->H = h;
  rc->fx = (float)fx; rc->cx = (float)cx; rc->fy = (float)fy; rc->cy = (float)cy;
  for (int i = 0; i < 9; ++i) rc->R[i] = (float)r.R[i];
  for (int i = 0; i < 3; ++i) rc->t[i] = (float)r.t[i];
  for (int i = 0; i < 3; ++i) rc->c[i] = (float)r.C[i];
  for (int i = 0; i < 9; ++i) rc->K9[i] = (float)r.K[i];
  rc->K9[0] = (float)fx; rc->K9[2] = (float)cx; rc->K9[4] = (float)fy; rc->K9[5] = (float)cy;
  rc->depth_min = r.depth_min * 0.6f;  // DPE.cpp:788-789
  rc->depth_max = r.depth_max * 1.2f;
  rc->view = view;
}

inline void fold_pair(const HostCam& r, const HostCam& s, int w, int h, int full_w, int full_h, SrcConst* sc) {
  double fx, cx, fy, cy, k8, sfx, scx, sfy, scy, sk8;
  scaled_K(r, w, h, full_w, full_h, &fx, &cx, &fy, &cy, &k8);
  scaled_K(s, w, h, full_w, full_h, &sfx, &scx, &sfy, &scy, &sk8);
  const double Kr[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1.0};
  const double Kr_inv[9] = {1.0 / fx, 0, -cx / fx, 0, 1.0 / fy, -cy / fy, 0, 0, 1.0};
  const double Ks[9] = {sfx, 0, scx, 0, sfy, scy, 0, 0, sk8};
  const double Ks_inv[9] = {1.0 / sfx, 0, -scx / sfx, 0, 1.0 / sfy, -scy / sfy, 0, 0, 1.0 / sk8};
  double Rrel[9], tmp[9], A[9], Ai[9], RrelT[9];
  mat3_mulT(s.R, r.R, Rrel);
  const double dC[3] = {r.C[0] - s.C[0], r.C[1] - s.C[1], r.C[2] - s.C[2]};
  const double dCi[3] = {-dC[0], -dC[1], -dC[2]};
  double trel[3], b[3], bi[3], t2[3];
  mat3_vec(s.R, dC, trel);
  mat3_mul(Ks, Rrel, tmp); mat3_mul(tmp, Kr_inv, A);
  mat3_vec(Ks, trel, b);
  mat3_mulT(r.R, s.R, RrelT);
  mat3_mul(Kr, RrelT, tmp); mat3_mul(tmp, Ks_inv, Ai);
  mat3_vec(r.R, dCi, t2); mat3_vec(Kr, t2, bi);
  for (int i = 0; i < 9; ++i) { sc->A[i] = (float)A[i]; sc->Ai[i] = (float)Ai[i]; }
  for (int i = 0; i < 3; ++i) { sc->b[i] = (float)b[i]; sc->bi[i] = (float)bi[i]; }
  const float c0 = (float)r.C[0] - (float)s.C[0], c1 = (float)r.C[1] - (float)s.C[1], c2 = (float)r.C[2] - (float)s.C[2];
  sc->baseline = sqrtf((float)((double)(c0 * c0 + c1 * c1 + c2 * c2)));  // DPE.cu:2640-2645
  sc->width = (float)w; sc->height = (float)h;
  // fp32 data of the reference-order arithmetic
  for (int i = 0; i < 9; ++i) sc->sR[i] = (float)s.R[i];
  for (int i = 0; i < 3; ++i) { sc->st[i] = (float)s.t[i]; sc->sc3[i] = (float)s.C[i]; }
  for (int i = 0; i < 9; ++i) sc->sK[i] = (float)s.K[i];
  sc->sK[0] = (float)sfx; sc->sK[2] = (float)scx; sc->sK[4] = (float)sfy; sc->sK[5] = (float)scy;
  float rR[9], rt[3];
  for (int i = 0; i < 9; ++i) rR[i] = (float)r.R[i];
  for (int i = 0; i < 3; ++i) rt[i] = (float)r.t[i];
  relative_pose_ref(rR, rt, sc->sR, sc->st, sc->Rrel, sc->trel);
}

}  // namespace dpe
