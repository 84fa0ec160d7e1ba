// dpe_rng.h — the random stream of the PatchMatch path.
//
// The reference draws every random number from the cuRAND device API's XORWOW generator,
// one 48-byte state per pixel, initialised per (view, stage) by
//   curand_init(seed, /*subsequence=*/p.y, /*offset=*/p.x, &state)      (DPE.cu:1020-1033)
// and then consumed in program order by GenNeighbours, RandomInitialization, the red/black
// sweeps, RANSACToGetFitPlane ... of that RunPatchMatch call.  cuRAND is a third-party
// dependency of the reference (CUDA toolkit; this image: CUDA 12.9, cuRAND 10.3.10); its
// XORWOW is Marsaglia's xorwow ("Xorshift RNGs", JSS 2003, section 3.1: five 32-bit xorshift
// words plus a Weyl counter with increment 362437), seeded and partitioned as the cuRAND
// documentation states: subsequence s starts 2^67 * s draws into the stream, offset o skips
// o further draws.  This header restates that so that, with the reference's seed pinned, a
// pixel sees the same random numbers in the same order as in the reference:
//   * the draw (xorwow_next, xorwow_uniform) is shared by the kernels and the CPU simulator;
//   * the initial states are a per-scale TABLE built on the host (xorwow_init_table): the
//     2^67-draw jump is a 160x160 GF(2) matrix obtained by 67 squarings of the one-draw
//     matrix; row y is the jump applied y times to the seed state, pixel (x, y) is x single
//     draws further.  The table depends only on (seed, width, height), so it is built once
//     per scale and copied into the per-stream state buffer at the start of every stage —
//     the reference re-runs the matrix skip-ahead for every pixel of every view-stage.
// tests/test_rng.py checks states and draws bit for bit against cuRAND's own header compiled for
// the host (oracle/curand_check.cu).
#pragma once
#include <stdint.h>
#include <string.h>
#include <vector>

#if defined(__CUDACC__)
#define DPE_RNG_HD __host__ __device__ __forceinline__
#else
#define DPE_RNG_HD inline
#endif

namespace dpe {

struct Xorwow {
  uint32_t v[5];
  uint32_t d;
};

DPE_RNG_HD void xorwow_step_v(uint32_t v[5]) {
  const uint32_t t = v[0] ^ (v[0] >> 2);
  v[0] = v[1]; v[1] = v[2]; v[2] = v[3]; v[3] = v[4];
  v[4] = (v[4] ^ (v[4] << 4)) ^ (t ^ (t << 1));
}

DPE_RNG_HD uint32_t xorwow_next(Xorwow& s) {
  xorwow_step_v(s.v);
  s.d += 362437u;
  return s.v[4] + s.d;
}

// curand_uniform: (0, 1]
DPE_RNG_HD float xorwow_uniform(Xorwow& s) {
  return (float)xorwow_next(s) * 2.3283064365386963e-10f + 1.1641532182693481e-10f;
}

// ---- host side: seeding and the initial-state table ----------------------------------------
inline void xorwow_seed(uint64_t seed, Xorwow* s) {
  // cuRAND's seed scrambling for XORWOW (salt, odd multipliers, Marsaglia's default words)
  const uint32_t s0 = (uint32_t)seed ^ 0xaad26b49u;
  const uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
  const uint32_t t0 = 1099087573u * s0;
  const uint32_t t1 = 2591861531u * s1;
  s->d = 6615241u + t1 + t0;
  s->v[0] = 123456789u + t0;
  s->v[1] = 362436069u ^ t0;
  s->v[2] = 521288629u + t1;
  s->v[3] = 88675123u ^ t1;
  s->v[4] = 5783321u + t0;
}

// 160x160 matrix over GF(2) acting on the five xorshift words: row[i] is the image of basis
// vector e_i (bit i%32 of word i/32).
struct XorwowMatrix {
  uint32_t row[160][5];
  void apply(const uint32_t in[5], uint32_t out[5]) const {
    uint32_t r[5] = {0, 0, 0, 0, 0};
    for (int w = 0; w < 5; ++w)
      for (int b = 0; b < 32; ++b)
        if ((in[w] >> b) & 1u)
          for (int k = 0; k < 5; ++k) r[k] ^= row[w * 32 + b][k];
    memcpy(out, r, sizeof(r));
  }
  static XorwowMatrix one_draw() {
    XorwowMatrix m;
    for (int i = 0; i < 160; ++i) {
      uint32_t v[5] = {0, 0, 0, 0, 0};
      v[i / 32] = 1u << (i % 32);
      xorwow_step_v(v);
      memcpy(m.row[i], v, sizeof(v));
    }
    return m;
  }
  XorwowMatrix squared() const {
    XorwowMatrix m;
    for (int i = 0; i < 160; ++i) apply(row[i], m.row[i]);
    return m;
  }
  // the jump between cuRAND subsequences: 2^67 draws
  static const XorwowMatrix& subsequence_jump() {
    static const XorwowMatrix J = [] {
      XorwowMatrix m = one_draw();
      for (int i = 0; i < 67; ++i) m = m.squared();
      return m;
    }();
    return J;
  }
};

// state of curand_init(seed, subsequence, offset) — any values (used by the tests)
inline Xorwow xorwow_init(uint64_t seed, uint64_t subsequence, uint64_t offset) {
  Xorwow s;
  xorwow_seed(seed, &s);
  const XorwowMatrix& J = XorwowMatrix::subsequence_jump();
  for (uint64_t i = 0; i < subsequence; ++i) J.apply(s.v, s.v);
  for (uint64_t i = 0; i < offset; ++i) xorwow_step_v(s.v);
  s.d += 362437u * (uint32_t)offset;
  return s;
}

// table[y * width + x] = curand_init(seed, y, x)
inline void xorwow_init_table(uint64_t seed, int width, int height, Xorwow* table) {
  Xorwow row;
  xorwow_seed(seed, &row);
  const XorwowMatrix& J = XorwowMatrix::subsequence_jump();
  for (int y = 0; y < height; ++y) {
    if (y > 0) J.apply(row.v, row.v);
    Xorwow s = row;
    for (int x = 0; x < width; ++x) {
      table[(size_t)y * width + x] = s;
      xorwow_step_v(s.v);
      s.d += 362437u;
    }
  }
}

}  // namespace dpe
