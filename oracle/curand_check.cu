// TEST INFRASTRUCTURE — checks the XORWOW restatement (dpe-mvs_b200/csrc/dpe_rng.h) against the
// CUDA toolkit's own cuRAND device-API code compiled for the host (QUALIFIERS override): states of
// curand_init(seed, y, x), curand() and curand_uniform() must be bit-identical.  Built and run by
// tests/test_rng.py (CPU only).
#define QUALIFIERS static __forceinline__ __host__ __device__
#include <curand_kernel.h>
#include <stdio.h>
#include "../dpe-mvs_b200/csrc/dpe_rng.h"
int main() {
  int bad = 0;
  unsigned long long seeds[2] = {20261018ULL, 0x123456789abcdefULL};
  for (int si = 0; si < 2; ++si)
  for (int y : {0, 1, 2, 7, 599, 1199}) for (int x : {0, 1, 5, 1599}) {
    curandState st; curand_init(seeds[si], y, x, &st);
    dpe::Xorwow s = dpe::xorwow_init(seeds[si], y, x);
    bool ok = s.d == st.d; for (int i = 0; i < 5; ++i) ok = ok && s.v[i] == st.v[i];
    unsigned a = curand(&st), b = dpe::xorwow_next(s);
    float fa = curand_uniform(&st), fb = dpe::xorwow_uniform(s);
    if (!ok || a != b || fa != fb) { bad++; printf("mismatch y=%d x=%d %u %u %g %g\n", y, x, a, b, fa, fb); }
  }
  std::vector<dpe::Xorwow> tab(64 * 48);
  dpe::xorwow_init_table(seeds[0], 64, 48, tab.data());
  for (int y = 0; y < 48; y += 5) for (int x = 0; x < 64; x += 7) {
    curandState st; curand_init(seeds[0], y, x, &st);
    const dpe::Xorwow& s = tab[y * 64 + x];
    bool ok = s.d == st.d; for (int i = 0; i < 5; ++i) ok = ok && s.v[i] == st.v[i];
    if (!ok) { bad++; printf("table mismatch %d %d\n", y, x); }
  }
  printf("bad=%d\n", bad);
  return bad;
}
