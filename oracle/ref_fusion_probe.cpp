// TEST INFRASTRUCTURE — runs the reference's OWN depth-map fusion (RunFusion, DPE.cpp:1220-1370) on a scene
// folder whose DPE/<view>/{depths.dmb, normals.dmb, weak.bin} were written beforehand, and leaves DPE/DPE.ply.
// The reference's main.cpp is included where it lies (its main() renamed), so the problem list comes from its own
// GenerateSampleList (main.cpp:264-308) and the cloud from its own ExportPointCloud (DPE.cpp:532-572).  RunFusion is
// plain CPU code: this probe needs no GPU, so the golden cloud of tests/golden/ref_fusion_c1.npz is made here in
// the build container (oracle/make_fusion_golden.py) and pins oracle/fusion_oracle.cpp, the checker of the device
// fusion.  Built by oracle/Makefile into _ref/ref_fusion_probe; never linked into the product.
#define main dpe_reference_main
#include "main.cpp"
#undef main

int main(int argc, char** argv) {
  if (argc < 2) { std::cerr << "usage: ref_fusion_probe <dense_folder>\n"; return 2; }
  const path dense_folder(argv[1]);
  std::vector<Problem> problems;
  GenerateSampleList(dense_folder, problems, false);
  if (problems.empty()) { std::cerr << "no problems in pair.txt\n"; return 1; }
  RunFusion(dense_folder, problems);
  return 0;
}
