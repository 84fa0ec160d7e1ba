// main.cpp — the ./DPE command line, drop-in for the reference's main() (main.cpp:602-635):
//   DPE <dense_folder> [gpu_index=0] [verbose=1] [viz=0] [fusion=0] [depth=1] [normal=0] [weak=0] [edge=0]
// Positional order (viz before fusion) and atoi parsing are the reference's;
// argc is checked before argv[1] is read.
#include <cstdlib>
#include <iostream>

#include "dpe_b200.h"

int main(int argc, char** argv) {
  if (argc < 2) {
    std::cerr << "USAGE: DPE dense_folder" << std::endl;
    return EXIT_FAILURE;
  }
  int gpu_index = 0;
  if (argc >= 3) gpu_index = std::atoi(argv[2]);
  bool verbose = true;
  if (argc >= 4) verbose = std::atoi(argv[3]);
  bool viz = false;
  if (argc >= 5) viz = std::atoi(argv[4]);
  bool fusion = false;
  if (argc >= 6) fusion = std::atoi(argv[5]);
  bool depth = true;
  if (argc >= 7) depth = std::atoi(argv[6]);
  bool normal = false;
  if (argc >= 8) normal = std::atoi(argv[7]);
  bool weak = false;
  if (argc >= 9) weak = std::atoi(argv[8]);
  bool edge = false;
  if (argc >= 10) edge = std::atoi(argv[9]);
  return dpe_run_pipeline(argv[1], gpu_index, verbose, fusion, viz, depth, normal, weak, edge) == 0 ? EXIT_SUCCESS : EXIT_FAILURE;
}
