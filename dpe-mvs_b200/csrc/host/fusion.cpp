// fusion.cpp — binary PLY writer of the fused cloud (ExportPointCloud, DPE.cpp:532-572).  The fusion itself
// runs on the device: csrc/dpe_fusion.cu through dpe_fuse_* of the C ABI.
#include "fusion.h"

#include <stdio.h>

#include <fstream>

namespace dpe_host {

bool write_ply(const std::string& path, const float* xyz, const uint8_t* bgr, size_t n) {  // DPE.cpp:532-572
  std::ofstream out(path, std::ios::binary);
  if (!out.is_open()) return false;
  out << "ply\nformat binary_little_endian 1.0\nelement vertex " << (long long)n << "\n";
  out << "property float x\nproperty float y\nproperty float z\n";
  out << "property uchar diffuse_blue\nproperty uchar diffuse_green\nproperty uchar diffuse_red\nend_header\n";
  std::vector<char> buf(n * 15);
  for (size_t i = 0; i < n; ++i) {
    memcpy(&buf[i * 15], xyz + 3 * i, 12);
    buf[i * 15 + 12] = (char)bgr[3 * i]; buf[i * 15 + 13] = (char)bgr[3 * i + 1]; buf[i * 15 + 14] = (char)bgr[3 * i + 2];
  }
  out.write(buf.data(), (std::streamsize)buf.size());
  return out.good();
}

// the same file from ready-made vertex records (15 bytes each: dpe_fuse_get_ply_records), several parts in order
bool write_ply_records(const std::string& path, const std::vector<const uint8_t*>& parts, const std::vector<size_t>& counts) {
  size_t n = 0;
  for (size_t c : counts) n += c;
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  fprintf(f, "ply\nformat binary_little_endian 1.0\nelement vertex %lld\n", (long long)n);
  fprintf(f, "property float x\nproperty float y\nproperty float z\n");
  fprintf(f, "property uchar diffuse_blue\nproperty uchar diffuse_green\nproperty uchar diffuse_red\nend_header\n");
  bool ok = true;
  for (size_t i = 0; i < parts.size(); ++i)
    if (counts[i]) ok = ok && fwrite(parts[i], 15, counts[i], f) == counts[i];
  ok = fclose(f) == 0 && ok;
  return ok;
}

}  // namespace dpe_host
