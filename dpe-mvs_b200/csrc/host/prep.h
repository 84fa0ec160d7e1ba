// prep.h — edge / label preparation on the host:
//   GetProblemEdges  main.cpp:331-388      EdgeSegment  DPE.cpp:136-291
//   Roberts          DPE.cpp:9-25          Connect      DPE.cpp:28-134
// plus native restatements of the OpenCV routines those call (no OpenCV C++ in this image):
// resize(INTER_LINEAR) for 8-bit and float images, Canny(L2gradient, aperture 3),
// HoughLinesP, line, threshold.  tests/test_prep.py checks each against cv2 4.13.
#pragma once
#include <stdint.h>
#include <vector>

namespace dpe_host {

struct ImageU8 {
  int rows = 0, cols = 0;
  std::vector<uint8_t> d;
  ImageU8() {}
  ImageU8(int r, int c, uint8_t v = 0) : rows(r), cols(c), d((size_t)r * c, v) {}
  uint8_t& at(int y, int x) { return d[(size_t)y * cols + x]; }
  uint8_t at(int y, int x) const { return d[(size_t)y * cols + x]; }
};

void resize_linear_u8(const ImageU8& src, int dcols, int drows, ImageU8* dst);
void resize_linear_f32(const float* src, int scols, int srows, float* dst, int dcols, int drows);
void canny_l2(const ImageU8& src, double low, double high, ImageU8* dst);
void roberts(const ImageU8& src, ImageU8* dst);
// returns label count (incl. label 0); counts per label in *cnt
int connect(const ImageU8& img, std::vector<int32_t>* label, std::vector<int>* cnt);
void hough_lines_p(const ImageU8& img, double rho, double theta, int threshold, int min_len, int max_gap,
                   std::vector<int>* lines /* x0,y0,x1,y1 per line */);
void draw_line(ImageU8* img, int x0, int y0, int x1, int y1, uint8_t color);

// edges_j / labels_j of one view for scale_size = 2^j (j = 0 is full resolution)
void problem_edges(const ImageU8& gray_full, int scale_size, ImageU8* edge, std::vector<int32_t>* label,
                   int* out_cols, int* out_rows);

}  // namespace dpe_host
