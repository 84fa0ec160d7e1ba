// prep.cpp — see prep.h.
#include "prep.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace dpe_host {

static inline int cv_round(double v) { return (int)lrint(v); }  // round half to even, like cvRound
static inline int cv_floor(double v) { return (int)floor(v); }
static inline int c_round(float v) { return (int)roundf(v); }   // std::round, half away from zero

// ---- cv::resize(INTER_LINEAR), 8-bit single channel ----------------------------------------
// Exact 2x2 decimation is routed to the area-fast kernel like OpenCV does
// ("INTER_AREA (fast) also is equal to INTER_LINEAR" for scale 2); everything else is the
// fixed-point separable path: 11-bit coefficients, horizontal pass in int, vertical pass
// ((b*(S>>4))>>16 summed, +2, >>2).
void resize_linear_u8(const ImageU8& src, int dcols, int drows, ImageU8* dst) {
  if (dcols <= 0 || drows <= 0 || src.cols <= 0 || src.rows <= 0) {  // cv::resize throws here; an empty image is the answer
    *dst = ImageU8(std::max(drows, 0), std::max(dcols, 0));
    return;
  }
  ImageU8 out(drows, dcols);
  const int sw = src.cols, sh = src.rows;
  const double scale_x = (double)sw / dcols, scale_y = (double)sh / drows;
  if (sw == 2 * dcols && sh == 2 * drows) {
    for (int y = 0; y < drows; ++y)
      for (int x = 0; x < dcols; ++x)
        out.at(y, x) = (uint8_t)((src.at(2 * y, 2 * x) + src.at(2 * y, 2 * x + 1) + src.at(2 * y + 1, 2 * x) + src.at(2 * y + 1, 2 * x + 1) + 2) >> 2);
    *dst = out;
    return;
  }
  std::vector<int> xofs(dcols), yofs(drows);
  std::vector<short> ia(2 * dcols), ib(2 * drows);
  for (int dx = 0; dx < dcols; ++dx) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = cv_floor(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    xofs[dx] = sx;
    ia[2 * dx] = (short)cv_round((1.f - fx) * 2048.f);
    ia[2 * dx + 1] = (short)cv_round(fx * 2048.f);
  }
  for (int dy = 0; dy < drows; ++dy) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = cv_floor(fy);
    fy -= sy;
    yofs[dy] = sy;
    ib[2 * dy] = (short)cv_round((1.f - fy) * 2048.f);
    ib[2 * dy + 1] = (short)cv_round(fy * 2048.f);
  }
  std::vector<int> r0(dcols), r1(dcols);
  auto hrow = [&](int sy, std::vector<int>& r) {
    sy = std::min(std::max(sy, 0), sh - 1);
    const uint8_t* S = &src.d[(size_t)sy * sw];
    for (int dx = 0; dx < dcols; ++dx) {
      const int sx = xofs[dx];
      const int s1 = sx + 1 < sw ? S[sx + 1] : S[sx];
      r[dx] = S[sx] * ia[2 * dx] + s1 * ia[2 * dx + 1];
    }
  };
  for (int dy = 0; dy < drows; ++dy) {
    hrow(yofs[dy], r0);
    hrow(yofs[dy] + 1, r1);
    const int b0 = ib[2 * dy], b1 = ib[2 * dy + 1];
    for (int dx = 0; dx < dcols; ++dx) {
      const int v = (((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2;
      out.at(dy, dx) = (uint8_t)std::min(std::max(v, 0), 255);
    }
  }
  *dst = out;
}

// cv::resize(INTER_LINEAR) for CV_32FC1 (same arithmetic as the GPU pyramid kernel)
void resize_linear_f32(const float* src, int sw, int sh, float* dst, int dw, int dh) {
  if (dw <= 0 || dh <= 0 || sw <= 0 || sh <= 0) return;
  const double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
  for (int dy = 0; dy < dh; ++dy) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = cv_floor(fy);
    fy -= sy;
    if (sy < 0) { fy = 0.f; sy = 0; }
    if (sy >= sh - 1) { fy = 0.f; sy = sh - 1; }
    const int sy1 = std::min(sy + 1, sh - 1);
    for (int dx = 0; dx < dw; ++dx) {
      float fx = (float)((dx + 0.5) * scale_x - 0.5);
      int sx = cv_floor(fx);
      fx -= sx;
      if (sx < 0) { fx = 0.f; sx = 0; }
      if (sx >= sw - 1) { fx = 0.f; sx = sw - 1; }
      const int sx1 = std::min(sx + 1, sw - 1);
      const float a0 = 1.f - fx, a1 = fx, b0 = 1.f - fy, b1 = fy;
      const float q0 = src[(size_t)sy * sw + sx] * a0 + src[(size_t)sy * sw + sx1] * a1;
      const float q1 = src[(size_t)sy1 * sw + sx] * a0 + src[(size_t)sy1 * sw + sx1] * a1;
      dst[(size_t)dy * dw + dx] = q0 * b0 + q1 * b1;
    }
  }
}

// ---- cv::Canny(src, dst, low, high, 3, L2gradient = true) ----------------------------------
void canny_l2(const ImageU8& src, double low_thresh, double high_thresh, ImageU8* dst) {
  const int rows = src.rows, cols = src.cols;
  if (low_thresh > high_thresh) std::swap(low_thresh, high_thresh);
  low_thresh = std::min(32767.0, low_thresh);
  high_thresh = std::min(32767.0, high_thresh);
  if (low_thresh > 0) low_thresh *= low_thresh;
  if (high_thresh > 0) high_thresh *= high_thresh;
  const int low = cv_floor(low_thresh), high = cv_floor(high_thresh);
  // Sobel 3x3, BORDER_REPLICATE
  std::vector<short> gx((size_t)rows * cols), gy((size_t)rows * cols);
  const int mstep = cols + 2;
  std::vector<int> mag((size_t)(rows + 2) * mstep, 0);
  auto px = [&](int y, int x) -> int {
    y = std::min(std::max(y, 0), rows - 1);
    x = std::min(std::max(x, 0), cols - 1);
    return src.d[(size_t)y * cols + x];
  };
  for (int y = 0; y < rows; ++y) {
    for (int x = 0; x < cols; ++x) {
      const int a = px(y - 1, x - 1), b = px(y - 1, x), c = px(y - 1, x + 1);
      const int d = px(y, x - 1), f = px(y, x + 1);
      const int g = px(y + 1, x - 1), h = px(y + 1, x), i = px(y + 1, x + 1);
      const int dx = (c + 2 * f + i) - (a + 2 * d + g);
      const int dy = (g + 2 * h + i) - (a + 2 * b + c);
      gx[(size_t)y * cols + x] = (short)dx;
      gy[(size_t)y * cols + x] = (short)dy;
      mag[(size_t)(y + 1) * mstep + x + 1] = dx * dx + dy * dy;
    }
  }
  // non-maximum suppression: map 1 = not an edge, 0 = candidate, 2 = edge
  std::vector<uint8_t> map((size_t)(rows + 2) * mstep, 1);
  std::vector<int> stack;
  const int TG22 = (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5);
  for (int y = 0; y < rows; ++y) {
    const int* ma = &mag[(size_t)(y + 1) * mstep + 1];
    const int* mp = ma - mstep;
    const int* mn = ma + mstep;
    uint8_t* pm = &map[(size_t)(y + 1) * mstep + 1];
    for (int j = 0; j < cols; ++j) {
      const int m = ma[j];
      bool is_max = false;
      if (m > low) {
        const int xs = gx[(size_t)y * cols + j], ys = gy[(size_t)y * cols + j];
        const int x = abs(xs), yy = abs(ys) << 15;
        const int tg22x = x * TG22;
        if (yy < tg22x) {
          is_max = (m > ma[j - 1] && m >= ma[j + 1]);
        } else {
          const int tg67x = tg22x + (x << 16);
          if (yy > tg67x) is_max = (m > mp[j] && m >= mn[j]);
          else {
            const int s = (xs ^ ys) < 0 ? -1 : 1;
            is_max = (m > mp[j - s] && m > mn[j + s]);
          }
        }
      }
      if (is_max) {
        if (m > high) { pm[j] = 2; stack.push_back((y + 1) * mstep + j + 1); }
        else pm[j] = 0;
      } else pm[j] = 1;
    }
  }
  // hysteresis, 8-connectivity
  while (!stack.empty()) {
    const int p = stack.back();
    stack.pop_back();
    const int nb[8] = {p - mstep - 1, p - mstep, p - mstep + 1, p - 1, p + 1, p + mstep - 1, p + mstep, p + mstep + 1};
    for (int k = 0; k < 8; ++k)
      if (map[nb[k]] == 0) { map[nb[k]] = 2; stack.push_back(nb[k]); }
  }
  ImageU8 out(rows, cols);
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) out.at(y, x) = map[(size_t)(y + 1) * mstep + x + 1] == 2 ? 255 : 0;
  *dst = out;
}

// ---- Roberts, DPE.cpp:9-25 -----------------------------------------------------------------
void roberts(const ImageU8& src, ImageU8* dst) {
  ImageU8 out(src.rows, src.cols);
  for (int i = 0; i < src.rows; i++) {
    for (int j = 0; j < src.cols; j++) {
      int t1 = 50, t2 = 50;  // image border
      if (i > 0 && i < src.rows - 1 && j > 0 && j < src.cols - 1) {
        t1 = (int)src.at(i, j) - (int)src.at(i + 1, j + 1);
        t2 = (int)src.at(i + 1, j) - (int)src.at(i, j + 1);
      }
      out.at(i, j) = (uint8_t)((int)sqrt((double)(t1 * t1 + t2 * t2)) & 0xFF);
    }
  }
  *dst = out;
}

// ---- Connect, DPE.cpp:28-134: two-pass labelling of the zero set (4-connectivity) with the
// reference's union rule (the larger provisional label is re-parented to the smaller one).
int connect(const ImageU8& img, std::vector<int32_t>* label_out, std::vector<int>* cnt) {
  const int rows = img.rows, cols = img.cols;
  std::vector<int32_t>& label = *label_out;
  label.assign((size_t)rows * cols, 0);
  std::vector<int> parent(1, 0);
  for (int y = 0; y < rows; y++) {
    for (int x = 0; x < cols; x++) {
      const size_t c = (size_t)y * cols + x;
      if (img.d[c] == 255) { label[c] = 0; continue; }
      const bool left = x > 0 && img.d[c] == 0 && img.d[c - 1] == 0;
      const bool up = y > 0 && img.d[c] == 0 && img.d[c - cols] == 0;
      if (left) label[c] = label[c - 1];
      if (up) label[c] = label[c - cols];
      if (!left && !up) {
        label[c] = (int32_t)parent.size();
        parent.push_back((int)parent.size());
      } else if (left && up) {
        const int ll = label[c - 1], ul = label[c - cols];
        if (ll > ul) { parent[ll] = ul; label[c] = ul; }
        else if (ll < ul) { parent[ul] = ll; label[c] = ll; }
      }
    }
  }
  const int n = (int)parent.size();
  for (int i = 1; i < n; i++) {
    int cur = parent[i], pre = parent[cur];
    while (pre != cur) { cur = pre; pre = parent[pre]; }
    parent[i] = cur;
  }
  int label_num = 1;
  std::vector<int> mapping(n, 0);
  for (int i = 1; i < n; i++)
    if (parent[i] == i) mapping[i] = label_num++;
  for (int i = 1; i < n; i++) parent[i] = mapping[parent[i]];
  cnt->assign(label_num, 0);
  for (size_t i = 0; i < label.size(); i++) { label[i] = parent[label[i]]; (*cnt)[label[i]]++; }
  return label_num;
}

// ---- cv::HoughLinesP (progressive probabilistic Hough transform) ---------------------------
namespace {
struct CvRng {  // cv::RNG, multiply-with-carry
  uint64_t state;
  explicit CvRng(uint64_t s) : state(s) {}
  unsigned next() { state = (uint64_t)(unsigned)state * 4164903690U + (unsigned)(state >> 32); return (unsigned)state; }
  int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};
}  // namespace

void hough_lines_p(const ImageU8& image, double rho, double theta, int threshold, int line_length, int line_gap,
                   std::vector<int>* lines) {
  lines->clear();
  const int width = image.cols, height = image.rows;
  const int numangle = cv_round(3.14159265358979323846 / theta);
  const int numrho = cv_round(((width + height) * 2 + 1) / rho);
  const float irho = 1.f / (float)rho;
  std::vector<int> accum((size_t)numangle * numrho, 0);
  std::vector<uint8_t> mask((size_t)width * height, 0);
  std::vector<float> trig((size_t)numangle * 2);
  for (int n = 0; n < numangle; n++) {
    trig[n * 2] = (float)(cos((double)n * theta) * irho);
    trig[n * 2 + 1] = (float)(sin((double)n * theta) * irho);
  }
  std::vector<int> nzx, nzy;
  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++)
      if (image.at(y, x)) { mask[(size_t)y * width + x] = 1; nzx.push_back(x); nzy.push_back(y); }
  int count = (int)nzx.size();
  CvRng rng((uint64_t)-1);
  const int shift = 16;
  for (; count > 0; count--) {
    const int idx = rng.uniform(0, count);
    int max_val = threshold - 1, max_n = 0;
    const int px = nzx[idx], py = nzy[idx];
    int line_end[2][2] = {{0, 0}, {0, 0}};  // [k] = {x, y}
    const int i = py, j = px;
    nzx[idx] = nzx[count - 1]; nzy[idx] = nzy[count - 1];
    if (!mask[(size_t)i * width + j]) continue;
    int* adata = accum.data();
    for (int n = 0; n < numangle; n++, adata += numrho) {
      int r = cv_round(j * trig[n * 2] + i * trig[n * 2 + 1]);
      r += (numrho - 1) / 2;
      const int val = ++adata[r];
      if (max_val < val) { max_val = val; max_n = n; }
    }
    if (max_val < threshold) continue;
    const float a = -trig[max_n * 2 + 1], b = trig[max_n * 2];
    int x0 = j, y0 = i, dx0, dy0, xflag;
    if (fabs(a) > fabs(b)) {
      xflag = 1;
      dx0 = a > 0 ? 1 : -1;
      dy0 = cv_round(b * (1 << shift) / fabs(a));
      y0 = (y0 << shift) + (1 << (shift - 1));
    } else {
      xflag = 0;
      dy0 = b > 0 ? 1 : -1;
      dx0 = cv_round(a * (1 << shift) / fabs(b));
      x0 = (x0 << shift) + (1 << (shift - 1));
    }
    for (int k = 0; k < 2; k++) {
      int gap = 0, x = x0, y = y0, dx = dx0, dy = dy0;
      if (k > 0) { dx = -dx; dy = -dy; }
      for (;; x += dx, y += dy) {
        int i1, j1;
        if (xflag) { j1 = x; i1 = y >> shift; } else { j1 = x >> shift; i1 = y; }
        if (j1 < 0 || j1 >= width || i1 < 0 || i1 >= height) break;
        if (mask[(size_t)i1 * width + j1]) { gap = 0; line_end[k][1] = i1; line_end[k][0] = j1; }
        else if (++gap > line_gap) break;
      }
    }
    const bool good_line = abs(line_end[1][0] - line_end[0][0]) >= line_length || abs(line_end[1][1] - line_end[0][1]) >= line_length;
    for (int k = 0; k < 2; k++) {
      int x = x0, y = y0, dx = dx0, dy = dy0;
      if (k > 0) { dx = -dx; dy = -dy; }
      for (;; x += dx, y += dy) {
        int i1, j1;
        if (xflag) { j1 = x; i1 = y >> shift; } else { j1 = x >> shift; i1 = y; }
        uint8_t* m = &mask[(size_t)i1 * width + j1];
        if (*m) {
          if (good_line) {
            int* ad = accum.data();
            for (int n = 0; n < numangle; n++, ad += numrho) {
              int r = cv_round(j1 * trig[n * 2] + i1 * trig[n * 2 + 1]);
              r += (numrho - 1) / 2;
              ad[r]--;
            }
          }
          *m = 0;
        }
        if (i1 == line_end[k][1] && j1 == line_end[k][0]) break;
      }
    }
    if (good_line) {
      lines->push_back(line_end[0][0]); lines->push_back(line_end[0][1]);
      lines->push_back(line_end[1][0]); lines->push_back(line_end[1][1]);
    }
  }
}

// ---- cv::line, thickness 1, 8-connected (LineIterator, left to right) ----------------------
void draw_line(ImageU8* img, int x0, int y0, int x1, int y1, uint8_t color) {
  // end points produced by hough_lines_p are inside the image, so no clipping is needed
  int dx = x1 - x0, dy = y1 - y0;
  int sx = 1, sy = 1;
  if (dx < 0) { dx = -dx; dy = -dy; x0 = x1; y0 = y1; }
  if (dy < 0) { dy = -dy; sy = -1; }
  const bool vert = dy > dx;
  int major = dx, minor = dy;
  if (vert) std::swap(major, minor);
  int err = major - (minor + minor);
  const int plus_delta = major + major, minus_delta = -(minor + minor);
  int x = x0, y = y0;
  for (int i = 0; i <= major; ++i) {
    if (x >= 0 && x < img->cols && y >= 0 && y < img->rows) img->at(y, x) = color;
    const bool m = err < 0;
    err += minus_delta + (m ? plus_delta : 0);
    if (vert) { y += sy; if (m) x += sx; }
    else { x += sx; if (m) y += sy; }
  }
}

static void threshold_binary(ImageU8* img, int thr) {
  for (auto& v : img->d) v = v > thr ? 255 : 0;
}

static void border_clean(ImageU8* dst) {  // DPE.cpp:239-250
  const int rows = dst->rows, cols = dst->cols;
  if (rows < 2 || cols < 2) return;  // the reference indexes column 1 / row 1 regardless; a one-pixel-wide image has none
  for (int y = 0; y < rows; y++) {
    if (dst->at(y, 1) == 0) dst->at(y, 0) = 0;
    if (dst->at(y, cols - 2) == 0) dst->at(y, cols - 1) = 0;
  }
  for (int x = 0; x < cols; x++) {
    if (dst->at(1, x) == 0) dst->at(0, x) = 0;
    if (dst->at(rows - 2, x) == 0) dst->at(rows - 1, x) = 0;
  }
}

// EdgeSegment(scale, srcImage, mode 0, use_canny = true), DPE.cpp:136-253
static void edge_map(const ImageU8& src, ImageU8* edge) {
  const int rows = src.rows, cols = src.cols;
  float hist[256] = {0};
  for (uint8_t v : src.d) hist[v]++;
  const int half = rows * cols / 2;
  int median_val = -1, acc = 0;
  for (int i = 0; i < 255; i++) {
    acc = acc + (int)hist[i];
    if (acc > half) { median_val = i; break; }
  }
  const float sigma = 0.67f;
  const int threshold1 = (int)((1 - sigma) * median_val);
  const int threshold2 = median_val;
  ImageU8 dst;
  canny_l2(src, threshold1, threshold2, &dst);
  // resize to the same size is the identity; threshold(>4) keeps 0/255
  threshold_binary(&dst, 4);
  border_clean(&dst);
  *edge = dst;
}

// EdgeSegment(scale, srcImage = full-resolution image, mode 1, use_canny = false)
static void label_map(int scale, const ImageU8& src, std::vector<int32_t>* label, int* out_cols, int* out_rows) {
  const int rows = src.rows, cols = src.cols;
  const int weak_tex_num = (int)(1.0 * rows * cols / ((1024 << scale) << scale));
  ImageU8 half, quarter, dst;
  resize_linear_u8(src, cols / 2, rows / 2, &half);
  resize_linear_u8(half, half.cols / 2, half.rows / 2, &quarter);
  const int m = std::min(quarter.cols, quarter.rows);
  const int houthr = (int)(m / 30.0);
  roberts(quarter, &dst);
  threshold_binary(&dst, 4);
  std::vector<int32_t> lab0;
  std::vector<int> cnt0;
  connect(dst, &lab0, &cnt0);
  ImageU8 img_weak(dst.rows, dst.cols);
  std::vector<int> lines;
  for (size_t k = 1; k < cnt0.size(); k++) {
    if (cnt0[k] < weak_tex_num) continue;
    const int wi = (int)k;
    std::fill(img_weak.d.begin(), img_weak.d.end(), 0);
    for (int y = 0; y < dst.rows; y++) {
      for (int x = 0; x < dst.cols; x++) {
        const size_t c = (size_t)y * dst.cols + x;
        if (lab0[c] == wi) continue;
        bool border = false;
        if (x > 0 && lab0[c - 1] == wi) border = true;
        if (x < dst.cols - 1 && lab0[c + 1] == wi) border = true;
        if (y > 0 && lab0[c - dst.cols] == wi) border = true;
        if (y < dst.rows - 1 && lab0[c + dst.cols] == wi) border = true;
        if (border) img_weak.d[c] = 255;
      }
    }
    hough_lines_p(img_weak, 1.0, 3.14159265358979323846 / 180, houthr, houthr, houthr, &lines);
    for (size_t i = 0; i + 3 < lines.size(); i += 4) draw_line(&dst, lines[i], lines[i + 1], lines[i + 2], lines[i + 3], 255);
  }
  const float factor = 1.0f / (float)(1 << scale);
  const int new_cols = c_round(cols * factor), new_rows = c_round(rows * factor);
  ImageU8 scaled;
  resize_linear_u8(dst, new_cols, new_rows, &scaled);
  threshold_binary(&scaled, 4);
  border_clean(&scaled);
  std::vector<int> cnt;
  connect(scaled, label, &cnt);
  for (auto& l : *label)
    if (l != 0 && cnt[l] <= weak_tex_num) l = -1;
  *out_cols = new_cols; *out_rows = new_rows;
}

void problem_edges(const ImageU8& gray_full, int scale_size, ImageU8* edge, std::vector<int32_t>* label, int* out_cols,
                   int* out_rows) {
  int scale = 0;
  while ((1 << scale) < scale_size) scale++;
  // main.cpp:338-346: float convert, bilinear resize, back to 8 bit (rounds half to even)
  const float factor = 1.0f / (float)scale_size;
  const int new_cols = c_round(gray_full.cols * factor), new_rows = c_round(gray_full.rows * factor);
  ImageU8 scaled(new_rows, new_cols);
  if (new_cols == gray_full.cols && new_rows == gray_full.rows) {
    scaled = gray_full;
  } else {
    std::vector<float> f(gray_full.d.begin(), gray_full.d.end()), g((size_t)new_rows * new_cols);
    resize_linear_f32(f.data(), gray_full.cols, gray_full.rows, g.data(), new_cols, new_rows);
    for (size_t i = 0; i < g.size(); ++i) scaled.d[i] = (uint8_t)std::min(std::max(cv_round(g[i]), 0), 255);
  }
  edge_map(scaled, edge);
  int lc, lr;
  label_map(scale, gray_full, label, &lc, &lr);
  *out_cols = new_cols; *out_rows = new_rows;
}

}  // namespace dpe_host

// C hooks for the CPU tests (tests/test_prep.py compares each routine with cv2)
extern "C" {
#define DPE_TEST_API __attribute__((visibility("default")))
DPE_TEST_API void dpe_host_resize_u8(const uint8_t* src, int scols, int srows, uint8_t* dst, int dcols, int drows) {
  dpe_host::ImageU8 s(srows, scols), d;
  memcpy(s.d.data(), src, s.d.size());
  dpe_host::resize_linear_u8(s, dcols, drows, &d);
  memcpy(dst, d.d.data(), d.d.size());
}
DPE_TEST_API void dpe_host_resize_f32(const float* src, int scols, int srows, float* dst, int dcols, int drows) {
  dpe_host::resize_linear_f32(src, scols, srows, dst, dcols, drows);
}
DPE_TEST_API void dpe_host_canny(const uint8_t* src, int cols, int rows, double low, double high, uint8_t* dst) {
  dpe_host::ImageU8 s(rows, cols), d;
  memcpy(s.d.data(), src, s.d.size());
  dpe_host::canny_l2(s, low, high, &d);
  memcpy(dst, d.d.data(), d.d.size());
}
DPE_TEST_API int dpe_host_hough(const uint8_t* src, int cols, int rows, int thr, int min_len, int max_gap, int* out, int max_lines) {
  dpe_host::ImageU8 s(rows, cols);
  memcpy(s.d.data(), src, s.d.size());
  std::vector<int> lines;
  dpe_host::hough_lines_p(s, 1.0, 3.14159265358979323846 / 180, thr, min_len, max_gap, &lines);
  const int n = std::min((int)lines.size() / 4, max_lines);
  memcpy(out, lines.data(), (size_t)n * 4 * sizeof(int));
  return (int)lines.size() / 4;
}
DPE_TEST_API void dpe_host_line(uint8_t* img, int cols, int rows, int x0, int y0, int x1, int y1) {
  dpe_host::ImageU8 s(rows, cols);
  memcpy(s.d.data(), img, s.d.size());
  dpe_host::draw_line(&s, x0, y0, x1, y1, 255);
  memcpy(img, s.d.data(), s.d.size());
}
DPE_TEST_API void dpe_host_problem_edges(const uint8_t* gray, int cols, int rows, int scale_size, uint8_t* edge, int32_t* label) {
  dpe_host::ImageU8 s(rows, cols), e;
  memcpy(s.d.data(), gray, s.d.size());
  std::vector<int32_t> l;
  int oc, orr;
  dpe_host::problem_edges(s, scale_size, &e, &l, &oc, &orr);
  memcpy(edge, e.d.data(), e.d.size());
  memcpy(label, l.data(), l.size() * sizeof(int32_t));
}
}
