// jpeg_luma.cpp — baseline JPEG -> 8-bit luma, bit-identical to libjpeg / libjpeg-turbo's default decode
// (Huffman, integer dequantisation, the "accurate integer" inverse DCT jpeg_idct_islow, range limit), which is
// what cv::imread(path, cv::IMREAD_GRAYSCALE) returns for a greyscale or YCbCr JPEG — the pixels the reference
// feeds its PatchMatch (DPE.cpp:745, 755).  nvJPEG's luma differs from it by up to two grey levels (another inverse
// DCT); the raw-moment NCC turns one grey level into ~1e-4 of cost, and end to end that was 23 points of the
// 1-degree normal agreement with the reference (tests/test_gpu_gate2.py), so the grey images are decoded here, on the
// host threads that do the Huffman decoding for nvJPEG's hybrid backend anyway.  Only the Y component is
// reconstructed; the chroma blocks are Huffman-decoded to stay in step with the bit stream and dropped.
// Not handled (the caller falls back to nvJPEG): progressive / lossless / arithmetic-coded / 12-bit files, CMYK,
// a luma component that is not at full resolution.
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "io.h"

namespace dpe_host {
namespace {

const uint8_t kZigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22,
                             15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff {
  // canonical decoding tables: codes of length l span [mincode[l], maxcode[l]]; valptr[l] indexes `vals`
  int32_t maxcode[18];
  int32_t valoff[17];
  uint8_t vals[256];
  uint8_t look_nbits[512];  // 9-bit lookahead: code length (0 = longer than 9 bits)
  uint8_t look_sym[512];
  bool defined = false;
};

bool build_huff(const uint8_t* bits /*[17], bits[0] unused*/, const uint8_t* vals, int nvals, Huff* h) {
  int code = 0, p = 0;
  int32_t huffcode[257];
  uint8_t huffsize[257];
  for (int l = 1; l <= 16; ++l)
    for (int i = 0; i < bits[l]; ++i) { if (p >= 256) return false; huffsize[p++] = (uint8_t)l; }
  if (p != nvals) return false;
  huffsize[p] = 0;
  int si = huffsize[0];
  p = 0;
  while (huffsize[p]) {
    while (huffsize[p] == si) { huffcode[p++] = code; code++; }
    if (code > (1 << si)) return false;
    code <<= 1; si++;
  }
  p = 0;
  for (int l = 1; l <= 16; ++l) {
    if (bits[l]) {
      h->valoff[l] = p - huffcode[p];
      p += bits[l];
      h->maxcode[l] = huffcode[p - 1];
    } else {
      h->maxcode[l] = -1;
    }
  }
  h->maxcode[17] = 0xFFFFF;
  memcpy(h->vals, vals, nvals);
  memset(h->look_nbits, 0, sizeof(h->look_nbits));
  p = 0;
  for (int l = 1; l <= 9; ++l)
    for (int i = 0; i < bits[l]; ++i, ++p) {
      const int look = huffcode[p] << (9 - l);
      for (int c = 0; c < (1 << (9 - l)); ++c) { h->look_nbits[look + c] = (uint8_t)l; h->look_sym[look + c] = vals[p]; }
    }
  h->defined = true;
  return true;
}

struct BitReader {
  const uint8_t* p;
  const uint8_t* end;
  uint64_t acc = 0;
  int n = 0;          // valid bits in acc (from the top of the low n bits)
  bool marker = false;  // ran into a marker: further bits read as zeros (like libjpeg's "insufficient data" path)
  void fill() {
    while (n <= 48) {
      int c = 0;
      if (!marker && p < end) {
        c = *p++;
        if (c == 0xFF) {
          int c2 = p < end ? *p : 0xD9;
          if (c2 == 0) ++p;            // stuffed zero
          else { marker = true; --p; c = 0; }  // a marker: leave it for the caller
        }
      }
      acc = (acc << 8) | (uint64_t)c;
      n += 8;
    }
  }
  inline int peek(int k) { if (n < k) fill(); return (int)((acc >> (n - k)) & ((1u << k) - 1)); }
  inline void skip(int k) { n -= k; }
  inline int get(int k) { if (k == 0) return 0; const int v = peek(k); n -= k; return v; }
  void reset() { acc = 0; n = 0; marker = false; }
};

inline int decode_sym(BitReader& br, const Huff& h) {
  const int look = br.peek(9);
  const int nb = h.look_nbits[look];
  if (nb) { br.skip(nb); return h.look_sym[look]; }
  int l = 10;
  int32_t code = br.peek(16);
  // walk the lengths 10..16
  for (; l <= 16; ++l) {
    const int32_t c = code >> (16 - l);
    if (c <= h.maxcode[l]) { br.skip(l); return h.vals[(c + h.valoff[l]) & 255]; }
  }
  br.skip(16);
  return 0;  // corrupt data
}

inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

// jpeg_idct_islow (jidctint.c), 8x8: two passes of the LL&M algorithm in 13-bit fixed point
const int CONST_BITS = 13, PASS1_BITS = 2;
const int32_t F_0_298631336 = 2446, F_0_390180644 = 3196, F_0_541196100 = 4433, F_0_765366865 = 6270, F_0_899976223 = 7373,
              F_1_175875602 = 9633, F_1_501321110 = 12299, F_1_847759065 = 15137, F_1_961570560 = 16069, F_2_053119869 = 16819,
              F_2_562915447 = 20995, F_3_072711026 = 25172;
inline int32_t descale(int32_t x, int n) { return (x + (1 << (n - 1))) >> n; }
inline uint8_t range_limit(int32_t x) {  // sample_range_limit + CENTERJSAMPLE, indexed & RANGE_MASK (jdmaster.c)
  x = (x + 128) & 1023;
  if (x < 256) return (uint8_t)x;
  if (x < 640) return 255;      // 256..639 (post-IDCT overshoot above white)
  if (x < 896) return 0;        // 640..895
  return 0;                     // 896..1023 = -128..-1 before the centre shift: below black
}

// Coefficients of a damaged file can drive the 32-bit intermediates past INT32_MAX (libjpeg has the same arithmetic and
// the same garbage-in, garbage-out behaviour): this file is built with -fwrapv (build.py), which makes the wrap-around
// defined behaviour (found with -fsanitize=undefined on corrupted files; valid files never get near the limit).
void idct_islow(const int16_t* coef, const uint16_t* quant, uint8_t* out, int stride) {
  int32_t ws[64];
  for (int c = 0; c < 8; ++c) {
    const int16_t* in = coef + c;
    const uint16_t* q = quant + c;
    int32_t* w = ws + c;
    if (in[8] == 0 && in[16] == 0 && in[24] == 0 && in[32] == 0 && in[40] == 0 && in[48] == 0 && in[56] == 0) {
      const int32_t dc = (int32_t)in[0] * q[0] * (1 << PASS1_BITS);
      for (int r = 0; r < 8; ++r) w[8 * r] = dc;
      continue;
    }
    int32_t z2 = (int32_t)in[16] * q[16], z3 = (int32_t)in[48] * q[48];
    int32_t z1 = (z2 + z3) * F_0_541196100;
    int32_t tmp2 = z1 + z3 * (-F_1_847759065);
    int32_t tmp3 = z1 + z2 * F_0_765366865;
    z2 = (int32_t)in[0] * q[0]; z3 = (int32_t)in[32] * q[32];
    int32_t tmp0 = (z2 + z3) * (1 << CONST_BITS);
    int32_t tmp1 = (z2 - z3) * (1 << CONST_BITS);
    const int32_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = (int32_t)in[56] * q[56]; tmp1 = (int32_t)in[40] * q[40]; tmp2 = (int32_t)in[24] * q[24]; tmp3 = (int32_t)in[8] * q[8];
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int32_t z4 = tmp1 + tmp3;
    const int32_t z5 = (z3 + z4) * F_1_175875602;
    tmp0 *= F_0_298631336; tmp1 *= F_2_053119869; tmp2 *= F_3_072711026; tmp3 *= F_1_501321110;
    z1 *= -F_0_899976223; z2 *= -F_2_562915447; z3 *= -F_1_961570560; z4 *= -F_0_390180644;
    z3 += z5; z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    w[0] = descale(tmp10 + tmp3, CONST_BITS - PASS1_BITS); w[56] = descale(tmp10 - tmp3, CONST_BITS - PASS1_BITS);
    w[8] = descale(tmp11 + tmp2, CONST_BITS - PASS1_BITS); w[48] = descale(tmp11 - tmp2, CONST_BITS - PASS1_BITS);
    w[16] = descale(tmp12 + tmp1, CONST_BITS - PASS1_BITS); w[40] = descale(tmp12 - tmp1, CONST_BITS - PASS1_BITS);
    w[24] = descale(tmp13 + tmp0, CONST_BITS - PASS1_BITS); w[32] = descale(tmp13 - tmp0, CONST_BITS - PASS1_BITS);
  }
  for (int r = 0; r < 8; ++r) {
    const int32_t* w = ws + 8 * r;
    uint8_t* o = out + (size_t)r * stride;
    int32_t z2 = w[2], z3 = w[6];
    int32_t z1 = (z2 + z3) * F_0_541196100;
    int32_t tmp2 = z1 + z3 * (-F_1_847759065);
    int32_t tmp3 = z1 + z2 * F_0_765366865;
    int32_t tmp0 = (w[0] + w[4]) * (1 << CONST_BITS);
    int32_t tmp1 = (w[0] - w[4]) * (1 << CONST_BITS);
    const int32_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int32_t z4 = tmp1 + tmp3;
    const int32_t z5 = (z3 + z4) * F_1_175875602;
    tmp0 *= F_0_298631336; tmp1 *= F_2_053119869; tmp2 *= F_3_072711026; tmp3 *= F_1_501321110;
    z1 *= -F_0_899976223; z2 *= -F_2_562915447; z3 *= -F_1_961570560; z4 *= -F_0_390180644;
    z3 += z5; z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    const int S = CONST_BITS + PASS1_BITS + 3;
    o[0] = range_limit(descale(tmp10 + tmp3, S)); o[7] = range_limit(descale(tmp10 - tmp3, S));
    o[1] = range_limit(descale(tmp11 + tmp2, S)); o[6] = range_limit(descale(tmp11 - tmp2, S));
    o[2] = range_limit(descale(tmp12 + tmp1, S)); o[5] = range_limit(descale(tmp12 - tmp1, S));
    o[3] = range_limit(descale(tmp13 + tmp0, S)); o[4] = range_limit(descale(tmp13 - tmp0, S));
  }
}

struct Comp { int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0; int pred = 0; };

inline int be16(const uint8_t* p) { return (p[0] << 8) | p[1]; }

}  // namespace

bool jpeg_decode_luma_islow(const uint8_t* data, size_t size, std::vector<uint8_t>* scratch, uint8_t* dst, size_t cap, int* width,
                            int* height, std::string* err) {
  auto fail = [&](const char* m) { if (err) *err = m; return false; };
  if (size < 4 || data[0] != 0xFF || data[1] != 0xD8) return fail("not a JPEG");
  uint16_t quant[4][64];
  bool have_q[4] = {false, false, false, false};
  Huff dc[4], ac[4];
  Comp comp[4];
  int ncomp = 0, W = 0, H = 0, restart = 0;
  size_t pos = 2;
  bool sof = false;
  while (pos + 4 <= size) {
    if (data[pos] != 0xFF) return fail("marker expected");
    while (pos < size && data[pos] == 0xFF) ++pos;  // fill bytes
    if (pos >= size) break;
    const int m = data[pos++];
    if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
    if (m == 0xD9) break;
    if (pos + 2 > size) return fail("truncated");
    const int len = be16(data + pos);
    if (len < 2 || pos + len > size) return fail("bad segment length");
    const uint8_t* seg = data + pos + 2;
    const int n = len - 2;
    if (m == 0xDB) {  // DQT
      int o = 0;
      while (o < n) {
        const int pq = seg[o] >> 4, tq = seg[o] & 15;
        ++o;
        if (tq > 3 || o + 64 * (pq ? 2 : 1) > n) return fail("bad DQT");
        for (int i = 0; i < 64; ++i) {
          const int v = pq ? be16(seg + o + 2 * i) : seg[o + i];
          quant[tq][kZigzag[i]] = (uint16_t)v;
        }
        o += 64 * (pq ? 2 : 1);
        have_q[tq] = true;
      }
    } else if (m == 0xC4) {  // DHT
      int o = 0;
      while (o < n) {
        if (o + 17 > n) return fail("bad DHT");
        const int tc = seg[o] >> 4, th = seg[o] & 15;
        uint8_t bits[17];
        bits[0] = 0;
        int cnt = 0;
        for (int i = 1; i <= 16; ++i) { bits[i] = seg[o + i]; cnt += bits[i]; }
        o += 17;
        if (th > 3 || tc > 1 || cnt > 256 || o + cnt > n) return fail("bad DHT");
        if (!build_huff(bits, seg + o, cnt, tc ? &ac[th] : &dc[th])) return fail("bad Huffman table");
        o += cnt;
      }
    } else if (m == 0xC0 || m == 0xC1) {  // baseline / extended sequential, Huffman
      if (n < 6 || seg[0] != 8) return fail("only 8-bit JPEG");
      H = be16(seg + 1); W = be16(seg + 3); ncomp = seg[5];
      if (W <= 0 || H <= 0 || (ncomp != 1 && ncomp != 3) || n < 6 + 3 * ncomp) return fail("unsupported frame");
      for (int i = 0; i < ncomp; ++i) {
        comp[i].id = seg[6 + 3 * i]; comp[i].h = seg[7 + 3 * i] >> 4; comp[i].v = seg[7 + 3 * i] & 15; comp[i].tq = seg[8 + 3 * i];
        if (comp[i].h < 1 || comp[i].h > 4 || comp[i].v < 1 || comp[i].v > 4 || comp[i].tq > 3) return fail("bad component");
      }
      sof = true;
    } else if (m == 0xC2 || m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
      return fail("progressive / lossless / arithmetic JPEG");
    } else if (m == 0xDD) {
      if (n < 2) return fail("bad DRI");
      restart = be16(seg);
    } else if (m == 0xDA) {  // SOS: the one scan of a sequential file with all components
      if (!sof) return fail("scan before frame");
      if (n < 1 + 2 * ncomp + 3 || seg[0] != ncomp) return fail("unsupported scan (not all components interleaved)");
      for (int i = 0; i < ncomp; ++i) {
        const int cs = seg[1 + 2 * i];
        int k = -1;
        for (int j = 0; j < ncomp; ++j) if (comp[j].id == cs) k = j;
        if (k != i) return fail("unexpected component order");
        comp[i].td = seg[2 + 2 * i] >> 4; comp[i].ta = seg[2 + 2 * i] & 15;
        if (comp[i].td > 3 || comp[i].ta > 3 || !dc[comp[i].td].defined || !ac[comp[i].ta].defined || !have_q[comp[i].tq]) return fail("missing table");
      }
      pos += len;
      // ---- entropy-coded data
      int hmax = 1, vmax = 1;
      for (int i = 0; i < ncomp; ++i) { hmax = comp[i].h > hmax ? comp[i].h : hmax; vmax = comp[i].v > vmax ? comp[i].v : vmax; }
      if (comp[0].h != hmax || comp[0].v != vmax) return fail("luma is not at full resolution");
      if (ncomp == 1) { hmax = vmax = 1; comp[0].h = comp[0].v = 1; }  // a single-component scan is not interleaved: MCU = one block
      const int mcu_w = 8 * hmax, mcu_h = 8 * vmax;
      const int mcus_x = (W + mcu_w - 1) / mcu_w, mcus_y = (H + mcu_h - 1) / mcu_h;
      const int PW = mcus_x * mcu_w, PH = mcus_y * mcu_h;  // padded luma plane
      if ((size_t)W * H > cap) return fail("image larger than expected");
      scratch->resize((size_t)PW * PH);
      uint8_t* plane = scratch->data();
      BitReader br;
      br.p = data + pos; br.end = data + size;
      int16_t block[64];
      int to_restart = restart, next_rst = 0;
      for (int my = 0; my < mcus_y; ++my) {
        for (int mx = 0; mx < mcus_x; ++mx) {
          if (restart && to_restart == 0) {
            // byte-align, expect RSTn
            br.reset();
            const uint8_t* q = br.p;
            while (q + 1 < br.end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) ++q;
            if (q + 1 >= br.end) return fail("restart marker missing");
            (void)next_rst;
            br.p = q + 2;
            for (int i = 0; i < ncomp; ++i) comp[i].pred = 0;
            to_restart = restart;
          }
          for (int ci = 0; ci < ncomp; ++ci) {
            Comp& c = comp[ci];
            const Huff& hd = dc[c.td];
            const Huff& ha = ac[c.ta];
            for (int by = 0; by < c.v; ++by)
              for (int bx = 0; bx < c.h; ++bx) {
                // DC
                int s = decode_sym(br, hd);
                int diff = 0;
                if (s) { if (s > 15) return fail("bad DC"); diff = extend(br.get(s), s); }
                c.pred += diff;
                if (ci == 0) {
                  memset(block, 0, sizeof(block));
                  block[0] = (int16_t)c.pred;
                  for (int k = 1; k < 64;) {
                    const int rs = decode_sym(br, ha);
                    const int r = rs >> 4, sz = rs & 15;
                    if (sz == 0) { if (r == 15) { k += 16; continue; } break; }
                    k += r;
                    if (k > 63) return fail("bad AC run");
                    block[kZigzag[k]] = (int16_t)extend(br.get(sz), sz);
                    ++k;
                  }
                  idct_islow(block, quant[c.tq], plane + (size_t)(my * mcu_h + by * 8) * PW + (mx * mcu_w + bx * 8), PW);
                } else {  // chroma: keep in step with the stream, drop the coefficients
                  for (int k = 1; k < 64;) {
                    const int rs = decode_sym(br, ha);
                    const int r = rs >> 4, sz = rs & 15;
                    if (sz == 0) { if (r == 15) { k += 16; continue; } break; }
                    k += r + 1;
                    br.skip(0); br.get(sz);
                  }
                }
              }
          }
          if (restart) --to_restart;
        }
      }
      for (int y = 0; y < H; ++y) memcpy(dst + (size_t)y * W, plane + (size_t)y * PW, (size_t)W);
      *width = W; *height = H;
      return true;
    }
    pos += len;
  }
  return fail("no scan found");
}

}  // namespace dpe_host

extern "C" __attribute__((visibility("default"))) int dpe_host_decode_luma_islow(const unsigned char* data, long size, unsigned char* out,
                                                                                 long cap, int* w, int* h) {
  std::vector<uint8_t> scratch;
  std::string err;
  return dpe_host::jpeg_decode_luma_islow(data, (size_t)size, &scratch, out, (size_t)cap, w, h, &err) ? 0 : -1;
}
