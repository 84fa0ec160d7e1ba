// io.cpp — see io.h.  Host C++ only (compiled by nvcc's host compiler; nvJPEG for decode).
#include "io.h"

#include <cuda_runtime.h>
#include <nvjpeg.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <fstream>
#include <iomanip>
#include <sstream>

namespace dpe_host {

std::string format_index(int index) {
  std::stringstream ss;
  ss << std::setw(8) << std::setfill('0') << index;
  return ss.str();
}

bool file_exists(const std::string& path) {
  struct stat st;
  return stat(path.c_str(), &st) == 0;
}

// pair.txt: line 1 = V; per view: line "ref_id", line "n id1 s1 ... idn sn"; sources with a
// score <= 0 are dropped (main.cpp:297-305).
bool read_pairs(const std::string& path, std::vector<ProblemDesc>* problems) {
  problems->clear();
  std::ifstream file(path);
  if (!file.good()) return false;
  std::string line;
  std::stringstream iss;
  int num_images = 0;
  std::getline(file, line);
  iss.str(line);
  // A count that does not parse (or overflows: the stream then stores INT_MAX) must not become a loop bound — the
  // reference's GenerateSampleList would spin through 2^31 failed extractions on such a file; here it is an error.
  if (!(iss >> num_images) || num_images < 0) return false;
  for (int i = 0; i < num_images; ++i) {
    ProblemDesc p;
    p.ref_image_id = 0;
    iss.clear();
    if (!std::getline(file, line)) return false;  // fewer problems than announced
    iss.str(line);
    if (!(iss >> p.ref_image_id)) return false;
    int n = 0;
    iss.clear();
    if (!std::getline(file, line)) return false;
    iss.str(line);
    if (!(iss >> n)) return false;
    for (int j = 0; j < n; ++j) {
      int id = 0;
      float score = 0.f;
      if (!(iss >> id >> score)) break;  // fewer sources on the line than announced: keep what is there
      if (score <= 0.0f) continue;
      p.src_image_ids.push_back(id);
    }
    problems->push_back(p);
  }
  return true;
}

bool read_cam(const std::string& path, CamFile* cam) {
  std::ifstream in(path);
  if (!in.good()) return false;
  std::string tok;
  in >> tok;  // "extrinsic"
  for (int i = 0; i < 3; ++i) in >> cam->R[3 * i + 0] >> cam->R[3 * i + 1] >> cam->R[3 * i + 2] >> cam->t[i];
  float tmp[4];
  in >> tmp[0] >> tmp[1] >> tmp[2] >> tmp[3];
  in >> tok;  // "intrinsic"
  for (int i = 0; i < 3; ++i) in >> cam->K[3 * i + 0] >> cam->K[3 * i + 1] >> cam->K[3 * i + 2];
  float interval = 0.f, depth_num = 0.f;
  in >> cam->depth_min >> interval >> depth_num >> cam->depth_max;
  return !in.fail();
}

// NPY v1.0; header dict exactly as the reference writes it, space padded so that
// 10 + header length is a multiple of 16 (main.cpp:71-85).
bool write_npy(const std::string& path, const void* data, const char* descr, size_t elem_size, int rows, int cols,
               int channels) {
  std::ofstream out(path, std::ios::binary);
  if (!out.is_open()) return false;
  std::ostringstream shape;
  shape << "(" << rows << ", " << cols;
  if (channels > 1) shape << ", " << channels;
  shape << ")";
  std::string header = std::string("{'descr': '") + descr + "', 'fortran_order': False, 'shape': " + shape.str() + ", }";
  const size_t header_len = header.size() + 1;
  const size_t padding = (16 - ((10 + header_len) % 16)) % 16;
  header.append(padding, ' ');
  header.push_back('\n');
  const char magic[] = "\x93NUMPY";
  out.write(magic, 6);
  out.put((char)1);
  out.put((char)0);
  const uint16_t hs = (uint16_t)header.size();
  out.write((const char*)&hs, 2);
  out.write(header.data(), header.size());
  out.write((const char*)data, (size_t)rows * cols * channels * elem_size);
  return out.good();
}

bool read_dmb(const std::string& path, int* rows, int* cols, int* type, std::vector<uint8_t>* data) {
  std::ifstream in(path, std::ios::binary);
  if (!in.good()) return false;
  int32_t h[4] = {0, 0, 0, 0};
  in.read((char*)h, 16);
  if (in.fail() || h[0] != 1 || h[1] <= 0 || h[2] <= 0) return false;
  size_t es;
  switch (h[3]) {
    case DMB_8UC1: es = 1; break;
    case DMB_32SC1: case DMB_32FC1: es = 4; break;
    case DMB_32FC3: es = 12; break;
    default: return false;
  }
  // the payload must be in the file before memory is set aside for it: a foreign header can announce 2^62 elements
  const size_t want = (size_t)h[1] * (size_t)h[2] * es;
  in.seekg(0, std::ios::end);
  const std::streamoff file_bytes = in.tellg();
  if (file_bytes < 16 || (size_t)(file_bytes - 16) < want) return false;
  in.seekg(16, std::ios::beg);
  *rows = h[1]; *cols = h[2]; *type = h[3];
  data->resize(want);
  in.read((char*)data->data(), data->size());
  return !in.fail();
}

bool write_dmb(const std::string& path, int rows, int cols, int type, const void* data, size_t bytes) {
  std::ofstream out(path, std::ios::binary);
  if (!out.is_open()) return false;
  const int32_t h[4] = {1, rows, cols, type};
  out.write((const char*)h, 16);
  out.write((const char*)data, bytes);
  return out.good();
}

// ---- nvJPEG ---------------------------------------------------------------------------------
struct JpegDecoder {
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
  cudaStream_t stream = nullptr;
  unsigned char* dev = nullptr;
  size_t dev_bytes = 0;
  std::vector<unsigned char> file;
  std::vector<uint8_t> scratch;  // padded luma plane of the host decoder
};

// The nvJPEG handle (tens to hundreds of milliseconds to create, and a CUDA context on the calling thread) comes up on
// first use: grey decodes of baseline files never need it (jpeg_luma.cpp), only colour, progressive files and
// DPE_JPEG_DECODER=nvjpeg do.
JpegDecoder* jpeg_decoder_create(std::string* err) {
  (void)err;
  return new JpegDecoder();
}
static bool ensure_nvjpeg(JpegDecoder* d, std::string* err) {
  if (d->handle) return true;
  if (nvjpegCreateSimple(&d->handle) != NVJPEG_STATUS_SUCCESS || nvjpegJpegStateCreate(d->handle, &d->state) != NVJPEG_STATUS_SUCCESS) {
    if (d->handle) { nvjpegDestroy(d->handle); d->handle = nullptr; }
    if (err) *err = "nvjpeg initialisation failed";
    return false;
  }
  cudaStreamCreate(&d->stream);
  return true;
}

// frame size from the SOFn marker of a JPEG stream (ITU T.81 B.2.2); false if none is found in the bytes given
static bool jpeg_frame_size(const unsigned char* p, size_t n, int* width, int* height) {
  if (n < 4 || p[0] != 0xFF || p[1] != 0xD8) return false;
  size_t i = 2;
  while (i + 4 <= n) {
    if (p[i] != 0xFF) { ++i; continue; }
    const unsigned char m = p[i + 1];
    if (m == 0xFF) { ++i; continue; }                                     // fill byte
    if (m == 0x01 || (m >= 0xD0 && m <= 0xD8)) { i += 2; continue; }      // TEM, RSTn, SOI: no length
    if (m == 0xD9 || m == 0xDA) return false;                             // EOI / start of scan before any frame header
    const size_t len = ((size_t)p[i + 2] << 8) | p[i + 3];
    if (len < 2) return false;
    if (m >= 0xC0 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
      if (i + 9 > n) return false;
      *height = (p[i + 5] << 8) | p[i + 6];
      *width = (p[i + 7] << 8) | p[i + 8];
      return *width > 0 && *height > 0;
    }
    i += 2 + len;
  }
  return false;
}

void jpeg_decoder_destroy(JpegDecoder* d) {
  if (!d) return;
  if (d->state) nvjpegJpegStateDestroy(d->state);
  if (d->handle) nvjpegDestroy(d->handle);
  if (d->stream) cudaStreamDestroy(d->stream);
  cudaFree(d->dev);
  delete d;
}

// decodes into `dst` (host memory of `cap` bytes; pinned memory makes the copy back asynchronous to other streams)
static bool jpeg_decode_to(JpegDecoder* d, const std::string& path, nvjpegOutputFormat_t fmt, int ch, uint8_t* dst, size_t cap,
                           std::vector<uint8_t>* grow, int* width, int* height, std::string* err) {
  std::ifstream in(path, std::ios::binary | std::ios::ate);
  if (!in.good()) { if (err) *err = "cannot open " + path; return false; }
  const std::streamsize n = in.tellg();
  in.seekg(0);
  d->file.resize((size_t)n);
  in.read((char*)d->file.data(), n);
  int w = 0, h = 0;
  if (!jpeg_frame_size(d->file.data(), d->file.size(), &w, &h)) {
    if (err) *err = "not a decodable JPEG: " + path;
    return false;
  }
  const size_t need = (size_t)w * h * ch;
  if (fmt == NVJPEG_OUTPUT_Y) {
    // grey images: the host decoder that reproduces libjpeg's pixels (jpeg_luma.cpp) unless DPE_JPEG_DECODER=nvjpeg;
    // files it does not handle (progressive, ..) go through nvJPEG below
    const char* force = getenv("DPE_JPEG_DECODER");
    if (!(force && std::string(force) == "nvjpeg")) {
      uint8_t* out = dst;
      size_t out_cap = cap;
      if (grow) { grow->resize(need); out = grow->data(); out_cap = need; }
      int ew = 0, eh = 0;
      std::string e2;
      if (need <= out_cap && jpeg_decode_luma_islow(d->file.data(), d->file.size(), &d->scratch, out, out_cap, &ew, &eh, &e2) && ew == w && eh == h) {
        *width = w; *height = h;
        return true;
      }
    }
  }
  if (!ensure_nvjpeg(d, err)) return false;
  {
    int comps = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
    nvjpegChromaSubsampling_t ss;
    if (nvjpegGetImageInfo(d->handle, d->file.data(), d->file.size(), &comps, &ss, ws, hs) != NVJPEG_STATUS_SUCCESS || ws[0] != w || hs[0] != h) {
      if (err) *err = "not a decodable JPEG: " + path;
      return false;
    }
  }
  if (need > d->dev_bytes) {
    cudaFree(d->dev);
    if (cudaMalloc(&d->dev, need) != cudaSuccess) { if (err) *err = "cudaMalloc failed in jpeg decode"; d->dev_bytes = 0; return false; }
    d->dev_bytes = need;
  }
  nvjpegImage_t img;
  memset(&img, 0, sizeof(img));
  img.channel[0] = d->dev;
  img.pitch[0] = (size_t)w * ch;
  if (nvjpegDecode(d->handle, d->state, d->file.data(), d->file.size(), fmt, &img, d->stream) != NVJPEG_STATUS_SUCCESS) {
    if (err) *err = "nvjpegDecode failed: " + path;
    return false;
  }
  if (grow) { grow->resize(need); dst = grow->data(); cap = need; }
  if (need > cap) { if (err) *err = "image larger than expected: " + path; return false; }
  if (cudaMemcpyAsync(dst, d->dev, need, cudaMemcpyDeviceToHost, d->stream) != cudaSuccess ||
      cudaStreamSynchronize(d->stream) != cudaSuccess) {
    if (err) *err = "copy after jpeg decode failed";
    return false;
  }
  *width = w; *height = h;
  return true;
}

static bool jpeg_decode(JpegDecoder* d, const std::string& path, nvjpegOutputFormat_t fmt, int ch, std::vector<uint8_t>* out,
                        int* width, int* height, std::string* err) {
  return jpeg_decode_to(d, path, fmt, ch, nullptr, 0, out, width, height, err);
}

bool jpeg_image_size(JpegDecoder* d, const std::string& path, int* width, int* height, std::string* err) {
  std::ifstream in(path, std::ios::binary);
  if (!in.good()) { if (err) *err = "cannot open " + path; return false; }
  // the frame header sits in the first kilobytes
  std::vector<unsigned char> head(65536);
  in.read((char*)head.data(), (std::streamsize)head.size());
  const size_t n = (size_t)in.gcount();
  (void)d;
  if (!jpeg_frame_size(head.data(), n, width, height)) {
    if (err) *err = "not a decodable JPEG: " + path;
    return false;
  }
  return true;
}

bool jpeg_decode_gray_into(JpegDecoder* d, const std::string& path, uint8_t* dst, size_t cap, int* width, int* height, std::string* err) {
  return jpeg_decode_to(d, path, NVJPEG_OUTPUT_Y, 1, dst, cap, nullptr, width, height, err);
}

bool jpeg_decode_gray(JpegDecoder* d, const std::string& path, std::vector<uint8_t>* gray, int* width, int* height, std::string* err) {
  return jpeg_decode(d, path, NVJPEG_OUTPUT_Y, 1, gray, width, height, err);
}
bool jpeg_decode_bgr(JpegDecoder* d, const std::string& path, std::vector<uint8_t>* bgr, int* width, int* height, std::string* err) {
  return jpeg_decode(d, path, NVJPEG_OUTPUT_BGRI, 3, bgr, width, height, err);
}

// ---- nvJPEG encoder (viz=True images) ---------------------------------------------------------
struct JpegEncoder {
  nvjpegHandle_t handle = nullptr;
  nvjpegEncoderState_t state = nullptr;
  nvjpegEncoderParams_t params = nullptr;
  cudaStream_t stream = nullptr;
};
JpegEncoder* jpeg_encoder_create(std::string* err) {
  JpegEncoder* e = new JpegEncoder();
  cudaStreamCreate(&e->stream);
  if (nvjpegCreateSimple(&e->handle) != NVJPEG_STATUS_SUCCESS || nvjpegEncoderStateCreate(e->handle, &e->state, e->stream) != NVJPEG_STATUS_SUCCESS ||
      nvjpegEncoderParamsCreate(e->handle, &e->params, e->stream) != NVJPEG_STATUS_SUCCESS) {
    if (err) *err = "nvjpeg encoder initialisation failed";
    jpeg_encoder_destroy(e);
    return nullptr;
  }
  nvjpegEncoderParamsSetQuality(e->params, 95, e->stream);                       // cv::imwrite's default
  nvjpegEncoderParamsSetSamplingFactors(e->params, NVJPEG_CSS_420, e->stream);  // libjpeg's default for YCbCr
  return e;
}
void jpeg_encoder_destroy(JpegEncoder* e) {
  if (!e) return;
  if (e->params) nvjpegEncoderParamsDestroy(e->params);
  if (e->state) nvjpegEncoderStateDestroy(e->state);
  if (e->handle) nvjpegDestroy(e->handle);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}
bool jpeg_encode_bgr_dev(JpegEncoder* e, const void* bgr_dev, int width, int height, const std::string& path, std::string* err) {
  nvjpegImage_t img;
  memset(&img, 0, sizeof(img));
  img.channel[0] = (unsigned char*)bgr_dev;
  img.pitch[0] = (size_t)width * 3;
  if (nvjpegEncodeImage(e->handle, e->state, e->params, &img, NVJPEG_INPUT_BGRI, width, height, e->stream) != NVJPEG_STATUS_SUCCESS) {
    if (err) *err = "nvjpegEncodeImage failed";
    return false;
  }
  size_t len = 0;
  if (nvjpegEncodeRetrieveBitstream(e->handle, e->state, nullptr, &len, e->stream) != NVJPEG_STATUS_SUCCESS) { if (err) *err = "nvjpeg bitstream size"; return false; }
  std::vector<unsigned char> buf(len);
  if (nvjpegEncodeRetrieveBitstream(e->handle, e->state, buf.data(), &len, e->stream) != NVJPEG_STATUS_SUCCESS ||
      cudaStreamSynchronize(e->stream) != cudaSuccess) { if (err) *err = "nvjpeg bitstream"; return false; }
  std::ofstream out(path, std::ios::binary);
  if (!out.is_open()) { if (err) *err = "cannot write " + path; return false; }
  out.write((const char*)buf.data(), (std::streamsize)len);
  return out.good();
}

bool jpeg_encode_bgr_host(JpegEncoder* e, const uint8_t* bgr, int width, int height, const std::string& path, std::string* err) {
  void* dev = nullptr;
  const size_t bytes = (size_t)width * height * 3;
  if (cudaMalloc(&dev, bytes) != cudaSuccess) { cudaGetLastError(); if (err) *err = "cudaMalloc for a viz image failed"; return false; }
  bool ok = cudaMemcpy(dev, bgr, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
  if (!ok && err) *err = "upload of a viz image failed";
  ok = ok && jpeg_encode_bgr_dev(e, dev, width, height, path, err);
  cudaFree(dev);
  return ok;
}

}  // namespace dpe_host

// C hooks for the CPU tests (tests/test_io.py)
extern "C" {
#define DPE_TEST_API __attribute__((visibility("default")))
DPE_TEST_API int dpe_host_write_npy(const char* path, const void* data, const char* descr, int elem_size, int rows, int cols,
                                    int channels) {
  return dpe_host::write_npy(path, data, descr, (size_t)elem_size, rows, cols, channels) ? 0 : -1;
}
// out: K[9] R[9] t[3] depth_min depth_max (23 floats)
DPE_TEST_API int dpe_host_read_cam(const char* path, float* out) {
  dpe_host::CamFile c;
  if (!dpe_host::read_cam(path, &c)) return -1;
  memcpy(out, c.K, 36); memcpy(out + 9, c.R, 36); memcpy(out + 18, c.t, 12);
  out[21] = c.depth_min; out[22] = c.depth_max;
  return 0;
}
// out: per problem: ref_id, n_src, src ids...; returns number of ints written (or -1)
DPE_TEST_API int dpe_host_read_pairs(const char* path, int* out, int cap) {
  std::vector<dpe_host::ProblemDesc> p;
  if (!dpe_host::read_pairs(path, &p)) return -1;
  int n = 0;
  for (const auto& q : p) {
    if (n + 2 + (int)q.src_image_ids.size() > cap) return -1;
    out[n++] = q.ref_image_id; out[n++] = (int)q.src_image_ids.size();
    for (int s : q.src_image_ids) out[n++] = s;
  }
  return n;
}
}

// .dmb reader: rows / cols / type / payload bytes of a file, -1 when it is refused
extern "C" DPE_TEST_API long dpe_host_read_dmb(const char* path, int* rows, int* cols, int* type) {
  std::vector<uint8_t> d;
  if (!dpe_host::read_dmb(path, rows, cols, type, &d)) return -1;
  return (long)d.size();
}
extern "C" DPE_TEST_API int dpe_host_jpeg_size(const char* path, int* w, int* h) {
  std::string err;
  dpe_host::JpegDecoder* d = dpe_host::jpeg_decoder_create(&err);
  if (!d) return -1;
  const bool ok = dpe_host::jpeg_image_size(d, path, w, h, &err);
  dpe_host::jpeg_decoder_destroy(d);
  return ok ? 0 : -2;
}

extern "C" DPE_TEST_API int dpe_host_decode_gray(const char* path, unsigned char* out, int cap, int* w, int* h) {
  std::string err;
  dpe_host::JpegDecoder* d = dpe_host::jpeg_decoder_create(&err);
  if (!d) return -1;
  std::vector<uint8_t> g;
  const bool ok = dpe_host::jpeg_decode_gray(d, path, &g, w, h, &err);
  dpe_host::jpeg_decoder_destroy(d);
  if (!ok || (int)g.size() > cap) return -2;
  memcpy(out, g.data(), g.size());
  return 0;
}
