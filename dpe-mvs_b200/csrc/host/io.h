// io.h — input / output contract of the pipeline (host C++).
//   pair.txt parser        = GenerateSampleList, main.cpp:264-308
//   cams/%08d_cam.txt      = ReadCamera ("TAT & ETH" 4-number variant), DPE.cpp:341-382
//   .npy writer            = WriteMatToNpy, main.cpp:48-97
//   .dmb reader / writer   = ReadBinMat / WriteBinMat, DPE.cpp:293-339
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

namespace dpe_host {

struct ProblemDesc {
  int ref_image_id;
  std::vector<int> src_image_ids;
};

struct CamFile {
  float K[9], R[9], t[3];
  float depth_min, depth_max;
};

std::string format_index(int index);  // ToFormatIndex, DPE.cpp:668-672
bool read_pairs(const std::string& path, std::vector<ProblemDesc>* problems);
bool read_cam(const std::string& path, CamFile* cam);

// descr: "<f4", "|i1", "|u1"; shape (rows, cols[, ch])
bool write_npy(const std::string& path, const void* data, const char* descr, size_t elem_size, int rows, int cols,
               int channels);

enum { DMB_8UC1 = 0, DMB_32SC1 = 4, DMB_32FC1 = 5, DMB_32FC3 = 21 };
bool read_dmb(const std::string& path, int* rows, int* cols, int* type, std::vector<uint8_t>* data);
bool write_dmb(const std::string& path, int rows, int cols, int type, const void* data, size_t bytes);
bool file_exists(const std::string& path);

// JPEG -> 8-bit luma (cv::imread(..., IMREAD_GRAYSCALE)), decoded by nvJPEG on the GPU.
struct JpegDecoder;
JpegDecoder* jpeg_decoder_create(std::string* err);
void jpeg_decoder_destroy(JpegDecoder* d);
bool jpeg_decode_gray(JpegDecoder* d, const std::string& path, std::vector<uint8_t>* gray, int* width, int* height,
                      std::string* err);
// baseline JPEG -> luma exactly as libjpeg(-turbo) / cv::imread(IMREAD_GRAYSCALE) decode it (jpeg_luma.cpp); false for
// file types it does not handle (progressive, arithmetic, 12-bit, ..): the caller falls back to nvJPEG
bool jpeg_decode_luma_islow(const uint8_t* data, size_t size, std::vector<uint8_t>* scratch, uint8_t* dst, size_t cap, int* width,
                            int* height, std::string* err);
// image size from the JPEG header alone
bool jpeg_image_size(JpegDecoder* d, const std::string& path, int* width, int* height, std::string* err);
// the same into caller-provided host memory of `cap` bytes (e.g. a slot of a pinned slab)
bool jpeg_decode_gray_into(JpegDecoder* d, const std::string& path, uint8_t* dst, size_t cap, int* width, int* height,
                           std::string* err);
// BGR interleaved (cv::IMREAD_COLOR), used by fusion only
bool jpeg_decode_bgr(JpegDecoder* d, const std::string& path, std::vector<uint8_t>* bgr, int* width, int* height,
                     std::string* err);

// JPEG encoding of an interleaved BGR image that lives on the current device (viz=True outputs), quality 95, 4:2:0
struct JpegEncoder;
JpegEncoder* jpeg_encoder_create(std::string* err);
void jpeg_encoder_destroy(JpegEncoder* e);
bool jpeg_encode_bgr_dev(JpegEncoder* e, const void* bgr_dev, int width, int height, const std::string& path, std::string* err);
// the same for an image in host memory (staged through a temporary device buffer on the current device)
bool jpeg_encode_bgr_host(JpegEncoder* e, const uint8_t* bgr, int width, int height, const std::string& path, std::string* err);

}  // namespace dpe_host
