// pipeline.cpp — C++ host orchestrator: dpe_run_pipeline(), the drop-in for
// RunDPEPipeline (main.cpp:474-600) + ProcessProblem (main.cpp:411-472).  It talks to the
// CUDA side only through the C ABI (include/dpe_b200.h).
//
// Differences from the reference that are deliberate:
//  * the scene is loaded once (one JPEG decode per image instead of (N+1) x 4R decodes) and
//    stays in HBM; per-view state never touches the disk between stages (the reference
//    round-trips depths.dmb / normals.dmb / weak.bin / selected_views.bin per view-stage);
//  * gpu_index >= 0 selects that GPU; gpu_index < 0 (or env DPE_GPUS=0,1,..) shards the
//    reference views over several GPUs of the box: one worker thread and one context per GPU, the
//    images uploaded to the first GPU and broadcast with NCCL, the depth atlas all-gathered with NCCL
//    behind every stage (include/dpe_b200.h, "several GPUs").  With several GPUs all views of a
//    stage read the previous stage's depth maps (Jacobi) instead of a mix of this stage's and the
//    previous stage's (Gauss-Seidel through files);
//  * the host work around the path overlaps it: contexts and the communicator come up while the
//    JPEGs decode, edge / label preparation (needed from the first fine stage on) runs on the host
//    cores under the coarse stages, and every view's .npy files are written while the views behind
//    it are still in their last stage;
//  * errors are returned (never exit()).
// Kept: the verbose lines, the stage schedule and parameters, the edges_k/labels_k .dmb
// cache (used if present and of the right size, deleted at the end), the .npy outputs, DPE.ply
// when fusion is on.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <functional>
#include <iostream>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/dpe_b200.h"
#include "fusion.h"
#include "io.h"
#include "prep.h"

using namespace dpe_host;

namespace {

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int compute_round_num(int width, int height) {  // main.cpp:390-408
  int max_size = width > height ? width : height;
  int round_num = 1;
  while (max_size > 800) { max_size /= 2; round_num++; }
  return round_num > 2 ? round_num : 2;
}

struct ViewPrep {
  std::vector<ImageU8> edge;                 // per file index j (0 = full resolution)
  std::vector<std::vector<int32_t>> label;
};

void parallel_for(int n, int n_threads, const std::function<void(int)>& fn) {
  std::atomic<int> next(0);
  std::vector<std::thread> th;
  n_threads = std::max(1, std::min(n_threads, n));
  for (int t = 0; t < n_threads; ++t)
    th.emplace_back([&]() { for (int i = next++; i < n; i = next++) fn(i); });
  for (auto& t : th) t.join();
}

struct Timing {
  double load = 0, prep = 0, upload = 0, stages = 0, output = 0, fusion = 0, total = 0, gpu_ms = 0;
  double ctx = 0, comm = 0, prep_wait = 0;
  long long launches = 0;
};

// the stage schedule of RunDPEPipeline (main.cpp:507-567)
struct Stage { int scale; dpe_stage_params p; bool resolution_up; };
std::vector<Stage> make_schedule(int round_num) {
  std::vector<Stage> out;
  for (int i = 0; i < round_num; ++i) {
    dpe_stage_params p;
    memset(&p, 0, sizeof(p));
    p.max_iterations = 3; p.top_k = 4; p.geom_factor = 0.2f;
    p.ransac_threshold = 0.005f; p.rotate_time = 4;
    if (i == 0) { p.state = DPE_FIRST_INIT; p.use_apd = 0; }
    else {
      p.state = DPE_REFINE_INIT; p.use_apd = 1;
      p.ransac_threshold = (float)(0.01 - i * 0.00125);
      p.rotate_time = std::min((int)(1 << i), 4);
    }
    p.geom_consistency = 0; p.weak_peak_radius = 6;
    out.push_back({i, p, false});
    for (int j = 0; j < 3; ++j) {
      p.state = DPE_REFINE_ITER; p.use_apd = i == 0 ? 0 : 1;
      p.ransac_threshold = (float)(0.01 - i * 0.00125);
      p.rotate_time = std::min((int)(1 << i), 4);
      p.geom_consistency = 1; p.weak_peak_radius = std::max(4 - 2 * j, 2);
      out.push_back({i, p, j == 2});
    }
  }
  return out;
}

// pinned host buffers for the export of one view, recycled through a small pool
struct ExportBuf {
  float* depth = nullptr; float* normal = nullptr; int8_t* weak = nullptr;
  int view = -1;
};

// bounded pool + queue between the GPU workers (producers) and the .npy writer threads
class ExportQueue {
 public:
  ExportQueue(size_t P, bool want_normal, bool want_weak, int n_bufs) {
    for (int i = 0; i < n_bufs; ++i) {
      ExportBuf* b = new ExportBuf();
      bool ok = cudaHostAlloc((void**)&b->depth, P * sizeof(float), cudaHostAllocPortable) == cudaSuccess;
      if (want_normal) ok = ok && cudaHostAlloc((void**)&b->normal, P * 3 * sizeof(float), cudaHostAllocPortable) == cudaSuccess;
      if (want_weak) ok = ok && cudaHostAlloc((void**)&b->weak, P, cudaHostAllocPortable) == cudaSuccess;
      if (!ok) { release(b); continue; }
      all_.push_back(b); free_.push_back(b);
    }
  }
  ~ExportQueue() { for (auto* b : all_) release(b); }
  bool usable() const { return !all_.empty(); }
  ExportBuf* acquire() {
    std::unique_lock<std::mutex> l(m_);
    cv_.wait(l, [&]() { return !free_.empty(); });
    ExportBuf* b = free_.back(); free_.pop_back();
    return b;
  }
  void submit(ExportBuf* b) { { std::lock_guard<std::mutex> l(m_); ready_.push_back(b); } cv_.notify_all(); }
  void recycle(ExportBuf* b) { { std::lock_guard<std::mutex> l(m_); free_.push_back(b); } cv_.notify_all(); }
  // nullptr once close() was called and the queue is drained
  ExportBuf* take() {
    std::unique_lock<std::mutex> l(m_);
    cv_.wait(l, [&]() { return !ready_.empty() || closed_; });
    if (ready_.empty()) return nullptr;
    ExportBuf* b = ready_.front(); ready_.pop_front();
    return b;
  }
  void close() { { std::lock_guard<std::mutex> l(m_); closed_ = true; } cv_.notify_all(); }

 private:
  static void release(ExportBuf* b) {
    if (b->depth) cudaFreeHost(b->depth);
    if (b->normal) cudaFreeHost(b->normal);
    if (b->weak) cudaFreeHost(b->weak);
    delete b;
  }
  std::mutex m_;
  std::condition_variable cv_;
  std::vector<ExportBuf*> all_, free_;
  std::deque<ExportBuf*> ready_;
  bool closed_ = false;
};

}  // namespace

extern "C" DPE_API int dpe_run_pipeline(const char* dense_folder_c, int gpu_index, int verbose, int fusion, int viz,
                                        int depth, int normal, int weak, int edge) {
  // viz: depth_<i>.jpg / normal_<i>.jpg / weak_<i>.jpg per view and iteration like the reference
  // (main.cpp:448-454) and the rawedge_<k>.jpg / connect_<k>.jpg images of the prep stage (main.cpp:361-364, 380-383);
  // the complex.jpg residue (Q12) is not written
  const double t_begin = now_s();
  Timing tm;
  const std::string dense = dense_folder_c ? dense_folder_c : "";
  const std::string out_root = dense + "/DPE";
  mkdir(out_root.c_str(), 0777);

  // RunFusion gates reference pixels with <dense>/blocks/mask_<id>.jpg when that folder exists (DPE.cpp:1242-1268, 1296)
  const bool use_block = fusion && file_exists(dense + "/blocks");

  // ---- GenerateSampleList --------------------------------------------------------------------
  std::vector<ProblemDesc> problems;
  if (!read_pairs(dense + "/pair.txt", &problems) || problems.empty()) {
    std::cerr << "Images may error, check it!\n";  // CheckImages fails on an empty problem list
    return 1;
  }
  for (const auto& p : problems) mkdir((out_root + "/" + format_index(p.ref_image_id)).c_str(), 0777);
  const int n_problems = (int)problems.size();
  // scene views: the problems in pair.txt order, then images that only appear as sources
  std::map<int, int> id_to_view;
  std::vector<int> view_ids;
  for (const auto& p : problems)
    if (!id_to_view.count(p.ref_image_id)) { id_to_view[p.ref_image_id] = (int)view_ids.size(); view_ids.push_back(p.ref_image_id); }
  if ((int)view_ids.size() != n_problems) { std::cerr << "pair.txt lists a reference image twice\n"; return 1; }
  for (const auto& p : problems)
    for (int s : p.src_image_ids)
      if (!id_to_view.count(s)) { id_to_view[s] = (int)view_ids.size(); view_ids.push_back(s); }
  for (const auto& p : problems)
    if ((int)p.src_image_ids.size() + 1 > DPE_MAX_IMAGES) {
      std::cerr << "Can't process so much images: " << p.src_image_ids.size() + 1 << std::endl;  // DPE.cpp:762-765
      return 1;
    }
  const int n_views = (int)view_ids.size();

  // ---- devices --------------------------------------------------------------------------------
  std::vector<int> gpus;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess) n_dev = 0;
  if (const char* env = getenv("DPE_GPUS")) {
    std::string s(env);
    size_t pos = 0;
    while (pos < s.size()) {
      size_t c = s.find(',', pos);
      if (c == std::string::npos) c = s.size();
      if (c > pos) gpus.push_back(atoi(s.substr(pos, c - pos).c_str()));
      pos = c + 1;
    }
    std::set<int> seen;
    for (int g : gpus)
      if (g < 0 || g >= n_dev || !seen.insert(g).second) {
        std::cerr << "DPE-MVS: DPE_GPUS names GPU " << g << " twice or out of range (" << n_dev << " visible)\n";
        return 1;
      }
  }
  if (gpus.empty()) {
    if (gpu_index >= 0) gpus.push_back(gpu_index);
    else for (int i = 0; i < n_dev; ++i) gpus.push_back(i);
  }
  if (gpus.empty() || n_dev == 0) { std::cerr << "DPE-MVS: no CUDA device\n"; return 1; }
  if ((int)gpus.size() > n_problems) gpus.resize(n_problems);
  const int G = (int)gpus.size();
  if (cudaSetDevice(gpus[0]) != cudaSuccess) { std::cerr << "DPE-MVS: cannot select GPU " << gpus[0] << "\n"; return 1; }

  // ---- contexts come up on their own threads while the images decode ----------------------------------
  std::vector<dpe_ctx*> ctxs(G, nullptr);
  std::vector<std::string> errs(G);
  std::vector<std::thread> ctx_threads;
  const double t_ctx0 = now_s();
  for (int g = 0; g < G; ++g) ctx_threads.emplace_back([&, g]() {
    if (dpe_ctx_create(&ctxs[g], gpus[g]) != DPE_OK) errs[g] = "cannot create a context on GPU " + std::to_string(gpus[g]);
  });
  auto destroy_all = [&]() {
    for (auto*& x : ctxs) { dpe_ctx_destroy(x); x = nullptr; }
  };

  // ---- images + cameras (CheckImages: all the same size) --------------------------------------
  double t0 = now_s();
  std::string err;
  JpegDecoder* dec = jpeg_decoder_create(&err);
  uint8_t* gray_slab = nullptr;  // pinned: all images, view-major (uploads overlap, prep reads it)
  auto bail = [&](const std::string& msg) {
    for (auto& t : ctx_threads) if (t.joinable()) t.join();
    destroy_all();
    jpeg_decoder_destroy(dec);
    if (gray_slab) cudaFreeHost(gray_slab);
    std::cerr << msg << "\n";
    return 1;
  };
  if (!dec) return bail("DPE-MVS: " + err);
  std::vector<CamFile> cams(n_views);
  int width = 0, height = 0;
  auto image_path = [&](int v) { return dense + "/images/" + format_index(view_ids[v]) + ".jpg"; };
  size_t P = 0;
  {
    // the header of the first image gives the size; then all images decode on a few threads, one nvJPEG decoder each
    // (Huffman decoding is host work, the rest runs on gpus[0]), straight into the pinned slab
    if (!jpeg_image_size(dec, image_path(0), &width, &height, &err) || width <= 0 || height <= 0) return bail("Images may error, check it!");
    P = (size_t)width * height;
    if (cudaHostAlloc((void**)&gray_slab, P * n_views, cudaHostAllocPortable) != cudaSuccess) return bail("DPE-MVS: cannot allocate pinned memory");
    const int n_load = std::max(1, std::min({(int)std::thread::hardware_concurrency(), 16, n_views}));
    std::vector<JpegDecoder*> decs(n_load, nullptr);
    decs[0] = dec;
    std::atomic<int> bad(0), bad_cam(0), next_dec(0);
    parallel_for(n_views, n_load, [&](int v) {
      thread_local int my = -1;
      thread_local const void* owner = nullptr;
      if (my < 0 || owner != (const void*)&decs) {
        my = next_dec++; owner = (const void*)&decs;
        cudaSetDevice(gpus[0]);
        if (my > 0) { std::string e; decs[my] = jpeg_decoder_create(&e); }
      }
      if (!read_cam(dense + "/cams/" + format_index(view_ids[v]) + "_cam.txt", &cams[v])) bad_cam++;
      std::string e;
      int w = 0, h = 0;
      if (!decs[my] || !jpeg_decode_gray_into(decs[my], image_path(v), gray_slab + P * v, P, &w, &h, &e) || w != width || h != height) bad++;
    });
    for (int i = 1; i < n_load; ++i) jpeg_decoder_destroy(decs[i]);
    if (bad.load()) return bail("Images may error, check it!");  // CheckImages (main.cpp:310-329)
    if (bad_cam.load()) return bail("DPE-MVS: cannot read a camera file");
  }
  tm.load = now_s() - t0;
  if (verbose) std::cout << "There are " << n_problems << " images to be processed!" << std::endl;
  const int round_num = compute_round_num(width, height);
  const int iteration_num = round_num * 4;

  // ---- edge / label preparation (GetProblemEdges) in the background: the first fine stage needs it, the
  //      coarse stages do not.  Cached .dmb files are honoured (main.cpp:351-355, 370-374).
  std::vector<ViewPrep> prep(n_problems);
  for (int v = 0; v < n_problems; ++v) { prep[v].edge.resize(round_num); prep[v].label.resize(round_num); }
  std::mutex prep_m;
  std::condition_variable prep_cv;
  std::vector<int> prep_left(n_problems, round_num);  // jobs outstanding per view
  const double t_prep0 = now_s();
  std::atomic<double> prep_done_at(0.0);
  std::thread prep_thread([&]() {
    const int hw = (int)std::thread::hardware_concurrency();
    // full-resolution jobs first (they take longest), views in order
    std::vector<std::pair<int, int>> jobs;
    for (int j = 0; j < round_num; ++j)
      for (int v = 0; v < n_problems; ++v) jobs.push_back({v, j});
    const int n_prep_threads = std::max(1, (hw > 0 ? hw : 4) - 1);
    std::vector<JpegEncoder*> encs(viz ? n_prep_threads : 0, nullptr);  // viz: one nvJPEG encoder per prep thread
    std::atomic<int> next_enc(0);
    parallel_for((int)jobs.size(), n_prep_threads, [&](int job) {
      const int v = jobs[job].first, j = jobs[job].second;
      const std::string dir = out_root + "/" + format_index(view_ids[v]);
      const std::string ep = dir + "/edges_" + std::to_string(j) + ".dmb";
      const std::string lp = dir + "/labels_" + std::to_string(j) + ".dmb";
      bool have_e = false, have_l = false;
      int r, c, t;
      std::vector<uint8_t> buf;
      // a cached file is honoured only if it has the size this scale needs (round(W/2^j) x round(H/2^j), the
      // reference's cv::resize target, main.cpp:343-345) and carries that many elements; a stale or foreign
      // file is ignored and the arrays recomputed
      const int want_c = (int)std::round(width * (1.0f / (float)(1 << j))), want_r = (int)std::round(height * (1.0f / (float)(1 << j)));
      if (file_exists(ep) && read_dmb(ep, &r, &c, &t, &buf) && t == DMB_8UC1 && r == want_r && c == want_c &&
          buf.size() == (size_t)r * c) {
        prep[v].edge[j].rows = r; prep[v].edge[j].cols = c; prep[v].edge[j].d = buf;
        have_e = true;
      }
      if (file_exists(lp) && read_dmb(lp, &r, &c, &t, &buf) && t == DMB_32SC1 && r == want_r && c == want_c &&
          buf.size() == (size_t)r * c * sizeof(int32_t)) {
        prep[v].label[j].resize((size_t)r * c);
        memcpy(prep[v].label[j].data(), buf.data(), buf.size());
        have_l = true;
      }
      if (!have_e || !have_l) {
        ImageU8 gimg;
        gimg.rows = height; gimg.cols = width;
        gimg.d.assign(gray_slab + P * v, gray_slab + P * (v + 1));
        ImageU8 e;
        std::vector<int32_t> l;
        int oc, orr;
        problem_edges(gimg, 1 << j, &e, &l, &oc, &orr);
        if (viz) {
          // show_medium_result: the edge map as rawedge_<j>.jpg and the connected regions as connect_<j>.jpg, each only
          // when it was computed rather than read from the cache, like the reference.  Region colours are arbitrary
          // there (rand()); here a label hashes to its colour, background is black, the regions too small to keep
          // (label -1) are grey.  Failing to write a picture does not fail the run.
          thread_local int my = -1;
          thread_local const void* owner = nullptr;
          if (my < 0 || owner != (const void*)&encs) {
            my = next_enc++; owner = (const void*)&encs;
            cudaSetDevice(gpus[0]);
            std::string ee;
            if (my < (int)encs.size()) encs[my] = jpeg_encoder_create(&ee);
          }
          JpegEncoder* enc = my < (int)encs.size() ? encs[my] : nullptr;
          const size_t n = (size_t)e.rows * e.cols;
          if (enc && n > 0 && l.size() == n) {
            std::vector<uint8_t> bgr(n * 3);
            std::string ee;
            if (!have_e) {
              for (size_t i = 0; i < n; ++i) bgr[3 * i] = bgr[3 * i + 1] = bgr[3 * i + 2] = e.d[i];
              jpeg_encode_bgr_host(enc, bgr.data(), e.cols, e.rows, dir + "/rawedge_" + std::to_string(j) + ".jpg", &ee);
            }
            if (!have_l) {
              for (size_t i = 0; i < n; ++i) {
                const int32_t lab = l[i];
                uint32_t h = (uint32_t)lab * 2654435761u;
                h ^= h >> 15;
                const uint8_t c0 = lab == 0 ? 0 : (lab < 0 ? 96 : (uint8_t)(64 + (h & 127)));
                const uint8_t c1 = lab == 0 ? 0 : (lab < 0 ? 96 : (uint8_t)(64 + ((h >> 8) & 127)));
                const uint8_t c2 = lab == 0 ? 0 : (lab < 0 ? 96 : (uint8_t)(64 + ((h >> 16) & 127)));
                bgr[3 * i] = c0; bgr[3 * i + 1] = c1; bgr[3 * i + 2] = c2;
              }
              jpeg_encode_bgr_host(enc, bgr.data(), e.cols, e.rows, dir + "/connect_" + std::to_string(j) + ".jpg", &ee);
            }
          }
        }
        if (!have_e) prep[v].edge[j] = e;
        if (!have_l) prep[v].label[j] = l;
      }
      { std::lock_guard<std::mutex> lk(prep_m); prep_left[v]--; }
      prep_cv.notify_all();
    });
    for (auto* enc : encs) jpeg_encoder_destroy(enc);
    prep_done_at = now_s();
  });
  // ---- fusion: colour images (cv::IMREAD_COLOR, DPE.cpp:1253) decode in the background too -----------------------
  std::vector<uint8_t> color_slab, block_slab;
  std::atomic<int> color_bad(0), block_bad(0);
  std::thread color_thread;
  if (fusion) {
    color_slab.resize(P * 3 * (size_t)n_problems);
    if (use_block) block_slab.resize(P * (size_t)n_problems);
    color_thread = std::thread([&]() {
      const int n_dec = std::max(1, std::min(4, n_problems));
      std::vector<JpegDecoder*> cd(n_dec, nullptr);
      std::atomic<int> next_dec(0);
      parallel_for(n_problems, n_dec, [&](int v) {
        thread_local int my = -1;
        thread_local const void* owner = nullptr;
        if (my < 0 || owner != (const void*)&cd) {
          my = next_dec++; owner = (const void*)&cd;
          cudaSetDevice(gpus[0]);
          std::string e; cd[my] = jpeg_decoder_create(&e);
        }
        std::vector<uint8_t> c;
        std::string e;
        int w = 0, h = 0;
        uint8_t* dst = color_slab.data() + P * 3 * (size_t)v;
        if (cd[my] && jpeg_decode_bgr(cd[my], image_path(v), &c, &w, &h, &e) && w == width && h == height) memcpy(dst, c.data(), P * 3);
        else  // grey-only JPEG: replicate luma (what cv::imread(IMREAD_COLOR) returns for it)
          for (size_t i = 0; i < P; ++i) dst[3 * i] = dst[3 * i + 1] = dst[3 * i + 2] = gray_slab[P * v + i];
        if (use_block) {  // cv::imread(IMREAD_GRAYSCALE) of blocks/mask_<id>.jpg, the id printed without padding (DPE.cpp:1265-1267)
          const std::string bp = dense + "/blocks/mask_" + std::to_string(view_ids[v]) + ".jpg";
          int bw = 0, bh = 0;
          if (!cd[my] || !jpeg_decode_gray_into(cd[my], bp, block_slab.data() + P * (size_t)v, P, &bw, &bh, &e) || bw != width || bh != height)
            block_bad++;
        }
      });
      for (auto* d : cd) jpeg_decoder_destroy(d);
    });
  }

  auto wait_prep = [&](int v) {
    std::unique_lock<std::mutex> lk(prep_m);
    prep_cv.wait(lk, [&]() { return prep_left[v] == 0; });
  };

  // ---- contexts ready?  communicator -----------------------------------------------------------------
  for (auto& t : ctx_threads) t.join();
  tm.ctx = now_s() - t_ctx0;
  auto fail = [&](const std::string& what) {
    if (prep_thread.joinable()) prep_thread.join();
    if (color_thread.joinable()) color_thread.join();
    if (G > 1) dpe_comm_reset_all();
    destroy_all();
    jpeg_decoder_destroy(dec);
    cudaFreeHost(gray_slab);
    std::cerr << "DPE-MVS: " << what << "\n";
    return 1;
  };
  for (int g = 0; g < G; ++g) if (!errs[g].empty()) return fail(errs[g]);
  if (G > 1) {
    const double tc = now_s();
    if (dpe_comm_init_all(ctxs.data(), G) != DPE_OK) return fail(std::string("communicator: ") + dpe_last_error(ctxs[0]));
    tm.comm = now_s() - tc;
  }

  if (verbose) {
    std::cout << "There are " << round_num << " resolution stages for coarse-to-fine processing!" << std::endl;
    std::cout << "Iteration nums: " << iteration_num << std::endl;
  }

  // RNG seed: the reference seeds cuRAND with clock64() (DPE.cu:1032); any value is "the reference's";
  // a fixed default makes runs reproducible, DPE_SEED overrides it.
  uint64_t seed = 20261018ull;
  if (const char* e = getenv("DPE_SEED")) seed = strtoull(e, nullptr, 10);
  // view order inside a stage: single GPU defaults to the reference's sequential order (view k reads
  // this stage's depth maps of views < k); several GPUs need the order-independent one.
  // DPE_VIEW_ORDER=parallel|sequential overrides.
  bool sequential = (G == 1);
  if (const char* e = getenv("DPE_VIEW_ORDER")) sequential = (std::string(e) == "sequential") && G == 1;
  // cost arithmetic (include/dpe_b200.h): the reference's own, operation by operation, unless DPE_ARITH says
  // fast (constant-folded homography, ~8 % faster) or centred
  int arith = -1;
  if (const char* e = getenv("DPE_ARITH")) {
    const std::string m(e);
    arith = m == "fast" ? DPE_COST_REFERENCE : (m == "centred" ? DPE_COST_CENTRED : DPE_COST_REFERENCE_EXACT);
  }
  // edge-mode propagation, direction 4 samples pixels of the colour the same launch is writing in the reference
  // (SURVEY Q3).  Default of the product surface: the reference's positions, read from a copy of the maps taken before
  // each half-sweep (dpe_set_reference_race(c, 2)) — what agrees best with the reference end to end
  // (tests/test_gpu_gate2.py) and, unlike the reference, the same result every run.  DPE_DIRECTION4=live reads them
  // live (racy exactly like the reference: two runs differ in the last digits), DPE_DIRECTION4=shifted (or the older
  // DPE_DETERMINISTIC=1) shifts direction 4 onto the other colour like directions 5-7.
  int direction4 = 2;
  if (const char* e = getenv("DPE_DETERMINISTIC")) direction4 = atoi(e) != 0 ? 0 : 2;
  if (const char* e = getenv("DPE_DIRECTION4")) {
    const std::string m(e);
    direction4 = m == "live" ? 1 : (m == "shifted" ? 0 : 2);
  }
  const std::vector<Stage> schedule = make_schedule(round_num);
  bool fusion_sharded = false;
  if (const char* e = getenv("DPE_FUSION_SHARDED")) fusion_sharded = atoi(e) != 0 && G > 1;

  // ---- .npy writers (main.cpp:570-575) run behind the GPU workers -------------------------------------
  const bool any_output = depth || normal || weak || edge;
  ExportQueue queue(P, normal != 0, weak != 0, any_output ? 2 * G + 2 : 0);
  if (any_output && !queue.usable()) return fail("cannot allocate pinned export buffers");
  std::atomic<int> write_fail(0);
  std::vector<std::thread> writers;
  const int n_writers = any_output ? std::min(4 + G, 12) : 0;
  for (int w = 0; w < n_writers; ++w) writers.emplace_back([&]() {
    std::vector<int8_t> i8;
    while (ExportBuf* b = queue.take()) {
      const int v = b->view;
      const std::string dir = out_root + "/" + format_index(view_ids[v]);
      bool ok = true;
      if (depth) ok &= write_npy(dir + "/depth.npy", b->depth, "<f4", 4, height, width, 1);
      if (normal) ok &= write_npy(dir + "/normal.npy", b->normal, "<f4", 4, height, width, 3);
      if (weak) ok &= write_npy(dir + "/weak.npy", b->weak, "|i1", 1, height, width, 1);
      queue.recycle(b);
      if (edge) {  // edges_0 > 0 (main.cpp:226-260)
        wait_prep(v);
        i8.resize(P);
        const ImageU8& e = prep[v].edge[0];
        for (size_t i = 0; i < P; ++i) i8[i] = e.d[i] > 0 ? 1 : 0;
        ok &= write_npy(dir + "/edge.npy", i8.data(), "|i1", 1, height, width, 1);
      }
      if (!ok) write_fail++;
    }
  });

  // ---- one worker per GPU: scene upload, the whole schedule, export of its views ----------------------
  std::atomic<bool> abort_flag(false);
  std::vector<double> t_upload(G, 0.0), t_stages(G, 0.0), t_prep_wait(G, 0.0), t_up_views(G, 0.0), t_up_commit(G, 0.0), t_fuse(G, 0.0);
  std::vector<std::vector<uint8_t>> cloud_records(G);  // per rank: the PLY vertex records of its views' points
  auto worker = [&](int g) {
    dpe_ctx* c = ctxs[g];
    auto bad = [&](const char* what) {
      errs[g] = std::string(what) + ": " + dpe_last_error(c);
      if (!abort_flag.exchange(true) && G > 1) dpe_comm_reset_all();  // peers blocked in a collective return
    };
    const double tu0 = now_s();
    if (dpe_scene_begin(c, n_views, width, height, round_num)) return bad("scene_begin");
    for (int v = 0; v < n_views; ++v)
      if (dpe_scene_set_view(c, v, g == 0 ? gray_slab + P * v : nullptr, cams[v].K, cams[v].R, cams[v].t, cams[v].depth_min, cams[v].depth_max))
        return bad("set_view");
    for (int v = 0; v < n_problems; ++v) {
      std::vector<int> src;
      for (int s : problems[v].src_image_ids) src.push_back(id_to_view[s]);
      if (dpe_scene_set_pairs(c, v, src.data(), (int)src.size())) return bad("set_pairs");
    }
    if (dpe_scene_set_shard(c, n_problems, g, G)) return bad("set_shard");
    if (dpe_scene_broadcast_images(c, 0)) return bad("broadcast_images");
    t_up_views[g] = now_s() - tu0;
    if (dpe_scene_commit(c)) return bad("commit");
    t_up_commit[g] = now_s() - tu0 - t_up_views[g];
    if (sequential) dpe_set_view_order(c, 1);
    if (arith >= 0) dpe_set_cost_arithmetic(c, arith);
    dpe_set_reference_race(c, direction4);
    t_upload[g] = now_s() - tu0;
    int first = 0, count = 0;
    dpe_shard_range(n_problems, G, g, &first, &count);
    JpegEncoder* venc = nullptr;  // viz images (created on first use, on this worker's GPU)
    struct EncGuard { JpegEncoder*& e; ~EncGuard() { jpeg_encoder_destroy(e); e = nullptr; } } enc_guard{venc};
    const double ts0 = now_s();
    bool prep_up = false;
    for (size_t si = 0; si < schedule.size(); ++si) {
      if (abort_flag.load()) return;
      const Stage& st = schedule[si];
      if (st.p.use_apd && !prep_up) {
        const double tw = now_s();
        for (int v = first; v < first + count; ++v) {
          wait_prep(v);
          for (int k = 0; k < round_num; ++k) {
            const int j = round_num - 1 - k;
            if (dpe_scene_set_prep(c, v, k, prep[v].edge[j].d.data(), prep[v].label[j].data(), prep[v].edge[j].d.size())) return bad("set_prep");
          }
        }
        t_prep_wait[g] = now_s() - tw;
        prep_up = true;
      }
      const bool last = si + 1 == schedule.size();
      if (dpe_stage_begin(c, st.scale, &st.p, seed)) return bad("stage_begin");
      if (last && (depth || normal || weak)) {
        for (int v = first; v < first + count; ++v) {
          if (dpe_stage_wait_view(c, v)) return bad("stage_wait_view");
          ExportBuf* b = queue.acquire();
          b->view = v;
          if (dpe_export_view(c, v, depth ? b->depth : nullptr, normal ? b->normal : nullptr, weak ? b->weak : nullptr)) {
            queue.recycle(b);
            return bad("export_view");
          }
          queue.submit(b);
        }
      } else if (last && edge) {
        for (int v = first; v < first + count; ++v) { ExportBuf* b = queue.acquire(); b->view = v; queue.submit(b); }
      }
      if (dpe_stage_end(c)) return bad("stage_end");
      if (dpe_stage_commit(c)) return bad("stage_commit");
      if (viz) {
        if (!venc) { std::string e; cudaSetDevice(gpus[g]); venc = jpeg_encoder_create(&e); }
        for (int v = first; venc && v < first + count; ++v) {
          void *d0, *d1, *d2; int vw = 0, vh = 0;
          if (dpe_viz_render(c, v, &d0, &d1, &d2, &vw, &vh)) return bad("viz_render");
          const std::string dir = out_root + "/" + format_index(view_ids[v]) + "/";
          std::string e;
          jpeg_encode_bgr_dev(venc, d0, vw, vh, dir + "depth_" + std::to_string(si) + ".jpg", &e);
          jpeg_encode_bgr_dev(venc, d1, vw, vh, dir + "normal_" + std::to_string(si) + ".jpg", &e);
          jpeg_encode_bgr_dev(venc, d2, vw, vh, dir + "weak_" + std::to_string(si) + ".jpg", &e);
        }
      }
      if (g == 0 && verbose) {
        std::cout << "Iteration " << si + 1 << " / " << iteration_num << " done" << std::endl;
        if (st.resolution_up) std::cout << "Resolution up" << std::endl;
      }
    }
    t_stages[g] = now_s() - ts0;
    if (fusion) {
      // the maps stay where the last stage left them: all-gather of the final planes + states (NCCL), colours
      // uploaded by the first GPU.  Default: the first GPU fuses all views in order, like the reference (a view sees
      // the marks of every view before it).  DPE_FUSION_SHARDED=1: colours broadcast, every GPU fuses its own block of
      // views against its own marks — G times faster, but surface seen from two blocks is fused twice (measured on a
      // 24-view ring scene, 2 GPUs: 14.6 M points instead of 11.9 M).
      const double tf0 = now_s();
      if (dpe_fuse_prepare(c)) return bad("fuse_prepare");
      if (g == 0) {
        if (color_thread.joinable()) color_thread.join();
        for (int v = 0; v < n_problems; ++v)
          if (dpe_fuse_set_color(c, v, color_slab.data() + P * 3 * (size_t)v)) return bad("fuse_set_color");
      }
      if (fusion_sharded && dpe_fuse_broadcast_colors(c, 0)) return bad("fuse_broadcast_colors");   // also: the colour thread has ended
      if (use_block) {
        // the reference reads an empty Mat for a missing mask and indexes it; here a missing or mis-sized mask is an error
        if (block_bad.load()) { errs[g] = "a blocks/mask_<id>.jpg is missing, undecodable or not of the images' size"; return; }
        const int b0 = fusion_sharded ? first : 0, b1 = fusion_sharded ? first + count : (g == 0 ? n_problems : 0);
        for (int v = b0; v < b1; ++v)
          if (dpe_fuse_set_block(c, v, block_slab.data() + P * (size_t)v)) return bad("fuse_set_block");
      }
      if (fusion_sharded || g == 0) {
        size_t n_points = 0;
        if (dpe_fuse_run(c, fusion_sharded ? first : 0, fusion_sharded ? count : n_problems, &n_points)) return bad("fuse_run");
        cloud_records[g].resize(n_points * 15);
        if (n_points && dpe_fuse_get_ply_records(c, cloud_records[g].data())) return bad("fuse_get_ply_records");
      }
      t_fuse[g] = now_s() - tf0;
    }
  };
  t0 = now_s();
  if (G == 1) worker(0);
  else {
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g) th.emplace_back(worker, g);
    for (auto& t : th) t.join();
  }
  queue.close();
  for (auto& w : writers) w.join();
  prep_thread.join();
  tm.prep = prep_done_at.load() - t_prep0;
  for (int g = 0; g < G; ++g) {
    tm.upload = std::max(tm.upload, t_upload[g]); tm.stages = std::max(tm.stages, t_stages[g]);
    tm.prep_wait = std::max(tm.prep_wait, t_prep_wait[g]);
  }
  { double mf = 0; for (int g = 0; g < G; ++g) mf = std::max(mf, t_fuse[g]);
    tm.output = std::max(0.0, now_s() - t0 - tm.upload - tm.stages - mf); }  // what the writers needed beyond the last stage
  for (int g = 0; g < G; ++g) if (!errs[g].empty()) return fail(errs[g]);
  if (write_fail.load()) return fail("cannot write the .npy outputs");
  for (int g = 0; g < G; ++g) { tm.gpu_ms = std::max(tm.gpu_ms, dpe_stage_gpu_ms(ctxs[g])); tm.launches += dpe_kernel_launches(ctxs[g]); }

  if (color_thread.joinable()) color_thread.join();
  if (fusion) {
    // the ranks' clouds in rank order = view order (ExportPointCloud, DPE.cpp:532-572)
    t0 = now_s();
    std::vector<const uint8_t*> parts; std::vector<size_t> counts;
    for (int g = 0; g < G; ++g) { parts.push_back(cloud_records[g].data()); counts.push_back(cloud_records[g].size() / 15); }
    if (!write_ply_records(dense + "/DPE/DPE.ply", parts, counts)) return fail("cannot write DPE.ply");
    for (int g = 0; g < G; ++g) tm.fusion = std::max(tm.fusion, t_fuse[g]);
    tm.fusion += now_s() - t0;
  }
  const double t_destroy0 = now_s();
  if (G == 1) destroy_all();
  else {  // freeing a few GB per GPU takes a while: all GPUs at once
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g) th.emplace_back([&, g]() { dpe_ctx_destroy(ctxs[g]); ctxs[g] = nullptr; });
    for (auto& t : th) t.join();
  }
  jpeg_decoder_destroy(dec);
  cudaFreeHost(gray_slab);
  const double t_destroy = now_s() - t_destroy0;

  // ---- cleanup of intermediates the reference deletes (main.cpp:581-595) ------------------------
  for (int v = 0; v < n_problems; ++v) {
    const std::string dir = out_root + "/" + format_index(view_ids[v]);
    for (int j = 0; j < round_num; j++) {
      remove((dir + "/edges_" + std::to_string(j) + ".dmb").c_str());
      remove((dir + "/labels_" + std::to_string(j) + ".dmb").c_str());
    }
  }
  tm.total = now_s() - t_begin;
  if (const char* tj = getenv("DPE_TIMING_JSON")) {
    FILE* f = fopen(tj, "w");
    if (f) {
      fprintf(f,
              "{\"views\": %d, \"gpus\": %d, \"width\": %d, \"height\": %d, \"load_s\": %.6f, \"ctx_s\": %.6f, \"comm_init_s\": %.6f, "
              "\"prep_background_s\": %.6f, \"prep_wait_s\": %.6f, \"upload_s\": %.6f, \"upload_parts_s\": [%.6f, %.6f], \"stages_s\": %.6f, \"output_tail_s\": %.6f, "
              "\"fusion_s\": %.6f, \"teardown_s\": %.6f, \"total_s\": %.6f, \"gpu_ms\": %.3f, \"kernel_launches\": %lld}\n",
              n_problems, G, width, height, tm.load, tm.ctx, tm.comm, tm.prep, tm.prep_wait, tm.upload, t_up_views[0], t_up_commit[0], tm.stages, tm.output, tm.fusion,
              t_destroy, tm.total, tm.gpu_ms, tm.launches);
      fclose(f);
    }
  }
  if (verbose) std::cout << "All done" << std::endl;
  return 0;
}
