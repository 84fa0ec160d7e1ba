// pipeline.cpp — C++ host orchestrator: dpe_run_pipeline(), the drop-in for
// RunDPEPipeline (main.cpp:474-600) + ProcessProblem (main.cpp:411-472).  It talks to the
// CUDA side only through the C ABI (include/dpe_b200.h).
//
// Differences from the reference that are deliberate:
//  * the scene is loaded once (one JPEG decode per image instead of (N+1) x 4R decodes) and
//    stays in HBM; per-view state never touches the disk between stages (the reference
//    round-trips depths.dmb / normals.dmb / weak.bin / selected_views.bin per view-stage);
//  * all views of a stage read the previous stage's depth maps (Jacobi) instead of a mix of
//    this stage's and the previous stage's (Gauss-Seidel through files), which
//    makes the result independent of how views are spread over streams and GPUs;
//  * gpu_index >= 0 selects that GPU; gpu_index < 0 (or env DPE_GPUS=0,1,..) shards the
//    views over several GPUs of the box, exchanging depth atlases with peer copies;
//  * errors are returned (never exit()).
// Kept: the verbose lines, the stage schedule and parameters, the edges_k/labels_k .dmb
// cache (used if present, deleted at the end), the .npy outputs, DPE.ply when fusion is on.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <chrono>
#include <iostream>
#include <map>
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
  double up_create = 0, up_views = 0, up_commit = 0;  // parts of `upload` on the first GPU
  long long launches = 0;
};

}  // namespace

extern "C" DPE_API int dpe_run_pipeline(const char* dense_folder_c, int gpu_index, int verbose, int fusion, int viz,
                                        int depth, int normal, int weak, int edge) {
  (void)viz;  // visualisation jpgs are out of scope; the flag is accepted
  const double t_begin = now_s();
  Timing tm;
  const std::string dense = dense_folder_c ? dense_folder_c : "";
  const std::string out_root = dense + "/DPE";
  mkdir(out_root.c_str(), 0777);

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
  if (const char* env = getenv("DPE_GPUS")) {
    std::string s(env);
    size_t pos = 0;
    while (pos < s.size()) {
      size_t c = s.find(',', pos);
      if (c == std::string::npos) c = s.size();
      if (c > pos) gpus.push_back(atoi(s.substr(pos, c - pos).c_str()));
      pos = c + 1;
    }
  }
  if (gpus.empty()) {
    if (gpu_index >= 0) gpus.push_back(gpu_index);
    else {
      int n = 0;
      cudaGetDeviceCount(&n);
      for (int i = 0; i < n; ++i) gpus.push_back(i);
    }
  }
  if (gpus.empty()) { std::cerr << "DPE-MVS: no CUDA device\n"; return 1; }
  if ((int)gpus.size() > n_problems) gpus.resize(n_problems);
  const int G = (int)gpus.size();
  if (cudaSetDevice(gpus[0]) != cudaSuccess) { std::cerr << "DPE-MVS: cannot select GPU " << gpus[0] << "\n"; return 1; }

  // ---- images + cameras (CheckImages: all the same size) --------------------------------------
  double t0 = now_s();
  std::string err;
  JpegDecoder* dec = jpeg_decoder_create(&err);
  if (!dec) { std::cerr << "DPE-MVS: " << err << "\n"; return 1; }
  std::vector<ImageU8> grays(n_views);
  std::vector<CamFile> cams(n_views);
  int width = 0, height = 0;
  {
    // one nvJPEG decoder per loader thread (Huffman decoding is host work, the rest runs on gpus[0])
    const int n_load = std::max(1, std::min({(int)std::thread::hardware_concurrency(), 8, n_views}));
    std::vector<JpegDecoder*> decs(n_load, nullptr);
    decs[0] = dec;
    std::vector<int> ws(n_views, 0), hs(n_views, 0);
    std::atomic<int> bad(0), next_dec(0);
    parallel_for(n_views, n_load, [&](int v) {
      thread_local int my = -1;
      thread_local const void* owner = nullptr;
      if (my < 0 || owner != (const void*)&decs) {
        my = next_dec++; owner = (const void*)&decs;
        cudaSetDevice(gpus[0]);
        if (my > 0) { std::string e; decs[my] = jpeg_decoder_create(&e); }
      }
      std::string e;
      const std::string ip = dense + "/images/" + format_index(view_ids[v]) + ".jpg";
      if (!decs[my] || !jpeg_decode_gray(decs[my], ip, &grays[v].d, &ws[v], &hs[v], &e)) { bad++; return; }
      grays[v].cols = ws[v]; grays[v].rows = hs[v];
      if (!read_cam(dense + "/cams/" + format_index(view_ids[v]) + "_cam.txt", &cams[v])) bad += 1000;
    });
    for (int i = 1; i < n_load; ++i) jpeg_decoder_destroy(decs[i]);
    width = ws[0]; height = hs[0];
    bool same = bad.load() % 1000 == 0;
    for (int v = 0; v < n_views && same; ++v) same = (ws[v] == width && hs[v] == height);
    if (!same) {  // CheckImages (main.cpp:310-329)
      std::cerr << "Images may error, check it!\n";
      jpeg_decoder_destroy(dec);
      return 1;
    }
    if (bad.load() >= 1000) {
      std::cerr << "DPE-MVS: cannot read a camera file\n";
      jpeg_decoder_destroy(dec);
      return 1;
    }
  }
  tm.load = now_s() - t0;
  if (verbose) std::cout << "There are " << n_problems << " images to be processed!" << std::endl;
  const int round_num = compute_round_num(width, height);
  const int iteration_num = round_num * 4;

  // ---- edge / label preparation (GetProblemEdges), cached .dmb files honoured -----------------
  t0 = now_s();
  std::vector<ViewPrep> prep(n_problems);
  {
    const int hw = (int)std::thread::hardware_concurrency();
    for (int v = 0; v < n_problems; ++v) { prep[v].edge.resize(round_num); prep[v].label.resize(round_num); }
    parallel_for(n_problems * round_num, hw > 0 ? hw : 4, [&](int job) {
      const int v = job / round_num, j = job % round_num;
      const std::string dir = out_root + "/" + format_index(view_ids[v]);
      const std::string ep = dir + "/edges_" + std::to_string(j) + ".dmb";
      const std::string lp = dir + "/labels_" + std::to_string(j) + ".dmb";
      bool have_e = false, have_l = false;
      int r, c, t;
      std::vector<uint8_t> buf;
      if (file_exists(ep) && read_dmb(ep, &r, &c, &t, &buf) && t == DMB_8UC1) {
        prep[v].edge[j].rows = r; prep[v].edge[j].cols = c; prep[v].edge[j].d = buf;
        have_e = true;
      }
      if (file_exists(lp) && read_dmb(lp, &r, &c, &t, &buf) && t == DMB_32SC1) {
        prep[v].label[j].resize((size_t)r * c);
        memcpy(prep[v].label[j].data(), buf.data(), buf.size());
        have_l = true;
      }
      if (!have_e || !have_l) {
        ImageU8 e;
        std::vector<int32_t> l;
        int oc, orr;
        problem_edges(grays[v], 1 << j, &e, &l, &oc, &orr);
        if (!have_e) prep[v].edge[j] = e;
        if (!have_l) prep[v].label[j] = l;
      }
    });
  }
  tm.prep = now_s() - t0;

  // ---- contexts + scene upload ------------------------------------------------------------------
  t0 = now_s();
  // views are split into G contiguous blocks of spr = ceil(V/G); atlas slot = view index, so the
  // block of rank g is also its chunk of the all-gather
  const int slots = (n_views + G - 1) / G;
  std::vector<dpe_ctx*> ctxs(G, nullptr);
  auto fail = [&](const char* what, dpe_ctx* c) {
    std::cerr << "DPE-MVS: " << what << ": " << (c ? dpe_last_error(c) : "") << "\n";
    for (auto* x : ctxs) dpe_ctx_destroy(x);
    jpeg_decoder_destroy(dec);
    return 1;
  };
  const int slots_per_rank = slots;
  {
    std::vector<std::string> errs(G);
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g) th.emplace_back([&, g]() {
      auto bad = [&](const char* what) { errs[g] = std::string(what) + ": " + (ctxs[g] ? dpe_last_error(ctxs[g]) : ""); };
      const double tu0 = now_s();
      if (dpe_ctx_create(&ctxs[g], gpus[g]) != DPE_OK) { errs[g] = "cannot create context"; return; }
      dpe_ctx* c = ctxs[g];
      const double tu1 = now_s();
      if (dpe_scene_begin(c, n_views, width, height, round_num)) return bad("scene_begin");
      for (int v = 0; v < n_views; ++v)
        if (dpe_scene_set_view(c, v, grays[v].d.data(), cams[v].K, cams[v].R, cams[v].t, cams[v].depth_min, cams[v].depth_max))
          return bad("set_view");
      for (int v = 0; v < n_problems; ++v) {
        std::vector<int> src;
        for (int s : problems[v].src_image_ids) src.push_back(id_to_view[s]);
        if (dpe_scene_set_pairs(c, v, src.data(), (int)src.size())) return bad("set_pairs");
      }
      const int first = std::min(g * slots, n_problems), count = std::max(0, std::min(slots, n_problems - first));
      for (int v = first; v < first + count; ++v)
        for (int k = 0; k < round_num; ++k) {
          const int j = round_num - 1 - k;
          if (dpe_scene_set_prep(c, v, k, prep[v].edge[j].d.data(), prep[v].label[j].data())) return bad("set_prep");
        }
      if (dpe_scene_set_shard(c, first, count, slots_per_rank, G)) return bad("set_shard");
      const double tu2 = now_s();
      if (dpe_scene_commit(c)) return bad("commit");
      if (g == 0) { tm.up_create = tu1 - tu0; tm.up_views = tu2 - tu1; tm.up_commit = now_s() - tu2; }
    });
    for (auto& t : th) t.join();
    for (int g = 0; g < G; ++g)
      if (!errs[g].empty()) {
        std::cerr << "DPE-MVS: " << errs[g] << "\n";
        for (auto* x : ctxs) dpe_ctx_destroy(x);
        jpeg_decoder_destroy(dec);
        return 1;
      }
    cudaSetDevice(gpus[0]);
  }
  tm.upload = now_s() - t0;

  if (verbose) {
    std::cout << "There are " << round_num << " resolution stages for coarse-to-fine processing!" << std::endl;
    std::cout << "Iteration nums: " << iteration_num << std::endl;
  }

  // ---- stage loop (main.cpp:507-567) ------------------------------------------------------------
  t0 = now_s();
  // RNG seed: the reference seeds cuRAND with clock64() (DPE.cu:1032); any value is "the reference's";
  // a fixed default makes runs reproducible, DPE_SEED overrides it.
  uint64_t seed = 20261018ull;
  if (const char* e = getenv("DPE_SEED")) seed = strtoull(e, nullptr, 10);
  // view order inside a stage: single GPU defaults to the reference's sequential order (view k reads
  // this stage's depth maps of views < k); several GPUs need the order-independent one.
  // DPE_VIEW_ORDER=parallel|sequential overrides.
  {
    bool sequential = (G == 1);
    if (const char* e = getenv("DPE_VIEW_ORDER")) sequential = (std::string(e) == "sequential") && G == 1;
    if (sequential) dpe_set_view_order(ctxs[0], 1);
  }
  // cost arithmetic (include/dpe_b200.h): the reference's own, operation by operation, unless DPE_ARITH says
  // fast (constant-folded homography, ~8 % faster) or centred
  if (const char* e = getenv("DPE_ARITH")) {
    const std::string m(e);
    const int mode = m == "fast" ? DPE_COST_REFERENCE : (m == "centred" ? DPE_COST_CENTRED : DPE_COST_REFERENCE_EXACT);
    for (auto* c : ctxs) dpe_set_cost_arithmetic(c, mode);
  }
  int iteration_index = 0;
  auto run_stage_all = [&](int k, const dpe_stage_params& p) -> int {
    std::vector<int> rcs(G, 0);
    if (G == 1) {
      rcs[0] = dpe_run_stage(ctxs[0], k, &p, seed);
    } else {
      std::vector<std::thread> th;
      for (int g = 0; g < G; ++g) th.emplace_back([&, g]() { rcs[g] = dpe_run_stage(ctxs[g], k, &p, seed); });
      for (auto& t : th) t.join();
    }
    for (int g = 0; g < G; ++g) if (rcs[g]) return fail("run_stage", ctxs[g]);
    if (G > 1) {
      // all-gather of the depth atlas: every rank's chunk is copied to every peer
      std::vector<void*> ptr(G);
      size_t chunk = 0, total = 0;
      for (int g = 0; g < G; ++g) dpe_stage_atlas(ctxs[g], &ptr[g], &chunk, &total);
      for (int dst = 0; dst < G; ++dst) {
        cudaSetDevice(gpus[dst]);
        for (int src = 0; src < G; ++src)
          if (src != dst)
            cudaMemcpyPeerAsync((char*)ptr[dst] + (size_t)src * chunk, gpus[dst], (char*)ptr[src] + (size_t)src * chunk, gpus[src], chunk, 0);
      }
      for (int g = 0; g < G; ++g) { cudaSetDevice(gpus[g]); cudaDeviceSynchronize(); }
      cudaSetDevice(gpus[0]);
    }
    for (int g = 0; g < G; ++g) if (dpe_stage_commit(ctxs[g])) return fail("stage_commit", ctxs[g]);
    return 0;
  };
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
    if (int rc = run_stage_all(i, p)) return rc;
    if (verbose) std::cout << "Iteration " << iteration_index + 1 << " / " << iteration_num << " done" << std::endl;
    iteration_index++;
    for (int j = 0; j < 3; ++j) {
      p.state = DPE_REFINE_ITER; p.use_apd = i == 0 ? 0 : 1;
      p.ransac_threshold = (float)(0.01 - i * 0.00125);
      p.rotate_time = std::min((int)(1 << i), 4);
      p.geom_consistency = 1; p.weak_peak_radius = std::max(4 - 2 * j, 2);
      if (int rc = run_stage_all(i, p)) return rc;
      if (verbose) std::cout << "Iteration " << iteration_index + 1 << " / " << iteration_num << " done" << std::endl;
      iteration_index++;
    }
    if (verbose) std::cout << "Resolution up" << std::endl;
  }
  tm.stages = now_s() - t0;
  for (int g = 0; g < G; ++g) { tm.gpu_ms = std::max(tm.gpu_ms, dpe_stage_gpu_ms(ctxs[g])); tm.launches += dpe_kernel_launches(ctxs[g]); }

  // ---- results: .npy export (main.cpp:570-575), optional fusion ---------------------------------
  t0 = now_s();
  const size_t P = (size_t)width * height;
  std::vector<std::vector<float>> all_depth, all_normal;
  std::vector<std::vector<uint8_t>> all_state;
  if (fusion) { all_depth.resize(n_problems); all_normal.resize(n_problems); all_state.resize(n_problems); }
  {
    // one reader thread per GPU pulls its views' maps off the device; the .npy files are written by
    // a few writer tasks behind it
    struct Job { int v; std::vector<float> d, n3; std::vector<uint8_t> st; };
    std::vector<std::string> errs(G);
    std::atomic<int> write_fail(0);
    auto write_view = [&](Job* job) {
      const int v = job->v;
      const std::string dir = out_root + "/" + format_index(view_ids[v]);
      std::vector<int8_t> i8;
      bool ok = true;
      if (depth) {
        std::vector<float> dz(job->d);
        for (size_t i = 0; i < P; ++i) if (job->st[i] == DPE_UNKNOWN) dz[i] = 0.0f;  // ZeroDepthForUnknown
        ok &= write_npy(dir + "/depth.npy", dz.data(), "<f4", 4, height, width, 1);
      }
      if (normal) ok &= write_npy(dir + "/normal.npy", job->n3.data(), "<f4", 4, height, width, 3);
      if (weak) {
        i8.resize(P);
        for (size_t i = 0; i < P; ++i) { const uint8_t st = job->st[i]; i8[i] = st == DPE_WEAK ? 1 : (st == DPE_STRONG ? 2 : 0); }
        ok &= write_npy(dir + "/weak.npy", i8.data(), "|i1", 1, height, width, 1);
      }
      if (edge) {
        i8.resize(P);
        const ImageU8& e = prep[v].edge[0];
        for (size_t i = 0; i < P; ++i) i8[i] = e.d[i] > 0 ? 1 : 0;
        ok &= write_npy(dir + "/edge.npy", i8.data(), "|i1", 1, height, width, 1);
      }
      if (!ok) write_fail++;
      if (fusion) { all_depth[v].swap(job->d); all_normal[v].swap(job->n3); all_state[v].swap(job->st); }
      delete job;
    };
    std::vector<std::thread> readers;
    for (int g = 0; g < G; ++g) readers.emplace_back([&, g]() {
      cudaSetDevice(gpus[g]);
      const int first = std::min(g * slots, n_problems), count = std::max(0, std::min(slots, n_problems - first));
      std::vector<std::thread> writers;
      for (int v = first; v < first + count; ++v) {
        Job* job = new Job{v, std::vector<float>(P), std::vector<float>((normal || fusion) ? P * 3 : 0), std::vector<uint8_t>(P)};
        if (dpe_get_maps(ctxs[g], v, job->d.data(), job->n3.empty() ? nullptr : job->n3.data(), job->st.data(), nullptr)) {
          errs[g] = dpe_last_error(ctxs[g]);
          delete job;
          break;
        }
        if (writers.size() >= 4) { writers.front().join(); writers.erase(writers.begin()); }
        writers.emplace_back(write_view, job);
      }
      for (auto& w : writers) w.join();
    });
    for (auto& r : readers) r.join();
    cudaSetDevice(gpus[0]);
    for (int g = 0; g < G; ++g) if (!errs[g].empty()) { std::cerr << "DPE-MVS: get_maps: " << errs[g] << "\n"; return fail("get_maps", nullptr); }
    if (write_fail.load()) { std::cerr << "DPE-MVS: cannot write the .npy outputs\n"; return fail("write", nullptr); }
  }
  tm.output = now_s() - t0;
  if (fusion) {
    // on the device of the first context: maps of every problem view + colour images go up, the cloud comes back
    t0 = now_s();
    dpe_ctx* fc = ctxs[0];
    cudaSetDevice(gpus[0]);
    std::vector<uint8_t> color;
    for (int v = 0; v < n_problems; ++v) {
      int w, h;
      if (!jpeg_decode_bgr(dec, dense + "/images/" + format_index(view_ids[v]) + ".jpg", &color, &w, &h, &err)) {
        color.resize(P * 3);  // grey-only JPEG: replicate luma
        for (size_t i = 0; i < P; ++i) color[3 * i] = color[3 * i + 1] = color[3 * i + 2] = grays[v].d[i];
      }
      if (dpe_fuse_set_view(fc, v, all_depth[v].data(), all_normal[v].data(), all_state[v].data(), color.data())) return fail("fuse_set_view", fc);
      std::vector<float>().swap(all_depth[v]); std::vector<float>().swap(all_normal[v]);
    }
    size_t n_points = 0;
    if (dpe_fuse_run(fc, &n_points)) return fail("fuse_run", fc);
    std::vector<float> xyz(n_points * 3);
    std::vector<uint8_t> bgr(n_points * 3);
    if (n_points && dpe_fuse_get(fc, xyz.data(), bgr.data())) return fail("fuse_get", fc);
    if (!write_ply(dense + "/DPE/DPE.ply", xyz.data(), bgr.data(), n_points)) { std::cerr << "DPE-MVS: cannot write DPE.ply\n"; return fail("write_ply", nullptr); }
    tm.fusion = now_s() - t0;
  }
  for (auto* c : ctxs) dpe_ctx_destroy(c);
  ctxs.assign(G, nullptr);
  jpeg_decoder_destroy(dec);

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
              "{\"views\": %d, \"gpus\": %d, \"width\": %d, \"height\": %d, \"load_s\": %.6f, \"prep_s\": %.6f, "
              "\"upload_s\": %.6f, \"upload_parts_s\": [%.6f, %.6f, %.6f], \"stages_s\": %.6f, \"output_s\": %.6f, \"fusion_s\": %.6f, \"total_s\": %.6f, "
              "\"gpu_ms\": %.3f, \"kernel_launches\": %lld}\n",
              n_problems, G, width, height, tm.load, tm.prep, tm.upload, tm.up_create, tm.up_views, tm.up_commit, tm.stages, tm.output, tm.fusion, tm.total,
              tm.gpu_ms, tm.launches);
      fclose(f);
    }
  }
  if (verbose) std::cout << "All done" << std::endl;
  return 0;
}
