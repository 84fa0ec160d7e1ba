// dpe_capi.cu — C-ABI layer (include/dpe_b200.h) over the sm_100a kernels: context,
// scene-resident device state, per-stage orchestration, the NCCL exchange between GPUs.
//
// Where the reference rebuilds everything per (view, stage) — N+1 JPEG decodes, cv::resize,
// 2(N+1) cudaMallocArray + texture objects, ~20 cudaMallocs, .dmb round trips
// (DPE.cpp:733-1023, main.cpp:411-446) — this layer uploads a scene once: all pyramid
// levels of all images as float textures + linear copies, cameras folded to per-pair
// constants in double precision, prep arrays, and per-view PatchMatch state that never
// leaves HBM between stages.  Source depth maps for geometric consistency live in one
// "depth atlas" per scale ([slot][pixel] floats); with several GPUs each rank owns a block of
// reference views and the atlas is all-gathered in place with NCCL, one collective per view
// slot, issued on a side stream as soon as that view's last kernel has written its slot.
// Nothing is allocated or freed while a stage is in flight (an implicit device synchronisation
// under an NCCL kernel that waits for a peer is how multi-GPU processes deadlock).
#include <cuda_runtime.h>
#include <cuda_profiler_api.h>
#include <cuda_fp16.h>
#include <nccl.h>  // types and prototypes only: the library itself is bound at run time (NcclApi below)
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>
#include "dpe_kernels.cuh"
#include "dpe_fusion.cuh"
#include "dpe_consts.h"

using namespace dpe;

namespace {

struct ScaleImg {
  float* lin = nullptr;          // W*H float
  // prep arrays (optional): pointers into dpe_ctx::edge_slab / label_slab / bits_slab once supplied
  uint8_t* edge = nullptr;       // edges_k
  uint32_t* edge_bits = nullptr; // coarsest scale only: the same map, one bit per pixel (rows of (w+31)/32 words)
  int32_t* label = nullptr;      // labels_k
  // prep handed in before dpe_scene_commit waits here until the slabs exist
  std::vector<uint8_t> pending_edge;
  std::vector<int32_t> pending_label;
};

struct ViewData {
  HostCam cam;
  bool have_cam = false;
  bool have_img = false;
  std::vector<int> src;
  std::vector<float> relpose;  // per source: R_rel[9], t_rel[3], baseline as the device computes them (k_relative_pose)
  std::vector<ScaleImg> scales;
  // state carried between stages (owner only): pointers into dpe_ctx::maps_*[cur_buf]
  float4* planes = nullptr;  // (world normal, depth)
  uint8_t* state = nullptr;
  uint32_t* selected = nullptr;
  int cur_scale = -1;  // scale index of planes/state/selected
  int cur_buf = 0;     // which of the two map buffers holds them (a stage at a new scale writes the other one)
};

struct Scratch {
  float4* planes = nullptr; float* costs = nullptr; uint32_t* selected = nullptr; uint4* view_w = nullptr;
  uint8_t* state = nullptr; float4* fit_planes = nullptr; int* radius = nullptr; short2* edge_neigh = nullptr;
  float* complexity = nullptr; short2* label_boundary = nullptr; uint8_t* weak_reliable = nullptr;
  short2* nearest_strong = nullptr; short2* neighbours = nullptr;
  Xorwow* rng = nullptr;
  int* weak_list = nullptr; int* weak_count = nullptr; int* weak_scan = nullptr;
  float4* snap_planes = nullptr; float* snap_costs = nullptr;  // dpe_set_reference_race(ctx, 2), allocated on first use
  cudaStream_t stream = nullptr;
};

// NCCL is bound at run time, on the first call that needs a communicator, not linked: a process that never
// leaves one GPU never loads it, and a process that already holds an NCCL (PyTorch ships its own, newer than
// the system's, under the same soname) keeps using that one instead of getting a second copy or the wrong one.
// Search order: the libnccl.so.2 already in the process, $DPE_NCCL_LIB, the loader's libnccl.so.2.
struct NcclApi {
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclCommAbort) CommAbort = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  bool ok = false;
  std::string why;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;
const NcclApi& nccl_api() {
  std::call_once(g_nccl_once, []() {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) if (const char* e = getenv("DPE_NCCL_LIB")) h = dlopen(e, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { g_nccl.why = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
    bool all = true;
    auto bind = [&](auto& fn, const char* name) {
      fn = reinterpret_cast<typename std::remove_reference<decltype(fn)>::type>(dlsym(h, name));
      if (!fn) { all = false; g_nccl.why = std::string("libnccl.so.2 lacks ") + name; }
    };
    bind(g_nccl.GetUniqueId, "ncclGetUniqueId"); bind(g_nccl.CommInitRank, "ncclCommInitRank");
    bind(g_nccl.CommInitAll, "ncclCommInitAll"); bind(g_nccl.CommDestroy, "ncclCommDestroy");
    bind(g_nccl.CommAbort, "ncclCommAbort"); bind(g_nccl.AllGather, "ncclAllGather");
    bind(g_nccl.Broadcast, "ncclBroadcast"); bind(g_nccl.GetErrorString, "ncclGetErrorString");
    g_nccl.ok = all;
  });
  return g_nccl;
}

// NCCL communicators of single-process multi-GPU runs (dpe_comm_init_all), kept for the life of the
// process: creating one costs far more than a stage, a second dpe_run_pipeline call reuses it.
struct CommSet { std::vector<ncclComm_t> comms; };
std::mutex g_comm_mutex;
std::map<std::vector<int>, CommSet> g_comm_cache;

}  // namespace

struct dpe_ctx {
  int device = 0;
  int num_sms = 148;
  std::string err;
  long long launches = 0;
  // scene
  int n_views = 0, W = 0, H = 0, n_scales = 0;
  std::vector<int> sw, sh;  // per scale index (0 = coarsest)
  std::vector<ViewData> views;
  // one layered array + texture object per scale: layer v = image of view v (float, linear filter)
  std::vector<cudaArray_t> scale_arr;
  std::vector<cudaTextureObject_t> scale_tex;
  // slabs: one allocation for all uploaded u8 images, one per scale for all float copies (a cudaMalloc /
  // cudaFree per view serialises on the driver lock and synchronises the device)
  uint8_t* gray_slab = nullptr;
  std::vector<float*> lin_slab;
  // prep arrays of the owned views: per scale [n_local][P_k]; bit-packed coarsest edges [n_local][words*h]
  std::vector<uint8_t*> edge_slab;
  std::vector<int32_t*> label_slab;
  uint32_t* bits_slab = nullptr;
  // carried maps of the owned views, double-buffered, sized for the finest scale: [2][n_local][P_full]
  float4* maps_planes[2] = {nullptr, nullptr};
  uint8_t* maps_state[2] = {nullptr, nullptr};
  uint32_t* maps_selected[2] = {nullptr, nullptr};
  bool committed = false;
  // shard: problems [0, n_problems) split into n_ranks contiguous balanced blocks (dpe_shard_range); atlas slot of
  // a problem view = local index * n_ranks + owner rank, so that the views with local index i of all ranks form
  // one contiguous chunk — the in-place ncclAllGather of view slot i
  int n_problems = 0, rank = 0, n_ranks = 1;
  int first_view = 0, n_local = 0, slots_per_rank = 0;
  ncclComm_t comm = nullptr;
  bool comm_owned = false;  // created by dpe_comm_init_rank (destroyed with the context); cached ones are not
  cudaStream_t comm_stream = nullptr, upload_stream = nullptr, copy_stream = nullptr;
  std::vector<cudaEvent_t> view_done;  // per local view: its last kernel of the running stage
  RefConst* h_rc = nullptr;            // pinned: folded cameras of the running stage's views (source of the async copies into c_rc)
  cudaEvent_t comm_done = nullptr;
  bool comm_queued = false;  // the running stage has all-gathers in flight
  // device staging of dpe_export_view (depth, normal3, weak as the .npy files hold them)
  float* exp_depth = nullptr; float* exp_normal = nullptr; int8_t* exp_weak = nullptr;
  uint8_t* viz_bgr = nullptr;  // dpe_viz_render: three BGR images
  // depth atlas per scale: front = committed (read by geom stages), back = being written
  std::vector<float*> atlas_front, atlas_back;
  int last_stage_scale = -1;
  bool stage_pending = false;
  bool stage_open = false;    // between dpe_stage_begin and dpe_stage_end
  bool gauss_seidel = false;  // dpe_set_view_order
  int ref_race = 0;           // dpe_set_reference_race
  bool cost_raw = true;       // dpe_set_cost_arithmetic
  bool exact = true;          // dpe_set_cost_arithmetic: DPE_COST_REFERENCE_EXACT is the default
  int variants = 0;           // dpe_debug_set_variants
  // scratch
  std::vector<Scratch> scratch;
  // device fusion (dpe_fuse_*): maps of all views at full resolution (own copies, or the carried maps themselves on
  // one GPU), colour images, marks, and the fused cloud (host)
  float4* fuse_planes = nullptr; uint8_t* fuse_state = nullptr;          // what dpe_fuse_run reads
  float4* fuse_planes_own = nullptr; uint8_t* fuse_state_own = nullptr;  // allocations (host-provided or gathered maps)
  uint8_t* fuse_bgr = nullptr; uint16_t* fuse_mask = nullptr;
  uint8_t* fuse_block = nullptr; std::vector<char> fuse_block_have;  // dpe_fuse_set_block
  std::vector<char> fuse_have;
  bool fuse_resident = false;
  std::vector<FusedPointDev> cloud;
  // initial XORWOW states per scale for rng_seed (dpe_rng.h): table[k][y*w+x] = curand_init(seed, y, x)
  std::vector<Xorwow*> rng_table;
  uint64_t rng_seed = 0;
  bool rng_ready = false;
  uint8_t* zero_edge = nullptr;  // placeholder when no prep was supplied
  int32_t* zero_label = nullptr;
  unsigned long long* d_eval_units = nullptr;
  bool count_evals = false;
  double eval_units_total = 0.0;
  double stage_ms = 0.0;
  double comm_ms = 0.0;
  uint32_t stage_counter = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_views = nullptr;
  // per-kernel-class profile (dpe_set_profile): single stream, CUDA events around each launch
  bool profile = false;
  int debug_stop_after = -1;  // dpe_debug_stop_after: last step of a view-stage that still runs (-1: all)
  int profile_views = 0;  // profile only the first n local views (the others run normally)
  double prof_ms[DPE_N_KERNEL_CLASSES] = {0};
  double prof_units[DPE_N_KERNEL_CLASSES] = {0};
  long long prof_launches[DPE_N_KERNEL_CLASSES] = {0};
  cudaEvent_t pa = nullptr, pb = nullptr;
};

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      char buf_[512];                                                                    \
      snprintf(buf_, sizeof(buf_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      ctx->err = buf_;                                                                   \
      return DPE_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define NCK(call)                                                                        \
  do {                                                                                   \
    ncclResult_t r_ = (call);                                                            \
    if (r_ != ncclSuccess) {                                                             \
      char buf_[512];                                                                    \
      snprintf(buf_, sizeof(buf_), "%s failed: %s (%s:%d)", #call, nccl_api().GetErrorString(r_), __FILE__, __LINE__); \
      ctx->err = buf_;                                                                   \
      return DPE_ERR_COMM;                                                               \
    }                                                                                    \
  } while (0)

// DPE_TRACE=1: wall-clock of the set-up steps on stderr (where does an end-to-end call spend its time outside the stages)
#include <chrono>
struct Trace {
  bool on; std::chrono::steady_clock::time_point t; int dev;
  explicit Trace(int device) : on(getenv("DPE_TRACE") != nullptr), t(std::chrono::steady_clock::now()), dev(device) {}
  void operator()(const char* what) {
    if (!on) return;
    cudaDeviceSynchronize();
    const auto n = std::chrono::steady_clock::now();
    fprintf(stderr, "[dpe trace gpu %d] %-28s %8.2f ms\n", dev, what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};

// Device memory comes from the device's stream-ordered pool with the release threshold lifted: what a context
// frees stays with the process, so the next scene (the next dpe_run_pipeline call) gets it back without a trip to
// the driver — a plain cudaMalloc / cudaFree of the ~6 GB a scene holds costs 0.1-2 s per call, more with peer
// mappings between several GPUs of one process.  DPE_RELEASE_MEMORY=1 hands everything back when a context is
// destroyed.  All allocation and release happens with the device drained (dpe_scene_begin / commit / destroy).
template <class T>
static cudaError_t dmalloc(T** p, size_t bytes) { return cudaMallocAsync((void**)p, bytes ? bytes : 1, (cudaStream_t)0); }
static cudaError_t dfree(void* p) { return p ? cudaFreeAsync(p, (cudaStream_t)0) : cudaSuccess; }

#define FAIL(code, msg) \
  do { ctx->err = (msg); return (code); } while (0)

static LaunchCfg cfg_of(dpe_ctx* ctx) { return LaunchCfg{ctx->num_sms, &ctx->launches}; }

// pixels per view in carried-map buffer b (see dpe_scene_commit)
static size_t map_stride(const dpe_ctx* ctx, int b) {
  const int k = ctx->n_scales - 1 - b;
  return k >= 0 ? (size_t)ctx->sw[k] * ctx->sh[k] : 1;
}
static int map_buffer_of_scale(const dpe_ctx* ctx, int k) { return (ctx->n_scales - 1 - k) & 1; }

// atlas slot of a view (see dpe_ctx::n_problems)
static int slot_of(const dpe_ctx* ctx, int view) {
  if (view >= ctx->n_problems) return ctx->slots_per_rank * ctx->n_ranks + (view - ctx->n_problems);
  int first = 0, count = 0, r = 0;
  for (; r < ctx->n_ranks; ++r) {
    dpe_shard_range(ctx->n_problems, ctx->n_ranks, r, &first, &count);
    if (view < first + count) break;
  }
  return (view - first) * ctx->n_ranks + r;
}
static int total_slots(const dpe_ctx* ctx) { return ctx->slots_per_rank * ctx->n_ranks + (ctx->n_views - ctx->n_problems); }

static void free_scene(dpe_ctx* ctx) {
  ctx->views.clear();
  dfree(ctx->gray_slab); ctx->gray_slab = nullptr;
  for (auto p : ctx->lin_slab) dfree(p);
  ctx->lin_slab.clear();
  for (auto p : ctx->edge_slab) dfree(p);
  for (auto p : ctx->label_slab) dfree(p);
  ctx->edge_slab.clear(); ctx->label_slab.clear();
  dfree(ctx->bits_slab); ctx->bits_slab = nullptr;
  for (int b = 0; b < 2; ++b) {
    dfree(ctx->maps_planes[b]); dfree(ctx->maps_state[b]); dfree(ctx->maps_selected[b]);
    ctx->maps_planes[b] = nullptr; ctx->maps_state[b] = nullptr; ctx->maps_selected[b] = nullptr;
  }
  for (auto t : ctx->scale_tex) if (t) cudaDestroyTextureObject(t);
  for (auto a : ctx->scale_arr) if (a) cudaFreeArray(a);
  ctx->scale_tex.clear(); ctx->scale_arr.clear();
  for (auto p : ctx->atlas_front) dfree(p);
  for (auto p : ctx->atlas_back) dfree(p);
  ctx->atlas_front.clear(); ctx->atlas_back.clear();
  for (auto& s : ctx->scratch) {
    dfree(s.planes); dfree(s.costs); dfree(s.selected); dfree(s.view_w); dfree(s.state);
    dfree(s.fit_planes); dfree(s.radius); dfree(s.edge_neigh); dfree(s.complexity);
    dfree(s.label_boundary); dfree(s.weak_reliable); dfree(s.nearest_strong); dfree(s.neighbours);
    dfree(s.rng); dfree(s.weak_list); dfree(s.weak_count); dfree(s.weak_scan);
    dfree(s.snap_planes); dfree(s.snap_costs);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  ctx->scratch.clear();
  for (auto e : ctx->view_done) cudaEventDestroy(e);
  ctx->view_done.clear();
  if (ctx->h_rc) cudaFreeHost(ctx->h_rc);
  ctx->h_rc = nullptr;
  dfree(ctx->exp_depth); dfree(ctx->exp_normal); dfree(ctx->exp_weak);
  ctx->exp_depth = nullptr; ctx->exp_normal = nullptr; ctx->exp_weak = nullptr;
  dfree(ctx->viz_bgr); ctx->viz_bgr = nullptr;
  dfree(ctx->fuse_planes_own); dfree(ctx->fuse_state_own); dfree(ctx->fuse_bgr); dfree(ctx->fuse_mask); dfree(ctx->fuse_block);
  ctx->fuse_planes_own = nullptr; ctx->fuse_state_own = nullptr; ctx->fuse_bgr = nullptr; ctx->fuse_mask = nullptr;
  ctx->fuse_block = nullptr; ctx->fuse_block_have.clear();
  ctx->fuse_planes = nullptr; ctx->fuse_state = nullptr; ctx->fuse_have.clear(); ctx->fuse_resident = false;
  ctx->cloud.clear();
  for (auto p : ctx->rng_table) dfree(p);
  ctx->rng_table.clear(); ctx->rng_ready = false;
  dfree(ctx->zero_edge); ctx->zero_edge = nullptr;
  dfree(ctx->zero_label); ctx->zero_label = nullptr;
  ctx->committed = false;
  ctx->stage_open = false; ctx->stage_pending = false;
}

extern "C" {

int dpe_ctx_create(dpe_ctx** out, int gpu_index) {
  if (!out) return DPE_ERR_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return DPE_ERR_NO_DEVICE;
  if (gpu_index < 0 || gpu_index >= n) return DPE_ERR_ARG;
  if (cudaSetDevice(gpu_index) != cudaSuccess) return DPE_ERR_CUDA;
  dpe_ctx* ctx = new dpe_ctx();
  ctx->device = gpu_index;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, gpu_index) == cudaSuccess) ctx->num_sms = prop.multiProcessorCount;
  {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, gpu_index) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  bool ok = true;
  ok &= cudaEventCreate(&ctx->ev0) == cudaSuccess;
  ok &= cudaEventCreate(&ctx->ev1) == cudaSuccess;
  ok &= cudaEventCreate(&ctx->ev_views) == cudaSuccess;
  ok &= cudaEventCreate(&ctx->pa) == cudaSuccess;
  ok &= cudaEventCreate(&ctx->pb) == cudaSuccess;
  ok &= cudaEventCreateWithFlags(&ctx->comm_done, cudaEventDisableTiming) == cudaSuccess;
  // the exchange runs beside the view kernels: highest priority, so that its few CTAs are placed as soon as a
  // persistent view kernel retires
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  ok &= cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, prio_hi) == cudaSuccess;
  ok &= cudaStreamCreateWithFlags(&ctx->upload_stream, cudaStreamNonBlocking) == cudaSuccess;
  ok &= cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
  ok &= dmalloc(&ctx->d_eval_units, sizeof(unsigned long long)) == cudaSuccess;
  ok = ok && cudaMemset(ctx->d_eval_units, 0, sizeof(unsigned long long)) == cudaSuccess;
  if (!ok) {  // a context with a missing event or stream would fail later, somewhere less obvious
    cudaGetLastError();
    dpe_ctx_destroy(ctx);
    return DPE_ERR_CUDA;
  }
  *out = ctx;
  return DPE_OK;
}

void dpe_ctx_destroy(dpe_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  if (ctx->comm && ctx->comm_owned) nccl_api().CommDestroy(ctx->comm);
  ctx->comm = nullptr;
  free_scene(ctx);
  dfree(ctx->d_eval_units);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ev_views) cudaEventDestroy(ctx->ev_views);
  if (ctx->pa) cudaEventDestroy(ctx->pa);
  if (ctx->pb) cudaEventDestroy(ctx->pb);
  if (ctx->comm_done) cudaEventDestroy(ctx->comm_done);
  if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
  if (ctx->upload_stream) cudaStreamDestroy(ctx->upload_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  cudaDeviceSynchronize();
  if (const char* e = getenv("DPE_RELEASE_MEMORY")) {
    cudaMemPool_t pool;
    if (atoi(e) != 0 && cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
  }
  delete ctx;
}

const char* dpe_last_error(const dpe_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
long long dpe_kernel_launches(const dpe_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- multi-GPU: shard arithmetic + NCCL communicator -------------------------------------------
int dpe_shard_range(int n_problems, int n_ranks, int rank, int* first, int* count) {
  if (n_problems < 0 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return DPE_ERR_ARG;
  const int q = n_problems / n_ranks, rem = n_problems % n_ranks;
  if (first) *first = rank * q + (rank < rem ? rank : rem);
  if (count) *count = q + (rank < rem ? 1 : 0);
  return DPE_OK;
}

int dpe_comm_get_unique_id(void* id, size_t bytes) {
  if (!id || bytes < sizeof(ncclUniqueId)) return DPE_ERR_ARG;
  ncclUniqueId u;
  if (!nccl_api().ok || nccl_api().GetUniqueId(&u) != ncclSuccess) return DPE_ERR_COMM;
  memset(id, 0, bytes);
  memcpy(id, &u, sizeof(u));
  return DPE_OK;
}

int dpe_comm_init_rank(dpe_ctx* ctx, const void* id, int n_ranks, int rank) {
  if (!ctx || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return DPE_ERR_ARG;
  if (ctx->comm) FAIL(DPE_ERR_STATE, "context already has a communicator");
  CK(cudaSetDevice(ctx->device));
  if (!nccl_api().ok) FAIL(DPE_ERR_COMM, nccl_api().why);
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  NCK(nccl_api().CommInitRank(&ctx->comm, n_ranks, u, rank));
  ctx->comm_owned = true;
  return DPE_OK;
}

int dpe_comm_init_all(dpe_ctx** ctxs, int n) {
  if (!ctxs || n < 1) return DPE_ERR_ARG;
  std::vector<int> devs(n);
  for (int i = 0; i < n; ++i) {
    if (!ctxs[i]) return DPE_ERR_ARG;
    devs[i] = ctxs[i]->device;
  }
  dpe_ctx* ctx = ctxs[0];
  for (int i = 0; i < n; ++i)
    if (ctxs[i]->comm) FAIL(DPE_ERR_STATE, "context already has a communicator");
  if (!nccl_api().ok) FAIL(DPE_ERR_COMM, nccl_api().why);
  // within one process NCCL moves data straight between the ranks' buffers over NVLink: the memory pools the
  // buffers come from must be mapped on the peers (the analogue of cudaDeviceEnablePeerAccess for pool memory)
  for (int i = 0; i < n; ++i) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, devs[i]) != cudaSuccess) continue;
    for (int j = 0; j < n; ++j) {
      int can = 0;
      if (j == i || cudaDeviceCanAccessPeer(&can, devs[j], devs[i]) != cudaSuccess || !can) continue;
      cudaMemAccessDesc d;
      memset(&d, 0, sizeof(d));
      d.location.type = cudaMemLocationTypeDevice; d.location.id = devs[j];
      d.flags = cudaMemAccessFlagsProtReadWrite;
      cudaMemPoolSetAccess(pool, &d, 1);
    }
  }
  cudaGetLastError();
  std::lock_guard<std::mutex> lock(g_comm_mutex);
  auto it = g_comm_cache.find(devs);
  if (it == g_comm_cache.end()) {
    CommSet cs;
    cs.comms.assign(n, nullptr);
    NCK(nccl_api().CommInitAll(cs.comms.data(), n, devs.data()));
    it = g_comm_cache.emplace(devs, cs).first;
  }
  for (int i = 0; i < n; ++i) { ctxs[i]->comm = it->second.comms[i]; ctxs[i]->comm_owned = false; }
  return DPE_OK;
}

void dpe_comm_reset_all(void) {
  std::lock_guard<std::mutex> lock(g_comm_mutex);
  for (auto& kv : g_comm_cache)
    for (ncclComm_t c : kv.second.comms) if (c) nccl_api().CommAbort(c);
  g_comm_cache.clear();
}

int dpe_scene_begin(dpe_ctx* ctx, int n_views, int width, int height, int n_scales) {
  if (!ctx) return DPE_ERR_ARG;
  if (n_views <= 0 || width <= 0 || height <= 0 || n_scales <= 0 || n_scales > 8) FAIL(DPE_ERR_ARG, "bad scene dimensions");
  CK(cudaSetDevice(ctx->device));
  CK(cudaDeviceSynchronize());
  free_scene(ctx);
  ctx->n_views = n_views; ctx->W = width; ctx->H = height; ctx->n_scales = n_scales;
  ctx->sw.assign(n_scales, 0); ctx->sh.assign(n_scales, 0);
  for (int k = 0; k < n_scales; ++k) {
    const int scale_size = 1 << (n_scales - 1 - k);
    const float factor = 1.0f / (float)scale_size;  // DPE.cpp:800-802
    ctx->sw[k] = (int)std::round(width * factor);
    ctx->sh[k] = (int)std::round(height * factor);
  }
  ctx->views.assign(n_views, ViewData());
  for (auto& v : ctx->views) v.scales.assign(n_scales, ScaleImg());
  ctx->n_problems = n_views; ctx->rank = 0; ctx->n_ranks = 1;
  ctx->first_view = 0; ctx->n_local = n_views; ctx->slots_per_rank = n_views;
  ctx->stage_counter = 0; ctx->last_stage_scale = -1; ctx->stage_pending = false; ctx->stage_open = false;
  CK(dmalloc(&ctx->gray_slab, (size_t)width * height * n_views));
  return DPE_OK;
}

int dpe_scene_set_view(dpe_ctx* ctx, int view, const uint8_t* gray, const float K[9], const float R[9],
                       const float t[3], float depth_min, float depth_max) {
  if (!ctx || view < 0 || view >= ctx->n_views || !K || !R || !t) return DPE_ERR_ARG;
  if (ctx->committed) FAIL(DPE_ERR_STATE, "scene already committed");
  CK(cudaSetDevice(ctx->device));
  ViewData& v = ctx->views[view];
  host_cam_set(&v.cam, K, R, t, depth_min, depth_max);
  v.have_cam = true;
  if (gray) {
    // asynchronous on the upload stream (truly so from pinned memory); dpe_scene_commit / dpe_scene_broadcast_images
    // order themselves behind it
    const size_t n = (size_t)ctx->W * ctx->H;
    CK(cudaMemcpyAsync(ctx->gray_slab + n * view, gray, n, cudaMemcpyHostToDevice, ctx->upload_stream));
    v.have_img = true;
  }
  return DPE_OK;
}

int dpe_scene_broadcast_images(dpe_ctx* ctx, int root) {
  if (!ctx || root < 0 || root >= ctx->n_ranks) return DPE_ERR_ARG;
  if (ctx->committed) FAIL(DPE_ERR_STATE, "scene already committed");
  if (ctx->n_ranks > 1 && !ctx->comm) FAIL(DPE_ERR_STATE, "no communicator (dpe_comm_init_rank / dpe_comm_init_all)");
  CK(cudaSetDevice(ctx->device));
  if (ctx->rank == root)
    for (const auto& v : ctx->views) if (!v.have_img) FAIL(DPE_ERR_STATE, "the broadcasting rank lacks an image");
  if (ctx->n_ranks > 1) {
    const size_t n = (size_t)ctx->W * ctx->H * ctx->n_views;
    NCK(nccl_api().Broadcast(ctx->gray_slab, ctx->gray_slab, n, ncclUint8, root, ctx->comm, ctx->upload_stream));
  }
  for (auto& v : ctx->views) v.have_img = true;
  return DPE_OK;
}

int dpe_scene_set_pairs(dpe_ctx* ctx, int view, const int* src_ids, int n_src) {
  if (!ctx || view < 0 || view >= ctx->n_views || n_src < 0 || (n_src > 0 && !src_ids)) return DPE_ERR_ARG;
  if (n_src > DPE_MAX_SRC) FAIL(DPE_ERR_TOO_MANY_IMAGES, "Can't process so much images");  // DPE.cpp:762-765
  for (int i = 0; i < n_src; ++i)
    if (src_ids[i] < 0 || src_ids[i] >= ctx->n_views) FAIL(DPE_ERR_ARG, "source id out of range");
  ctx->views[view].src.assign(src_ids, src_ids + n_src);
  return DPE_OK;
}

// uploads prep arrays of an owned view into its slab slots (after commit)
static int upload_prep(dpe_ctx* ctx, int view, int scale, const uint8_t* edge, const int32_t* label) {
  const int li = view - ctx->first_view;
  ScaleImg& s = ctx->views[view].scales[scale];
  const size_t n = (size_t)ctx->sw[scale] * ctx->sh[scale];
  if (edge) {
    uint8_t* dst = ctx->edge_slab[scale] + (size_t)li * n;
    CK(cudaMemcpyAsync(dst, edge, n, cudaMemcpyHostToDevice, ctx->upload_stream));
    if (scale == 0) {  // edge_low_res of every stage of this view: bit-packed copy for the Bresenham walks
      const int w = ctx->sw[0], h = ctx->sh[0], words = (w + 31) / 32;
      std::vector<uint32_t> bits((size_t)words * h, 0u);
      for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
          if (edge[(size_t)y * w + x]) bits[(size_t)y * words + (x >> 5)] |= 1u << (x & 31);
      uint32_t* bdst = ctx->bits_slab + (size_t)li * words * h;
      CK(cudaMemcpyAsync(bdst, bits.data(), bits.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->upload_stream));
      CK(cudaStreamSynchronize(ctx->upload_stream));  // `bits` is a local
      s.edge_bits = bdst;
    }
    s.edge = dst;
  }
  if (label) {
    int32_t* dst = ctx->label_slab[scale] + (size_t)li * n;
    CK(cudaMemcpyAsync(dst, label, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->upload_stream));
    s.label = dst;
  }
  CK(cudaStreamSynchronize(ctx->upload_stream));
  return DPE_OK;
}

int dpe_scene_set_prep(dpe_ctx* ctx, int view, int scale, const uint8_t* edge, const int32_t* label, size_t n_elems) {
  if (!ctx || view < 0 || view >= ctx->n_views || scale < 0 || scale >= ctx->n_scales) return DPE_ERR_ARG;
  const size_t n = (size_t)ctx->sw[scale] * ctx->sh[scale];
  if (n_elems != n) FAIL(DPE_ERR_ARG, "prep array size does not match the scale (dpe_get_size)");
  if (!ctx->committed) {  // kept on the host until the slabs exist
    ScaleImg& s = ctx->views[view].scales[scale];
    if (edge) s.pending_edge.assign(edge, edge + n);
    if (label) s.pending_label.assign(label, label + n);
    return DPE_OK;
  }
  if (view < ctx->first_view || view >= ctx->first_view + ctx->n_local) return DPE_OK;  // not this rank's view
  CK(cudaSetDevice(ctx->device));
  return upload_prep(ctx, view, scale, edge, label);
}

int dpe_scene_set_shard(dpe_ctx* ctx, int n_problems, int rank, int n_ranks) {
  if (!ctx || n_problems < 1 || n_problems > ctx->n_views || n_ranks < 1 || rank < 0 || rank >= n_ranks) return DPE_ERR_ARG;
  if (ctx->committed) FAIL(DPE_ERR_STATE, "scene already committed");
  ctx->n_problems = n_problems; ctx->rank = rank; ctx->n_ranks = n_ranks;
  dpe_shard_range(n_problems, n_ranks, rank, &ctx->first_view, &ctx->n_local);
  ctx->slots_per_rank = (n_problems + n_ranks - 1) / n_ranks;
  return DPE_OK;
}

int dpe_scene_set_active(dpe_ctx* ctx, int first_view, int count) {
  if (!ctx || first_view < 0 || count < 1) return DPE_ERR_ARG;
  if (ctx->committed) FAIL(DPE_ERR_STATE, "scene already committed");
  if (ctx->n_ranks != 1 || first_view + count > ctx->n_problems) FAIL(DPE_ERR_ARG, "an active sub-range needs an unsharded context and views inside it");
  ctx->first_view = first_view; ctx->n_local = count;
  return DPE_OK;
}

int dpe_view_slot(dpe_ctx* ctx, int view, int* slot) {
  if (!ctx || view < 0 || view >= ctx->n_views || !slot) return DPE_ERR_ARG;
  *slot = slot_of(ctx, view);
  return DPE_OK;
}

int dpe_scene_commit(dpe_ctx* ctx) {
  if (!ctx) return DPE_ERR_ARG;
  if (ctx->committed) return DPE_OK;
  CK(cudaSetDevice(ctx->device));
  const LaunchCfg cfg = cfg_of(ctx);
  const int top = ctx->n_scales - 1;
  for (const auto& v : ctx->views) {
    if (!v.have_cam) FAIL(DPE_ERR_STATE, "view without camera");
    if (!v.have_img) FAIL(DPE_ERR_STATE, "view without image");
  }
  Trace trace(ctx->device);
  CK(cudaStreamSynchronize(ctx->upload_stream));  // images (uploads and / or the broadcast)
  trace("images arrived");
  ctx->scale_arr.assign(ctx->n_scales, nullptr); ctx->scale_tex.assign(ctx->n_scales, 0);
  for (int k = 0; k < ctx->n_scales; ++k) {
    cudaChannelFormatDesc cd = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
    CK(cudaMalloc3DArray(&ctx->scale_arr[k], &cd, make_cudaExtent(ctx->sw[k], ctx->sh[k], ctx->n_views), cudaArrayLayered));
    cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeArray; rd.res.array.array = ctx->scale_arr[k];
    cudaTextureDesc td; memset(&td, 0, sizeof(td));
    // the reference asks for Wrap with unnormalised coordinates, which CUDA turns into
    // Clamp (DPE.cpp:929-933)
    td.addressMode[0] = cudaAddressModeClamp; td.addressMode[1] = cudaAddressModeClamp; td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    CK(cudaCreateTextureObject(&ctx->scale_tex[k], &rd, &td, nullptr));
  }
  ctx->lin_slab.assign(ctx->n_scales, nullptr);
  for (int k = 0; k < ctx->n_scales; ++k)
    CK(dmalloc(&ctx->lin_slab[k], (size_t)ctx->n_views * ctx->sw[k] * ctx->sh[k] * sizeof(float)));
  const size_t P = (size_t)ctx->W * ctx->H;
  for (int vi = 0; vi < ctx->n_views; ++vi) {
    ViewData& v = ctx->views[vi];
    for (int k = 0; k < ctx->n_scales; ++k) v.scales[k].lin = ctx->lin_slab[k] + (size_t)vi * ctx->sw[k] * ctx->sh[k];
    launch_u8_to_f32(ctx->gray_slab + P * vi, v.scales[top].lin, ctx->W * ctx->H, cfg, 0);
    // every level is resized from the full-resolution image (DPE.cpp:798-820)
    for (int k = 0; k < top; ++k)
      launch_resize_linear(v.scales[top].lin, ctx->W, ctx->H, v.scales[k].lin, ctx->sw[k], ctx->sh[k], cfg, 0);
    for (int k = 0; k < ctx->n_scales; ++k) {
      const int w = ctx->sw[k], h = ctx->sh[k];
      cudaMemcpy3DParms cp; memset(&cp, 0, sizeof(cp));
      cp.srcPtr = make_cudaPitchedPtr(v.scales[k].lin, (size_t)w * sizeof(float), w, h);
      cp.dstArray = ctx->scale_arr[k];
      cp.dstPos = make_cudaPos(0, 0, vi);
      cp.extent = make_cudaExtent(w, h, 1);
      cp.kind = cudaMemcpyDeviceToDevice;
      CK(cudaMemcpy3DAsync(&cp, 0));
    }
  }
  trace("pyramids + textures");
  // depth atlases
  const int slots = total_slots(ctx);
  ctx->atlas_front.assign(ctx->n_scales, nullptr); ctx->atlas_back.assign(ctx->n_scales, nullptr);
  for (int k = 0; k < ctx->n_scales; ++k) {
    const size_t bytes = (size_t)slots * ctx->sw[k] * ctx->sh[k] * sizeof(float);
    CK(dmalloc(&ctx->atlas_front[k], bytes)); CK(cudaMemset(ctx->atlas_front[k], 0, bytes));
    CK(dmalloc(&ctx->atlas_back[k], bytes)); CK(cudaMemset(ctx->atlas_back[k], 0, bytes));
  }
  trace("atlases");
  // owned views: prep slabs, carried maps (two buffers), completion events
  const int nl = ctx->n_local > 0 ? ctx->n_local : 1;
  ctx->edge_slab.assign(ctx->n_scales, nullptr); ctx->label_slab.assign(ctx->n_scales, nullptr);
  for (int k = 0; k < ctx->n_scales; ++k) {
    const size_t n = (size_t)ctx->sw[k] * ctx->sh[k];
    CK(dmalloc(&ctx->edge_slab[k], (size_t)nl * n));
    CK(dmalloc(&ctx->label_slab[k], (size_t)nl * n * sizeof(int32_t)));
  }
  CK(dmalloc(&ctx->bits_slab, (size_t)nl * ((ctx->sw[0] + 31) / 32) * ctx->sh[0] * sizeof(uint32_t)));
  // buffer 0 holds the scales top, top-2, .. (sized for the finest), buffer 1 the scales top-1, top-3, .. (sized for
  // the second finest): consecutive scales never share a buffer, which is all a stage at a new scale needs
  for (int b = 0; b < 2; ++b) {
    const size_t Pb = map_stride(ctx, b);
    CK(dmalloc(&ctx->maps_planes[b], (size_t)nl * Pb * sizeof(float4)));
    CK(dmalloc(&ctx->maps_state[b], (size_t)nl * Pb));
    CK(dmalloc(&ctx->maps_selected[b], (size_t)nl * Pb * sizeof(uint32_t)));
  }
  ctx->view_done.assign(ctx->n_local, nullptr);
  for (auto& e : ctx->view_done) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CK(cudaHostAlloc((void**)&ctx->h_rc, (size_t)nl * sizeof(RefConst), cudaHostAllocDefault));
  CK(dmalloc(&ctx->exp_depth, P * sizeof(float))); CK(dmalloc(&ctx->exp_normal, P * 3 * sizeof(float)));
  CK(dmalloc(&ctx->exp_weak, P));
  trace("prep + map slabs");
  // scratch: one set per stream, sized for the finest scale
  const int n_streams = 4;
  ctx->scratch.assign(n_streams, Scratch());
  for (auto& s : ctx->scratch) {
    CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CK(dmalloc(&s.planes, P * sizeof(float4))); CK(dmalloc(&s.costs, P * sizeof(float)));
    CK(dmalloc(&s.selected, P * sizeof(uint32_t))); CK(dmalloc(&s.view_w, P * sizeof(uint4)));
    CK(dmalloc(&s.state, P)); CK(dmalloc(&s.fit_planes, P * sizeof(float4)));
    CK(dmalloc(&s.radius, P * sizeof(int))); CK(dmalloc(&s.edge_neigh, P * 8 * sizeof(short2)));
    CK(dmalloc(&s.complexity, P * sizeof(float))); CK(dmalloc(&s.label_boundary, P * 8 * sizeof(short2)));
    CK(dmalloc(&s.weak_reliable, P)); CK(dmalloc(&s.nearest_strong, P * sizeof(short2)));
    CK(dmalloc(&s.neighbours, P * DPE_NEIGHBOUR_NUM * sizeof(short2)));
    CK(dmalloc(&s.rng, P * sizeof(Xorwow)));
    CK(dmalloc(&s.weak_list, 2 * P * sizeof(int))); CK(dmalloc(&s.weak_count, 4 * sizeof(int)));
    CK(cudaMemset(s.weak_count, 0, 4 * sizeof(int)));
    CK(dmalloc(&s.weak_scan, (size_t)compact_scan_entries(ctx->num_sms) * sizeof(int)));
    CK(cudaMemset(s.view_w, 0, P * sizeof(uint4)));
    CK(cudaMemset(s.radius, 0, P * sizeof(int)));
  }
  CK(dmalloc(&ctx->zero_edge, P)); CK(cudaMemset(ctx->zero_edge, 0, P));
  CK(dmalloc(&ctx->zero_label, P * sizeof(int32_t))); CK(cudaMemset(ctx->zero_label, 0xFF, P * sizeof(int32_t)));
  trace("scratch");
  // relative poses of all (reference, source) pairs of this context's views, from the device (see k_relative_pose)
  {
    std::vector<float> in;
    size_t n_pairs = 0;
    for (int v = 0; v < ctx->n_views; ++v) {
      const ViewData& rv = ctx->views[v];
      for (int sv : rv.src) {
        const HostCam& r = rv.cam; const HostCam& sc = ctx->views[sv].cam;
        for (int i = 0; i < 9; ++i) in.push_back((float)r.R[i]);
        for (int i = 0; i < 3; ++i) in.push_back((float)r.t[i]);
        for (int i = 0; i < 9; ++i) in.push_back((float)sc.R[i]);
        for (int i = 0; i < 3; ++i) in.push_back((float)sc.t[i]);
        for (int i = 0; i < 3; ++i) in.push_back((float)r.C[i]);
        for (int i = 0; i < 3; ++i) in.push_back((float)sc.C[i]);
        ++n_pairs;
      }
    }
    if (n_pairs) {
      float *d_in = nullptr, *d_out = nullptr;
      std::vector<float> out(n_pairs * 13);
      CK(dmalloc(&d_in, in.size() * sizeof(float))); CK(dmalloc(&d_out, out.size() * sizeof(float)));
      CK(cudaMemcpy(d_in, in.data(), in.size() * sizeof(float), cudaMemcpyHostToDevice));
      launch_relative_pose(d_in, d_out, (int)n_pairs, 0);
      CK(cudaGetLastError());
      CK(cudaMemcpy(out.data(), d_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost));
      dfree(d_in); dfree(d_out);
      size_t o = 0;
      for (int v = 0; v < ctx->n_views; ++v) {
        ViewData& rv = ctx->views[v];
        rv.relpose.assign(out.begin() + o, out.begin() + o + rv.src.size() * 13);
        o += rv.src.size() * 13;
      }
    }
  }
  CK(cudaDeviceSynchronize());
  trace("relative poses");
  ctx->committed = true;
  // prep that arrived before the slabs existed
  for (int vi = ctx->first_view; vi < ctx->first_view + ctx->n_local; ++vi)
    for (int k = 0; k < ctx->n_scales; ++k) {
      ScaleImg& s = ctx->views[vi].scales[k];
      if (s.pending_edge.empty() && s.pending_label.empty()) continue;
      if (int rc = upload_prep(ctx, vi, k, s.pending_edge.empty() ? nullptr : s.pending_edge.data(),
                               s.pending_label.empty() ? nullptr : s.pending_label.data()))
        return rc;
    }
  for (auto& v : ctx->views)
    for (auto& s : v.scales) { std::vector<uint8_t>().swap(s.pending_edge); std::vector<int32_t>().swap(s.pending_label); }
  return DPE_OK;
}

}  // extern "C"

static void build_ref_const(const dpe_ctx* ctx, int view, int k, bool geom, RefConst* rc) {
  const ViewData& rv = ctx->views[view];
  const int w = ctx->sw[k], h = ctx->sh[k];
  fold_ref(rv.cam, w, h, ctx->W, ctx->H, view, rc);
  rc->n_src = (int)rv.src.size();
  const size_t P = (size_t)w * h;
  for (int si = 0; si < rc->n_src; ++si) {
    const int sv = rv.src[si];
    SrcConst& sc = rc->src[si];
    fold_pair(rv.cam, ctx->views[sv].cam, w, h, ctx->W, ctx->H, &sc);
    if (rv.relpose.size() >= (size_t)(si + 1) * 13) {
      memcpy(sc.Rrel, &rv.relpose[(size_t)si * 13], 9 * sizeof(float));
      memcpy(sc.trel, &rv.relpose[(size_t)si * 13 + 9], 3 * sizeof(float));
      sc.baseline = rv.relpose[(size_t)si * 13 + 12];  // the device's value (k_relative_pose)
    }
    sc.src_view = sv;
    sc.tex = 0;
    sc.depth = geom ? ctx->atlas_front[k] + (size_t)slot_of(ctx, sv) * P : nullptr;
  }
}

// (re)builds the per-scale initial-state tables for `seed`
static int ensure_rng_tables(dpe_ctx* ctx, uint64_t seed) {
  if (ctx->rng_ready && ctx->rng_seed == seed) return DPE_OK;
  if (ctx->rng_table.empty()) ctx->rng_table.assign(ctx->n_scales, nullptr);
  std::vector<Xorwow> host;
  for (int k = 0; k < ctx->n_scales; ++k) {
    const size_t P = (size_t)ctx->sw[k] * ctx->sh[k];
    host.resize(P);
    xorwow_init_table(seed, ctx->sw[k], ctx->sh[k], host.data());
    if (!ctx->rng_table[k]) CK(dmalloc(&ctx->rng_table[k], P * sizeof(Xorwow)));
    CK(cudaMemcpy(ctx->rng_table[k], host.data(), P * sizeof(Xorwow), cudaMemcpyHostToDevice));
  }
  ctx->rng_seed = seed; ctx->rng_ready = true;
  return DPE_OK;
}

static void fill_args(dpe_ctx* ctx, int view, int k, const dpe_stage_params* p, uint64_t seed, const Scratch& s,
                      KernelParams* KP) {
  memset(KP, 0, sizeof(*KP));
  build_ref_const(ctx, view, k, p && p->geom_consistency, &KP->rc);
  StageArgs& a = KP->a;
  ViewData& v = ctx->views[view];
  a.rc = nullptr;
  a.ref_img = v.scales[k].lin;
  a.W = ctx->sw[k]; a.H = ctx->sh[k];
  a.planes = s.planes; a.costs = s.costs; a.selected = s.selected; a.view_w = s.view_w; a.state = s.state;
  a.fit_planes = s.fit_planes; a.radius = s.radius;
  a.edge = v.scales[k].edge ? v.scales[k].edge : ctx->zero_edge;
  a.edge_low = v.scales[0].edge ? v.scales[0].edge : ctx->zero_edge;  // coarsest scale (DPE.cpp:1036-1045)
  a.low_w = ctx->sw[0]; a.low_h = ctx->sh[0];
  a.edge_low_bits = v.scales[0].edge ? v.scales[0].edge_bits : nullptr; a.low_words = (ctx->sw[0] + 31) / 32;
  a.edge_neigh = s.edge_neigh; a.complexity = s.complexity;
  a.label = v.scales[k].label ? v.scales[k].label : ctx->zero_label;
  a.label_boundary = s.label_boundary; a.weak_reliable = s.weak_reliable; a.nearest_strong = s.nearest_strong;
  a.neighbours = s.neighbours;
  a.rng = s.rng;
  a.ref_race = ctx->ref_race;
  a.snap_planes = s.snap_planes; a.snap_costs = s.snap_costs;
  a.cost_raw = ctx->cost_raw ? 1 : 0;
  a.exact = ctx->exact ? 1 : 0;
  a.variants = ctx->variants;
  {  // a stage truncated after a strong sweep leaves the fit-plane scratch free for the accepted-candidate codes
    const int st = ctx->debug_stop_after;
    a.debug_accept = (st == 2 || st == 5 || st == 8) ? reinterpret_cast<unsigned char*>(s.fit_planes) : nullptr;
  }
  a.weak_list = s.weak_list; a.weak_count = s.weak_count; a.list_stride = ctx->W * ctx->H; a.weak_scan = s.weak_scan;
  a.eval_units = ctx->count_evals ? ctx->d_eval_units : nullptr;
  if (p) {
    a.run_state = p->state; a.geom = p->geom_consistency; a.use_apd = p->use_apd; a.top_k = p->top_k;
    a.weak_peak_radius = p->weak_peak_radius; a.rotate_time = p->rotate_time;
    a.ransac_threshold = p->ransac_threshold; a.geom_factor = p->geom_factor;
  }
  (void)seed;
}

// host-side unpacking of the cloud on a few threads (10^8 points are seconds of work for one)
template <class F>
static void cloud_parallel(size_t n, F fn) {
  const unsigned hw = std::thread::hardware_concurrency();
  const size_t nt = n < (1u << 20) ? 1 : (hw ? (hw > 16 ? 16 : hw) : 4);
  std::vector<std::thread> th;
  for (size_t t = 0; t < nt; ++t) th.emplace_back([=]() { for (size_t i = n * t / nt; i < n * (t + 1) / nt; ++i) fn(i); });
  for (auto& x : th) x.join();
}

// the carried-map buffers of local view li in buffer b
static void map_buffers(dpe_ctx* ctx, int li, int b, float4** planes, uint8_t** state, uint32_t** sel) {
  const size_t P = map_stride(ctx, b);
  *planes = ctx->maps_planes[b] + (size_t)li * P;
  *state = ctx->maps_state[b] + (size_t)li * P;
  *sel = ctx->maps_selected[b] + (size_t)li * P;
}

extern "C" {

int dpe_stage_begin(dpe_ctx* ctx, int k, const dpe_stage_params* p, uint64_t seed) {
  if (!ctx || !p || k < 0 || k >= ctx->n_scales) return DPE_ERR_ARG;
  if (!ctx->committed) FAIL(DPE_ERR_STATE, "scene not committed");
  if (ctx->stage_open) FAIL(DPE_ERR_STATE, "a stage is already running (dpe_stage_end)");
  if (ctx->stage_pending) FAIL(DPE_ERR_STATE, "previous stage not committed (dpe_stage_commit)");
  CK(cudaSetDevice(ctx->device));
  const LaunchCfg cfg = cfg_of(ctx);
  const size_t P = (size_t)ctx->sw[k] * ctx->sh[k];
  const int ns = (int)ctx->scratch.size();
  {
    Trace trace(ctx->device);
    const bool had = ctx->rng_ready && ctx->rng_seed == seed;
    if (int rc = ensure_rng_tables(ctx, seed)) return rc;
    if (!had) trace("rng tables");
  }
  // all views of a scale are layers of one texture (layer = view, linear filter, clamp); the handle sits in a
  // device-wide constant, so nothing of another scale may be in flight (it is not: dpe_stage_end drained it)
  launch_set_scale_tex((unsigned long long)ctx->scale_tex[k], 0);
  if (ctx->ref_race == 2 && !ctx->scratch.empty() && !ctx->scratch[0].snap_planes) {
    const size_t Pmax = (size_t)ctx->sw[ctx->n_scales - 1] * ctx->sh[ctx->n_scales - 1];
    for (auto& s : ctx->scratch) { CK(dmalloc(&s.snap_planes, Pmax * sizeof(float4))); CK(dmalloc(&s.snap_costs, Pmax * sizeof(float))); }
  }
  CK(cudaEventRecord(ctx->ev0, 0));
  for (auto& s : ctx->scratch) CK(cudaStreamWaitEvent(s.stream, ctx->ev0, 0));
  const int stop = ctx->debug_stop_after;
  const bool truncated = stop >= 0 && stop < 11;
  for (int li = 0; li < ctx->n_local; ++li) {
    const int view = ctx->first_view + li;
    ViewData& v = ctx->views[view];
    if (p->state != DPE_FIRST_INIT && v.cur_scale < 0) FAIL(DPE_ERR_STATE, "refine stage before first init");
    const bool prof_view = ctx->profile && li < ctx->profile_views;
    // an external profiler started with --profile-from-start off sees exactly the profiled views
    if (ctx->profile && li == 0) cudaProfilerStart();
    if (ctx->profile && li == ctx->profile_views) cudaProfilerStop();
    const int si = (prof_view || ctx->gauss_seidel) ? 0 : li % ns;
    Scratch& s = ctx->scratch[si];
    KernelParams KP;
    fill_args(ctx, view, k, p, seed, s, &KP);
    StageArgs& a = KP.a;
    a.slot = si;
    // stream-ordered: lands after the previous view's kernels on this stream; from pinned memory, so the host does not
    // wait for that stream to drain (a pageable source would make every view-stage's enqueue block on the one before)
    memcpy(&ctx->h_rc[li], &KP.rc, offsetof(RefConst, src) + (size_t)KP.rc.n_src * sizeof(SrcConst));
    launch_set_ref_const(si, &ctx->h_rc[li], s.stream);
    // outputs: the other map buffer when the scale changes, in place otherwise
    const int out_buf = map_buffer_of_scale(ctx, k);
    float4* new_planes; uint8_t* new_state; uint32_t* new_sel;
    map_buffers(ctx, li, out_buf, &new_planes, &new_state, &new_sel);
    a.prev_planes = v.planes; a.prev_state = v.state; a.prev_selected = v.selected;
    a.prev_W = v.cur_scale >= 0 ? ctx->sw[v.cur_scale] : a.W;
    a.prev_H = v.cur_scale >= 0 ? ctx->sh[v.cur_scale] : a.H;
    a.out_planes = new_planes; a.out_state = new_state; a.out_selected = new_sel;
    // sequential order: publish straight into the committed atlas, so that later views of this stage
    // read it (the reference's depths.dmb files)
    a.atlas_out = (ctx->gauss_seidel ? ctx->atlas_front[k] : ctx->atlas_back[k]) + (size_t)slot_of(ctx, view) * P;

    cudaStream_t st = s.stream;
    // every (view, stage) starts from curand_init(seed, y, x) like the reference (DPE.cu:1020-1033)
    CK(cudaMemcpyAsync(s.rng, ctx->rng_table[k], P * sizeof(Xorwow), cudaMemcpyDeviceToDevice, st));
    // radius_cuda is a fresh allocation per view-stage in the reference and only written for WEAK pixels with
    // >= 3 anchors (DPE.cu:2945-2948 returns before): start every view-stage from zeros, not from the last view's
    if (p->use_apd) CK(cudaMemsetAsync(s.radius, 0, P * sizeof(int), st));
    // in profile mode every launch is bracketed by CUDA events on its own stream and the
    // eval-unit counter is read back after it
    auto L = [&](int cls, void (*fn)(const KernelParams&, const LaunchCfg&, cudaStream_t)) {
      if (!prof_view) { fn(KP, cfg, st); return; }
      unsigned long long u0 = 0, u1 = 0;
      cudaMemcpyAsync(&u0, ctx->d_eval_units, sizeof(u0), cudaMemcpyDeviceToHost, st);
      cudaEventRecord(ctx->pa, st);
      fn(KP, cfg, st);
      cudaEventRecord(ctx->pb, st);
      cudaMemcpyAsync(&u1, ctx->d_eval_units, sizeof(u1), cudaMemcpyDeviceToHost, st);
      cudaStreamSynchronize(st);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ctx->pa, ctx->pb);
      ctx->prof_ms[cls] += ms; ctx->prof_units[cls] += (double)(u1 - u0); ctx->prof_launches[cls]++;
    };
    // steps as oracle/ref_stage_probe.cu numbers them: 0 anchors, 1 init, 2+3i strong sweeps of iteration i,
    // 3+3i fit plane, 4+3i weak sweeps, 11 the tail; dpe_debug_stop_after(n) leaves the scratch arrays as they
    // are after step n (test hook for the per-kernel differential comparison)
    auto on = [&](int step) { return stop < 0 || step <= stop; };
    if (p->state != DPE_FIRST_INIT) L(DPE_K_LOAD, launch_load);
    if (p->use_apd) {
      L(DPE_K_NEAREST, launch_compact_weak);       // the WEAK pixels the previous stage left: input of the next three
      L(DPE_K_EDGE_INFO, launch_edge_info);
      L(DPE_K_NEAREST, launch_nearest_strong);
      L(DPE_K_NEIGHBOURS, launch_gen_neighbours);
      L(DPE_K_NEIGHBOURS, launch_compact_weak);    // again: the anchor search demoted the unreliable ones
    }
    if (on(1)) L(DPE_K_INIT, launch_init);
    for (int it = 0; it < p->max_iterations; ++it) {
      a.iter = it;
      if (on(2 + 3 * it)) {
        a.colour = 0; L(DPE_K_STRONG, launch_strong);
        a.colour = 1; L(DPE_K_STRONG, launch_strong);
      }
      if (p->use_apd) {
        if (on(3 + 3 * it)) L(DPE_K_FIT, launch_fit_plane);
        if (on(4 + 3 * it)) {
          a.colour = 0; L(DPE_K_WEAK, launch_weak);
          a.colour = 1; L(DPE_K_WEAK, launch_weak);
        }
      }
    }
    if (on(11)) {
      L(DPE_K_EXTRACT, launch_extract);
      a.colour = 0; L(DPE_K_MEDIAN, launch_median);
      a.colour = 1; L(DPE_K_MEDIAN, launch_median);
      L(DPE_K_CLASSIFY, launch_classify_refine);
      L(DPE_K_FINISH, launch_finish);
    }
    CK(cudaEventRecord(ctx->view_done[li], st));
    if (!truncated) {  // a truncated stage (test hook) leaves the carried maps as they were
      v.planes = new_planes; v.state = new_state; v.selected = new_sel; v.cur_scale = k; v.cur_buf = out_buf;
    }
  }
  if (ctx->profile && ctx->n_local <= ctx->profile_views) cudaProfilerStop();
  CK(cudaGetLastError());
  // the exchange: view slot i of every rank forms one contiguous chunk of the atlas (slot = i * n_ranks + rank);
  // its in-place all-gather waits only for this rank's view i, so all but the last slot travel under the kernels
  // of the views behind them.  Every rank issues slots_per_rank collectives in the same order; a rank that owns
  // fewer views contributes a slot nobody reads.
  ctx->comm_queued = false;
  if (ctx->n_ranks > 1 && ctx->comm && !truncated) {
    float* atlas = ctx->atlas_back[k];
    for (int i = 0; i < ctx->slots_per_rank; ++i) {
      if (i < ctx->n_local) CK(cudaStreamWaitEvent(ctx->comm_stream, ctx->view_done[i], 0));
      float* chunk = atlas + (size_t)i * ctx->n_ranks * P;
      NCK(nccl_api().AllGather(chunk + (size_t)ctx->rank * P, chunk, P, ncclFloat, ctx->comm, ctx->comm_stream));
    }
    CK(cudaEventRecord(ctx->comm_done, ctx->comm_stream));
    ctx->comm_queued = true;
  }
  ctx->last_stage_scale = k;
  ctx->stage_open = true;
  return DPE_OK;
}

int dpe_stage_wait_view(dpe_ctx* ctx, int view) {
  if (!ctx || view < ctx->first_view || view >= ctx->first_view + ctx->n_local) return DPE_ERR_ARG;
  if (!ctx->stage_open) return DPE_OK;
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventSynchronize(ctx->view_done[view - ctx->first_view]));
  return DPE_OK;
}

int dpe_stage_end(dpe_ctx* ctx) {
  if (!ctx) return DPE_ERR_ARG;
  if (!ctx->stage_open) return DPE_OK;
  CK(cudaSetDevice(ctx->device));
  for (int li = 0; li < ctx->n_local; ++li) CK(cudaStreamWaitEvent(0, ctx->view_done[li], 0));
  CK(cudaEventRecord(ctx->ev_views, 0));  // all of this rank's views done; what follows is the exposed part of the exchange
  if (ctx->comm_queued) CK(cudaStreamWaitEvent(0, ctx->comm_done, 0));
  CK(cudaEventRecord(ctx->ev1, 0));
  CK(cudaEventSynchronize(ctx->ev1));
  CK(cudaGetLastError());
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  ctx->stage_ms += ms;
  CK(cudaEventElapsedTime(&ms, ctx->ev_views, ctx->ev1));
  ctx->comm_ms += ms;
  ctx->stage_counter++;
  ctx->stage_open = false;
  ctx->stage_pending = true;
  return DPE_OK;
}

int dpe_run_stage(dpe_ctx* ctx, int k, const dpe_stage_params* p, uint64_t seed) {
  if (int rc = dpe_stage_begin(ctx, k, p, seed)) return rc;
  return dpe_stage_end(ctx);
}

int dpe_stage_atlas(dpe_ctx* ctx, void** dev_ptr, size_t* slot_bytes, size_t* total_bytes) {
  if (!ctx || ctx->last_stage_scale < 0) return DPE_ERR_ARG;
  const int k = ctx->last_stage_scale;
  const size_t P = (size_t)ctx->sw[k] * ctx->sh[k];
  if (dev_ptr) *dev_ptr = ctx->gauss_seidel ? ctx->atlas_front[k] : ctx->atlas_back[k];
  if (slot_bytes) *slot_bytes = P * sizeof(float);
  if (total_bytes) *total_bytes = (size_t)total_slots(ctx) * P * sizeof(float);
  return DPE_OK;
}

int dpe_stage_commit(dpe_ctx* ctx) {
  if (!ctx) return DPE_ERR_ARG;
  if (ctx->stage_open) if (int rc = dpe_stage_end(ctx)) return rc;
  if (!ctx->stage_pending) return DPE_OK;
  const int k = ctx->last_stage_scale;
  if (!ctx->gauss_seidel) std::swap(ctx->atlas_front[k], ctx->atlas_back[k]);
  ctx->stage_pending = false;
  return DPE_OK;
}

int dpe_set_view_order(dpe_ctx* ctx, int sequential) {
  if (!ctx) return DPE_ERR_ARG;
  if (ctx->stage_pending || ctx->stage_open) FAIL(DPE_ERR_STATE, "stage not committed");
  if (sequential && ctx->n_ranks > 1) FAIL(DPE_ERR_ARG, "sequential view order needs all views on one GPU");
  ctx->gauss_seidel = sequential != 0;
  return DPE_OK;
}

// ---- device fusion (RunFusion, DPE.cpp:1220-1370) ---------------------------------------------
// Fusion reads, for every view, its final (world normal, depth) planes, pixel states and colour image, and keeps a
// mark per pixel.  The maps either come from the host (dpe_fuse_set_view: tests, external maps) or are the ones
// the last stage left on the GPUs (dpe_fuse_prepare: no trip through the host; with several ranks the blocks are
// all-gathered with NCCL so that every rank holds the maps of every view its own views use as sources).
static int fuse_slot(const dpe_ctx* ctx, int view) {  // rank-major: the layout one all-gather of padded blocks produces
  if (ctx->n_ranks == 1 || view >= ctx->n_problems) return view;
  int first = 0, count = 0, r = 0;
  for (; r < ctx->n_ranks; ++r) {
    dpe_shard_range(ctx->n_problems, ctx->n_ranks, r, &first, &count);
    if (view < first + count) break;
  }
  return r * ctx->slots_per_rank + (view - first);
}
static int fuse_total_slots(const dpe_ctx* ctx) {
  return ctx->n_ranks == 1 ? ctx->n_views : ctx->slots_per_rank * ctx->n_ranks;
}
static int fuse_ensure(dpe_ctx* ctx, bool own_maps) {
  const size_t P = (size_t)ctx->W * ctx->H;
  const size_t slots = (size_t)fuse_total_slots(ctx);
  if (own_maps && !ctx->fuse_planes_own) {
    CK(dmalloc(&ctx->fuse_planes_own, slots * P * sizeof(float4)));
    CK(dmalloc(&ctx->fuse_state_own, slots * P));
  }
  if (!ctx->fuse_bgr) {
    CK(dmalloc(&ctx->fuse_bgr, (size_t)ctx->n_views * P * 3));
    CK(cudaMemsetAsync(ctx->fuse_bgr, 0, (size_t)ctx->n_views * P * 3, ctx->upload_stream));
  }
  if (!ctx->fuse_mask) CK(dmalloc(&ctx->fuse_mask, (size_t)ctx->n_views * P * sizeof(uint16_t)));
  if (ctx->fuse_have.empty()) ctx->fuse_have.assign(ctx->n_views, 0);
  return DPE_OK;
}

int dpe_fuse_set_view(dpe_ctx* ctx, int view, const float* depth, const float* normal3, const uint8_t* state, const uint8_t* bgr) {
  if (!ctx || view < 0 || view >= ctx->n_views || !depth || !normal3 || !state || !bgr) return DPE_ERR_ARG;
  if (ctx->n_ranks != 1) FAIL(DPE_ERR_STATE, "host-provided fusion maps need an unsharded context");
  CK(cudaSetDevice(ctx->device));
  if (int rc = fuse_ensure(ctx, true)) return rc;
  ctx->fuse_resident = false;
  const size_t P = (size_t)ctx->W * ctx->H;
  std::vector<float4> h(P);
  for (size_t i = 0; i < P; ++i) h[i] = make_float4(normal3[3 * i], normal3[3 * i + 1], normal3[3 * i + 2], depth[i]);
  CK(cudaMemcpy(ctx->fuse_planes_own + (size_t)view * P, h.data(), P * sizeof(float4), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->fuse_state_own + (size_t)view * P, state, P, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->fuse_bgr + (size_t)view * P * 3, bgr, P * 3, cudaMemcpyHostToDevice));
  ctx->fuse_have[view] = 1;
  return DPE_OK;
}

int dpe_fuse_set_color(dpe_ctx* ctx, int view, const uint8_t* bgr) {
  if (!ctx || view < 0 || view >= ctx->n_views || !bgr) return DPE_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (int rc = fuse_ensure(ctx, false)) return rc;
  const size_t P = (size_t)ctx->W * ctx->H;
  CK(cudaMemcpyAsync(ctx->fuse_bgr + (size_t)view * P * 3, bgr, P * 3, cudaMemcpyHostToDevice, ctx->upload_stream));
  return DPE_OK;
}

int dpe_fuse_set_block(dpe_ctx* ctx, int view, const uint8_t* mask) {
  if (!ctx || view < 0 || view >= ctx->n_views || !mask) return DPE_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (int rc = fuse_ensure(ctx, false)) return rc;
  const size_t P = (size_t)ctx->W * ctx->H;
  if (!ctx->fuse_block) {
    CK(dmalloc(&ctx->fuse_block, (size_t)ctx->n_views * P));
    ctx->fuse_block_have.assign(ctx->n_views, 0);
  }
  // pageable or pinned source alike: the copy is complete (or staged) when the call returns
  CK(cudaMemcpyAsync(ctx->fuse_block + (size_t)view * P, mask, P, cudaMemcpyHostToDevice, ctx->upload_stream));
  CK(cudaStreamSynchronize(ctx->upload_stream));
  ctx->fuse_block_have[view] = 1;
  return DPE_OK;
}

int dpe_fuse_broadcast_colors(dpe_ctx* ctx, int root) {
  if (!ctx || root < 0 || root >= ctx->n_ranks) return DPE_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (int rc = fuse_ensure(ctx, false)) return rc;
  if (ctx->n_ranks > 1) {
    if (!ctx->comm) FAIL(DPE_ERR_STATE, "no communicator");
    NCK(nccl_api().Broadcast(ctx->fuse_bgr, ctx->fuse_bgr, (size_t)ctx->n_views * ctx->W * ctx->H * 3, ncclUint8, root, ctx->comm, ctx->upload_stream));
  }
  CK(cudaStreamSynchronize(ctx->upload_stream));
  return DPE_OK;
}

int dpe_fuse_prepare(dpe_ctx* ctx) {
  if (!ctx) return DPE_ERR_ARG;
  if (!ctx->committed) FAIL(DPE_ERR_STATE, "scene not committed");
  if (ctx->stage_open) if (int rc = dpe_stage_end(ctx)) return rc;
  CK(cudaSetDevice(ctx->device));
  const int top = ctx->n_scales - 1;
  for (int li = 0; li < ctx->n_local; ++li)
    if (ctx->views[ctx->first_view + li].cur_scale != top) FAIL(DPE_ERR_STATE, "fusion needs the maps of the finest scale (run the whole schedule first)");
  const size_t P = (size_t)ctx->W * ctx->H;
  const int b = map_buffer_of_scale(ctx, top);  // = 0: stride P per view, the local block is contiguous
  if (ctx->n_ranks == 1 && ctx->first_view == 0) {
    if (int rc = fuse_ensure(ctx, false)) return rc;
    ctx->fuse_planes = ctx->maps_planes[b]; ctx->fuse_state = ctx->maps_state[b];  // the carried maps themselves
  } else {
    if (int rc = fuse_ensure(ctx, true)) return rc;
    if (ctx->n_ranks > 1 && !ctx->comm) FAIL(DPE_ERR_STATE, "no communicator");
    const size_t spr = (size_t)ctx->slots_per_rank;
    float4* mine_p = ctx->fuse_planes_own + (size_t)ctx->rank * spr * P;
    uint8_t* mine_s = ctx->fuse_state_own + (size_t)ctx->rank * spr * P;
    CK(cudaMemcpyAsync(mine_p, ctx->maps_planes[b], (size_t)ctx->n_local * P * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->comm_stream));
    CK(cudaMemcpyAsync(mine_s, ctx->maps_state[b], (size_t)ctx->n_local * P, cudaMemcpyDeviceToDevice, ctx->comm_stream));
    if (ctx->n_ranks > 1) {
      NCK(nccl_api().AllGather(mine_p, ctx->fuse_planes_own, spr * P * 4, ncclFloat, ctx->comm, ctx->comm_stream));
      NCK(nccl_api().AllGather(mine_s, ctx->fuse_state_own, spr * P, ncclUint8, ctx->comm, ctx->comm_stream));
    }
    CK(cudaStreamSynchronize(ctx->comm_stream));
    ctx->fuse_planes = ctx->fuse_planes_own; ctx->fuse_state = ctx->fuse_state_own;
  }
  ctx->fuse_resident = true;
  for (int v = 0; v < ctx->n_views; ++v) ctx->fuse_have[v] = v < ctx->n_problems ? 1 : 0;
  return DPE_OK;
}

int dpe_fuse_run(dpe_ctx* ctx, int first_view, int count, size_t* n_points) {
  if (!ctx || !n_points || first_view < 0 || count < 0 || first_view + count > ctx->n_views) return DPE_ERR_ARG;
  if (ctx->fuse_have.empty()) FAIL(DPE_ERR_STATE, "no fusion inputs (dpe_fuse_set_view / dpe_fuse_prepare)");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->upload_stream));
  if (!ctx->fuse_resident) { ctx->fuse_planes = ctx->fuse_planes_own; ctx->fuse_state = ctx->fuse_state_own; }
  const int V = ctx->n_views, W = ctx->W, H = ctx->H;
  const size_t P = (size_t)W * H;
  std::vector<FuseView> hv(V);
  CK(cudaMemset(ctx->fuse_mask, 0, (size_t)V * P * sizeof(uint16_t)));
  for (int v = 0; v < V; ++v) {
    FuseView& fv = hv[v];
    memset(&fv, 0, sizeof(fv));
    if (ctx->fuse_have[v]) {
      const size_t slot = ctx->fuse_resident ? (size_t)fuse_slot(ctx, v) : (size_t)v;
      fv.planes = ctx->fuse_planes + slot * P; fv.state = ctx->fuse_state + slot * P;
    }
    fv.bgr = ctx->fuse_bgr + (size_t)v * P * 3; fv.mask = ctx->fuse_mask + (size_t)v * P;
    if (ctx->fuse_block && ctx->fuse_block_have[v]) fv.block = ctx->fuse_block + (size_t)v * P;
    const HostCam& c = ctx->views[v].cam;
    for (int i = 0; i < 9; ++i) { fv.K[i] = (float)c.K[i]; fv.R[i] = (float)c.R[i]; }
    for (int i = 0; i < 3; ++i) fv.t[i] = (float)c.t[i];
    // camera centre as the reference's fusion computes it: float arithmetic (DPE.cpp:1170-1194)
    for (int j = 0; j < 3; ++j) fv.C[j] = -(fv.R[0 + j] * fv.t[0] + fv.R[3 + j] * fv.t[1] + fv.R[6 + j] * fv.t[2]);
  }
  FuseView* dv; CK(dmalloc(&dv, V * sizeof(FuseView)));
  CK(cudaMemcpy(dv, hv.data(), V * sizeof(FuseView), cudaMemcpyHostToDevice));
  // the cloud is assembled on the device (ordered append, the running total stays on the device) and copied out once;
  // capacity: every pixel of every fused view, or what the pool still gives
  FusedPointDev *pts = nullptr, *cloud_dev = nullptr; uint8_t* accept = nullptr; int* warp_counts = nullptr;
  unsigned long long* running = nullptr;
  CK(dmalloc(&pts, P * sizeof(FusedPointDev)));
  CK(dmalloc(&accept, P)); CK(dmalloc(&warp_counts, (size_t)fuse_append_blocks(ctx->num_sms) * 8 * sizeof(int)));
  CK(dmalloc(&running, 3 * sizeof(unsigned long long)));
  CK(cudaMemset(running, 0, 3 * sizeof(unsigned long long)));
  int n_fuse = 0;
  for (int i = first_view; i < first_view + count; ++i) if (hv[i].planes) ++n_fuse;
  unsigned long long capacity = (unsigned long long)(n_fuse > 0 ? n_fuse : 1) * P;
  while (capacity >= P && dmalloc(&cloud_dev, capacity * sizeof(FusedPointDev)) != cudaSuccess) {
    cudaGetLastError();
    cloud_dev = nullptr;
    capacity /= 2;
  }
  if (!cloud_dev) FAIL(DPE_ERR_CUDA, "no device memory for the fused cloud");
  ctx->cloud.clear();
  // views in order: a view sees every mark of the views before it
  for (int i = first_view; i < first_view + count; ++i) {
    if (!hv[i].planes) continue;
    FuseSrcList sl; memset(&sl, 0, sizeof(sl));
    const std::vector<int>& src = ctx->views[i].src;
    sl.n = (int)src.size();
    for (int j = 0; j < sl.n; ++j) sl.id[j] = ctx->fuse_have[src[j]] ? src[j] : -1;
    launch_fuse_view(dv, i, sl, W, H, pts, accept, ctx->num_sms, 0);
    launch_fuse_append(pts, accept, (int)P, warp_counts, running, capacity, cloud_dev, ctx->num_sms, 0);
    ctx->launches += 4;
  }
  CK(cudaGetLastError());
  unsigned long long h_run[3] = {0, 0, 0};
  CK(cudaMemcpy(h_run, running, sizeof(h_run), cudaMemcpyDeviceToHost));
  ctx->cloud.resize((size_t)h_run[0]);
  if (h_run[0]) CK(cudaMemcpy(ctx->cloud.data(), cloud_dev, (size_t)h_run[0] * sizeof(FusedPointDev), cudaMemcpyDeviceToHost));
  dfree(dv); dfree(pts); dfree(accept); dfree(warp_counts); dfree(running); dfree(cloud_dev);
  if (h_run[2]) FAIL(DPE_ERR_CUDA, "the fused cloud did not fit into device memory");
  *n_points = ctx->cloud.size();
  return DPE_OK;
}

int dpe_fuse_get(dpe_ctx* ctx, float* xyz, uint8_t* bgr) {
  if (!ctx || !xyz || !bgr) return DPE_ERR_ARG;
  const FusedPointDev* c = ctx->cloud.data();
  cloud_parallel(ctx->cloud.size(), [=](size_t i) {
    const FusedPointDev& p = c[i];
    xyz[3 * i] = p.x; xyz[3 * i + 1] = p.y; xyz[3 * i + 2] = p.z;
    bgr[3 * i] = (uint8_t)(p.bgr & 255u); bgr[3 * i + 1] = (uint8_t)((p.bgr >> 8) & 255u); bgr[3 * i + 2] = (uint8_t)((p.bgr >> 16) & 255u);
  });
  return DPE_OK;
}

int dpe_fuse_get_ply_records(dpe_ctx* ctx, void* out) {
  if (!ctx || !out) return DPE_ERR_ARG;
  const FusedPointDev* c = ctx->cloud.data();
  uint8_t* o = (uint8_t*)out;
  cloud_parallel(ctx->cloud.size(), [=](size_t i) {
    const FusedPointDev& p = c[i];
    uint8_t* r = o + 15 * i;
    memcpy(r, &p.x, 12);
    r[12] = (uint8_t)(p.bgr & 255u); r[13] = (uint8_t)((p.bgr >> 8) & 255u); r[14] = (uint8_t)((p.bgr >> 16) & 255u);
  });
  return DPE_OK;
}

int dpe_set_cost_arithmetic(dpe_ctx* ctx, int mode) {
  if (!ctx || mode < 0 || mode > 2) return DPE_ERR_ARG;
  ctx->cost_raw = mode != DPE_COST_CENTRED;
  ctx->exact = mode == DPE_COST_REFERENCE_EXACT;
  return DPE_OK;
}

int dpe_debug_set_variants(dpe_ctx* ctx, int mask) {
  if (!ctx) return DPE_ERR_ARG;
  ctx->variants = mask;
  return DPE_OK;
}

int dpe_set_reference_race(dpe_ctx* ctx, int on) {
  if (!ctx) return DPE_ERR_ARG;
  ctx->ref_race = on == 2 ? 2 : (on != 0);
  return DPE_OK;
}

// scratch buffers of the last stage run for `view` (test hook; only meaningful when the view ran on the
// first stream: sequential view order, or a shard of one view)
int dpe_debug_stop_after(dpe_ctx* ctx, int step) {
  if (!ctx || step < -1 || step > 11) return DPE_ERR_ARG;
  ctx->debug_stop_after = step;
  // a truncated stage has written neither the carried maps nor the atlas: it may be run again without a commit
  ctx->stage_pending = false;
  return DPE_OK;
}

int dpe_debug_set_maps(dpe_ctx* ctx, int view, int k, const float* planes4, const uint8_t* state, const uint32_t* selected,
                       const float* atlas_depth) {
  if (!ctx || view < 0 || view >= ctx->n_views || k < 0 || k >= ctx->n_scales) return DPE_ERR_ARG;
  if (!ctx->committed) FAIL(DPE_ERR_STATE, "scene not committed");
  if (planes4 && (!state || !selected)) return DPE_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  const size_t P = (size_t)ctx->sw[k] * ctx->sh[k];
  ViewData& v = ctx->views[view];
  if (planes4) {
    if (view < ctx->first_view || view >= ctx->first_view + ctx->n_local) FAIL(DPE_ERR_ARG, "not a view of this context");
    v.cur_buf = map_buffer_of_scale(ctx, k);
    map_buffers(ctx, view - ctx->first_view, v.cur_buf, &v.planes, &v.state, &v.selected);
    v.cur_scale = k;
    CK(cudaMemcpy(v.planes, planes4, P * sizeof(float4), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(v.state, state, P, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(v.selected, selected, P * sizeof(uint32_t), cudaMemcpyHostToDevice));
  }
  if (atlas_depth) CK(cudaMemcpy(ctx->atlas_front[k] + (size_t)slot_of(ctx, view) * P, atlas_depth, P * sizeof(float), cudaMemcpyHostToDevice));
  return DPE_OK;
}

int dpe_debug_read(dpe_ctx* ctx, int what, void* out, size_t bytes) {
  if (!ctx || !out || ctx->scratch.empty() || ctx->last_stage_scale < 0) return DPE_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  const Scratch& s = ctx->scratch[0];
  const size_t P = (size_t)ctx->sw[ctx->last_stage_scale] * ctx->sh[ctx->last_stage_scale];
  const void* src = nullptr; size_t n = 0;
  switch (what) {
    case 0: src = s.neighbours; n = P * DPE_NEIGHBOUR_NUM * sizeof(short2); break;
    case 1: src = s.fit_planes; n = P * sizeof(float4); break;
    case 2: src = s.radius; n = P * sizeof(int); break;
    case 3: src = s.costs; n = P * sizeof(float); break;
    case 4: src = s.weak_reliable; n = P; break;
    case 5: src = s.nearest_strong; n = P * sizeof(short2); break;
    case 6: src = s.complexity; n = P * sizeof(float); break;
    case 7: src = s.planes; n = P * sizeof(float4); break;
    case 8: src = s.selected; n = P * sizeof(uint32_t); break;
    case 9: src = s.state; n = P; break;
    case 11: src = s.fit_planes; n = P; break;  // accepted-candidate codes of a stage truncated after a strong sweep
    case 12: src = s.edge_neigh; n = P * 8 * sizeof(short2); break;
    case 13: src = s.label_boundary; n = P * 8 * sizeof(short2); break;
    case 10: src = ctx->views[ctx->first_view].scales[ctx->last_stage_scale].lin; n = P * sizeof(float); break;
    default: return DPE_ERR_ARG;
  }
  if (bytes < n) return DPE_ERR_ARG;
  CK(cudaMemcpy(out, src, n, cudaMemcpyDeviceToHost));
  return DPE_OK;
}

int dpe_cost_eval(dpe_ctx* ctx, int view, int k, int n_pix, const int* xy, const float* planes, int mode, float* out) {
  if (!ctx || view < 0 || view >= ctx->n_views || k < 0 || k >= ctx->n_scales || n_pix <= 0 || !xy || !planes || !out)
    return DPE_ERR_ARG;
  if (!ctx->committed) FAIL(DPE_ERR_STATE, "scene not committed");
  CK(cudaSetDevice(ctx->device));
  KernelParams KP;
  fill_args(ctx, view, k, nullptr, 0, ctx->scratch[0], &KP);
  const int N = KP.rc.n_src;
  int* d_xy; float4* d_pl; float* d_out;
  CK(dmalloc(&d_xy, (size_t)n_pix * 2 * sizeof(int))); CK(dmalloc(&d_pl, (size_t)n_pix * sizeof(float4)));
  CK(dmalloc(&d_out, (size_t)n_pix * N * sizeof(float)));
  CK(cudaMemcpy(d_xy, xy, (size_t)n_pix * 2 * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_pl, planes, (size_t)n_pix * sizeof(float4), cudaMemcpyHostToDevice));
  CK(cudaDeviceSynchronize());
  launch_set_scale_tex((unsigned long long)ctx->scale_tex[k], 0);
  KP.a.slot = DPE_RC_SLOTS - 1;
  launch_set_ref_const(KP.a.slot, &KP.rc, 0);
  launch_cost_eval(KP, n_pix, d_xy, d_pl, mode, 0ull, d_out, cfg_of(ctx), 0);
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, d_out, (size_t)n_pix * N * sizeof(float), cudaMemcpyDeviceToHost));
  dfree(d_xy); dfree(d_pl); dfree(d_out);
  return DPE_OK;
}

int dpe_geom_eval(dpe_ctx* ctx, int view, int k, int n_pix, const int* xy, const float* planes, float* out) {
  if (!ctx || view < 0 || view >= ctx->n_views || k < 0 || k >= ctx->n_scales || n_pix <= 0 || !xy || !planes || !out)
    return DPE_ERR_ARG;
  if (!ctx->committed) FAIL(DPE_ERR_STATE, "scene not committed");
  CK(cudaSetDevice(ctx->device));
  dpe_stage_params p; memset(&p, 0, sizeof(p)); p.geom_consistency = 1;
  KernelParams KP;
  fill_args(ctx, view, k, &p, 0, ctx->scratch[0], &KP);
  const int N = KP.rc.n_src;
  int* d_xy; float4* d_pl; float* d_out;
  CK(dmalloc(&d_xy, (size_t)n_pix * 2 * sizeof(int))); CK(dmalloc(&d_pl, (size_t)n_pix * sizeof(float4)));
  CK(dmalloc(&d_out, (size_t)n_pix * N * sizeof(float)));
  CK(cudaMemcpy(d_xy, xy, (size_t)n_pix * 2 * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_pl, planes, (size_t)n_pix * sizeof(float4), cudaMemcpyHostToDevice));
  CK(cudaDeviceSynchronize());
  KP.a.slot = DPE_RC_SLOTS - 1;
  launch_set_ref_const(KP.a.slot, &KP.rc, 0);
  launch_geom_eval(KP, n_pix, d_xy, d_pl, d_out, cfg_of(ctx), 0);
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, d_out, (size_t)n_pix * N * sizeof(float), cudaMemcpyDeviceToHost));
  dfree(d_xy); dfree(d_pl); dfree(d_out);
  return DPE_OK;
}

int dpe_get_size(dpe_ctx* ctx, int k, int* width, int* height) {
  if (!ctx || k < 0 || k >= ctx->n_scales) return DPE_ERR_ARG;
  if (width) *width = ctx->sw[k];
  if (height) *height = ctx->sh[k];
  return DPE_OK;
}

int dpe_get_maps(dpe_ctx* ctx, int view, float* depth, float* normal3, uint8_t* state, uint32_t* selected) {
  if (!ctx || view < 0 || view >= ctx->n_views) return DPE_ERR_ARG;
  ViewData& v = ctx->views[view];
  if (v.cur_scale < 0 || !v.planes) FAIL(DPE_ERR_STATE, "view has no result on this context");
  CK(cudaSetDevice(ctx->device));
  const size_t P = (size_t)ctx->sw[v.cur_scale] * ctx->sh[v.cur_scale];
  if (depth || normal3) {
    std::vector<float4> h(P);
    CK(cudaMemcpy(h.data(), v.planes, P * sizeof(float4), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < P; ++i) {
      if (depth) depth[i] = h[i].w;
      if (normal3) { normal3[3 * i] = h[i].x; normal3[3 * i + 1] = h[i].y; normal3[3 * i + 2] = h[i].z; }
    }
  }
  if (state) CK(cudaMemcpy(state, v.state, P, cudaMemcpyDeviceToHost));
  if (selected) CK(cudaMemcpy(selected, v.selected, P * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return DPE_OK;
}

// the arrays the .npy writers hold, packed on the device and copied out on the copy stream (beside running stages)
int dpe_export_view(dpe_ctx* ctx, int view, float* depth, float* normal3, int8_t* weak) {
  if (!ctx || view < 0 || view >= ctx->n_views) return DPE_ERR_ARG;
  ViewData& v = ctx->views[view];
  if (v.cur_scale < 0 || !v.planes) FAIL(DPE_ERR_STATE, "view has no result on this context");
  CK(cudaSetDevice(ctx->device));
  if (ctx->stage_open) CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->view_done[view - ctx->first_view], 0));
  const size_t P = (size_t)ctx->sw[v.cur_scale] * ctx->sh[v.cur_scale];
  launch_export(v.planes, v.state, depth ? ctx->exp_depth : nullptr, normal3 ? ctx->exp_normal : nullptr,
                weak ? ctx->exp_weak : nullptr, (int)P, cfg_of(ctx), ctx->copy_stream);
  if (depth) CK(cudaMemcpyAsync(depth, ctx->exp_depth, P * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_stream));
  if (normal3) CK(cudaMemcpyAsync(normal3, ctx->exp_normal, P * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_stream));
  if (weak) CK(cudaMemcpyAsync(weak, ctx->exp_weak, P, cudaMemcpyDeviceToHost, ctx->copy_stream));
  CK(cudaStreamSynchronize(ctx->copy_stream));
  CK(cudaGetLastError());
  return DPE_OK;
}

// viz=True: the three colour images of a view's current maps, rendered into device staging (valid until the next call)
int dpe_viz_render(dpe_ctx* ctx, int view, void** bgr_depth, void** bgr_normal, void** bgr_weak, int* width, int* height) {
  if (!ctx || view < 0 || view >= ctx->n_views || !bgr_depth || !bgr_normal || !bgr_weak) return DPE_ERR_ARG;
  ViewData& v = ctx->views[view];
  if (v.cur_scale < 0 || !v.planes) FAIL(DPE_ERR_STATE, "view has no result on this context");
  CK(cudaSetDevice(ctx->device));
  const size_t Pf = (size_t)ctx->W * ctx->H;
  if (!ctx->viz_bgr) CK(dmalloc(&ctx->viz_bgr, 3 * Pf * 3));
  if (ctx->stage_open) CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->view_done[view - ctx->first_view], 0));
  const int k = v.cur_scale;
  const size_t P = (size_t)ctx->sw[k] * ctx->sh[k];
  uint8_t* d0 = ctx->viz_bgr; uint8_t* d1 = d0 + Pf * 3; uint8_t* d2 = d1 + Pf * 3;
  // GetDepthMin / GetDepthMax of the view-stage: the PatchMatch range (cam file min * 0.6, max * 1.2, DPE.cpp:788-789)
  launch_viz(v.planes, v.state, v.cam.depth_min * 0.6f, v.cam.depth_max * 1.2f, d0, d1, d2, (int)P, cfg_of(ctx), ctx->copy_stream);
  CK(cudaStreamSynchronize(ctx->copy_stream));
  CK(cudaGetLastError());
  *bgr_depth = d0; *bgr_normal = d1; *bgr_weak = d2;
  if (width) *width = ctx->sw[k];
  if (height) *height = ctx->sh[k];
  return DPE_OK;
}

int dpe_set_count_evals(dpe_ctx* ctx, int on) {
  if (!ctx) return DPE_ERR_ARG;
  ctx->count_evals = on != 0;
  return DPE_OK;
}

double dpe_eval_units(dpe_ctx* ctx) {
  if (!ctx) return 0.0;
  cudaSetDevice(ctx->device);
  unsigned long long h = 0;
  if (cudaMemcpy(&h, ctx->d_eval_units, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return -1.0;
  return (double)h;
}

double dpe_stage_gpu_ms(dpe_ctx* ctx) { return ctx ? ctx->stage_ms : 0.0; }
double dpe_stage_comm_ms(dpe_ctx* ctx) { return ctx ? ctx->comm_ms : 0.0; }

int dpe_set_profile(dpe_ctx* ctx, int on) {
  if (!ctx) return DPE_ERR_ARG;
  ctx->profile = on != 0;
  ctx->profile_views = on;
  if (on) {
    ctx->count_evals = true;
    for (int i = 0; i < DPE_N_KERNEL_CLASSES; ++i) { ctx->prof_ms[i] = 0; ctx->prof_units[i] = 0; ctx->prof_launches[i] = 0; }
  }
  return DPE_OK;
}

int dpe_get_profile(dpe_ctx* ctx, double* ms, double* units, long long* launches) {
  if (!ctx) return DPE_ERR_ARG;
  for (int i = 0; i < DPE_N_KERNEL_CLASSES; ++i) {
    if (ms) ms[i] = ctx->prof_ms[i];
    if (units) units[i] = ctx->prof_units[i];
    if (launches) launches[i] = ctx->prof_launches[i];
  }
  return DPE_OK;
}

int dpe_bench_ncc(dpe_ctx* ctx, int view, int variant, int n_cand, int reps, double* units_per_s, double* checksum) {
  if (!ctx || view < 0 || view >= ctx->n_views || n_cand <= 0 || reps <= 0 || !units_per_s) return DPE_ERR_ARG;
  ViewData& v = ctx->views[view];
  if (v.cur_scale < 0 || !v.planes) FAIL(DPE_ERR_STATE, "view has no result on this context");
  CK(cudaSetDevice(ctx->device));
  const int k = v.cur_scale;
  KernelParams KP;
  fill_args(ctx, view, k, nullptr, 0, ctx->scratch[0], &KP);
  const size_t P = (size_t)ctx->sw[k] * ctx->sh[k];
  float* out = ctx->scratch[0].costs;
  const LaunchCfg cfg = cfg_of(ctx);
  CK(cudaDeviceSynchronize());
  launch_set_scale_tex((unsigned long long)ctx->scale_tex[k], 0);
  KP.a.slot = DPE_RC_SLOTS - 1;
  launch_set_ref_const(KP.a.slot, &KP.rc, 0);
  launch_ncc_bench(KP, v.planes, n_cand, variant, out, cfg, 0);
  CK(cudaEventRecord(ctx->ev0, 0));
  for (int r = 0; r < reps; ++r) launch_ncc_bench(KP, v.planes, n_cand, variant, out, cfg, 0);
  CK(cudaEventRecord(ctx->ev1, 0));
  CK(cudaEventSynchronize(ctx->ev1));
  CK(cudaGetLastError());
  float ms = 0.f; CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  *units_per_s = (double)(P / 2) * n_cand * KP.rc.n_src * reps / (ms * 1e-3);
  if (checksum) {
    std::vector<float> h(P);
    CK(cudaMemcpy(h.data(), out, P * sizeof(float), cudaMemcpyDeviceToHost));
    double s = 0; for (size_t i = 0; i < P; ++i) if (((i % ctx->sw[k]) + (i / ctx->sw[k])) % 2 == 0 && h[i] == h[i]) s += h[i];
    *checksum = s;
  }
  return DPE_OK;
}

int dpe_probe_tex_rate(dpe_ctx* ctx, int width, int height, int iters, double* taps_per_s) {
  if (!ctx || width < 64 || height < 64 || iters <= 0 || !taps_per_s) return DPE_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  std::vector<float> h((size_t)width * height);
  uint32_t s = 12345u;
  for (auto& x : h) { s = s * 1664525u + 1013904223u; x = (float)(s >> 24); }
  cudaArray_t arr; cudaTextureObject_t tex;
  cudaChannelFormatDesc cd = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
  CK(cudaMallocArray(&arr, &cd, width, height));
  CK(cudaMemcpy2DToArray(arr, 0, 0, h.data(), (size_t)width * 4, (size_t)width * 4, height, cudaMemcpyHostToDevice));
  cudaResourceDesc rd; memset(&rd, 0, sizeof(rd)); rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
  cudaTextureDesc td; memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModeLinear;
  td.readMode = cudaReadModeElementType;
  CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  float* sink; CK(dmalloc(&sink, 4));
  const int blocks = ctx->num_sms * 8, threads = 256;
  const LaunchCfg cfg = cfg_of(ctx);
  launch_probe_tex((unsigned long long)tex, width, height, iters / 4 + 1, sink, blocks, threads, cfg, 0);  // warm-up
  CK(cudaEventRecord(ctx->ev0, 0));
  launch_probe_tex((unsigned long long)tex, width, height, iters, sink, blocks, threads, cfg, 0);
  CK(cudaEventRecord(ctx->ev1, 0));
  CK(cudaEventSynchronize(ctx->ev1));
  float ms = 0.f; CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  *taps_per_s = (double)blocks * threads * (double)iters * 36.0 / (ms * 1e-3);
  dfree(sink); cudaDestroyTextureObject(tex); cudaFreeArray(arr);
  return DPE_OK;
}

int dpe_probe_tex_pattern(dpe_ctx* ctx, int fmt, int layout, int width, int height, int iters, const float m[4],
                          int threads, int blocks_per_sm, double* taps_per_s) {
  if (!ctx || width < 128 || height < 128 || iters <= 0 || !taps_per_s || !m || fmt < 0 || fmt > 2) return DPE_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)width * height;
  std::vector<float> h(n);
  uint32_t s = 12345u;
  for (auto& x : h) { s = s * 1664525u + 1013904223u; x = (float)(s >> 24); }
  cudaArray_t arr; cudaTextureObject_t tex;
  cudaTextureDesc td; memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModeLinear;
  td.readMode = cudaReadModeElementType;
  if (fmt == 0) {
    cudaChannelFormatDesc cd = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
    CK(cudaMallocArray(&arr, &cd, width, height));
    CK(cudaMemcpy2DToArray(arr, 0, 0, h.data(), (size_t)width * 4, (size_t)width * 4, height, cudaMemcpyHostToDevice));
  } else if (fmt == 1) {
    std::vector<__half> hh(n);
    for (size_t i = 0; i < n; ++i) hh[i] = __float2half(h[i]);
    cudaChannelFormatDesc cd = cudaCreateChannelDescHalf();
    CK(cudaMallocArray(&arr, &cd, width, height));
    CK(cudaMemcpy2DToArray(arr, 0, 0, hh.data(), (size_t)width * 2, (size_t)width * 2, height, cudaMemcpyHostToDevice));
  } else {
    std::vector<uint8_t> hb(n);
    for (size_t i = 0; i < n; ++i) hb[i] = (uint8_t)h[i];
    cudaChannelFormatDesc cd = cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned);
    CK(cudaMallocArray(&arr, &cd, width, height));
    CK(cudaMemcpy2DToArray(arr, 0, 0, hb.data(), (size_t)width, (size_t)width, height, cudaMemcpyHostToDevice));
    td.readMode = cudaReadModeNormalizedFloat;
  }
  cudaResourceDesc rd; memset(&rd, 0, sizeof(rd)); rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
  CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  float* sink; CK(dmalloc(&sink, 4));
  const int blocks = ctx->num_sms * blocks_per_sm;
  const LaunchCfg cfg = cfg_of(ctx);
  launch_probe_tex_pattern((unsigned long long)tex, width, height, iters / 4 + 1, layout, m, sink, blocks, threads, cfg, 0);
  CK(cudaEventRecord(ctx->ev0, 0));
  launch_probe_tex_pattern((unsigned long long)tex, width, height, iters, layout, m, sink, blocks, threads, cfg, 0);
  CK(cudaEventRecord(ctx->ev1, 0));
  CK(cudaEventSynchronize(ctx->ev1));
  float ms = 0.f; CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  *taps_per_s = (double)blocks * threads * (double)iters * 36.0 / (ms * 1e-3);
  dfree(sink); cudaDestroyTextureObject(tex); cudaFreeArray(arr);
  return DPE_OK;
}

int dpe_probe_fma_rate(dpe_ctx* ctx, int iters, double* fma_per_s) {
  if (!ctx || iters <= 0 || !fma_per_s) return DPE_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  float* sink; CK(dmalloc(&sink, 4));
  const int blocks = ctx->num_sms * 8, threads = 256;
  const LaunchCfg cfg = cfg_of(ctx);
  launch_probe_fma(iters / 4 + 1, sink, blocks, threads, cfg, 0);
  CK(cudaEventRecord(ctx->ev0, 0));
  launch_probe_fma(iters, sink, blocks, threads, cfg, 0);
  CK(cudaEventRecord(ctx->ev1, 0));
  CK(cudaEventSynchronize(ctx->ev1));
  float ms = 0.f; CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  *fma_per_s = (double)blocks * threads * (double)iters * 128.0 / (ms * 1e-3);
  dfree(sink);
  return DPE_OK;
}

int dpe_probe_tex_weights(dpe_ctx* ctx, int n, float* weights) {
  if (!ctx || n <= 0 || !weights) return DPE_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  const float h[2] = {0.f, 1.f};
  cudaArray_t arr; cudaTextureObject_t tex;
  cudaChannelFormatDesc cd = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
  CK(cudaMallocArray(&arr, &cd, 2, 1));
  CK(cudaMemcpy2DToArray(arr, 0, 0, h, 8, 8, 1, cudaMemcpyHostToDevice));
  cudaResourceDesc rd; memset(&rd, 0, sizeof(rd)); rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
  cudaTextureDesc td; memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModeLinear;
  td.readMode = cudaReadModeElementType;
  CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  float* d; CK(dmalloc(&d, (size_t)(n + 1) * 4));
  launch_probe_weights((unsigned long long)tex, n, d, cfg_of(ctx), 0);
  CK(cudaMemcpy(weights, d, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost));
  dfree(d); cudaDestroyTextureObject(tex); cudaFreeArray(arr);
  return DPE_OK;
}

}  // extern "C"
