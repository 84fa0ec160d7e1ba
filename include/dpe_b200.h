/*
 * dpe_b200.h — C ABI of the B200-native DPE-MVS PatchMatch path.
 *
 * The reference (shunkenney/DPE-MVS) exports no C symbols: its only exported
 * symbol is PyInit__dpe (csrc/bindings.cpp:31) and the seam between host code
 * and CUDA is the C++ class `DPE` (csrc/DPE-MVS/DPE.h:88-107).  This header is
 * the thin C layer that replaces that seam: plain pointers and sizes, no C++
 * or torch types.  Every entry point cites the reference interface it stands
 * in for.  All functions return 0 on success or a negative dpe_status; none
 * of them ever calls exit() (the reference does: DPE.cpp:633-641).
 */
#ifndef DPE_B200_H_
#define DPE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DPE_API __attribute__((visibility("default")))
#else
#define DPE_API
#endif

#define DPE_MAX_IMAGES 32 /* main.h:39 MAX_IMAGES (ref + <=31 sources)      */
#define DPE_MAX_SRC 31
#define DPE_NEIGHBOUR_NUM 9 /* main.h:40                                     */

/* main.h:66-70 RunState */
enum { DPE_FIRST_INIT = 0, DPE_REFINE_INIT = 1, DPE_REFINE_ITER = 2 };
/* main.h:72-76 PixelState (values persisted in weak.bin) */
enum { DPE_WEAK = 0, DPE_STRONG = 1, DPE_UNKNOWN = 2 };

typedef enum dpe_status {
  DPE_OK = 0,
  DPE_ERR_ARG = -1,
  DPE_ERR_CUDA = -2,
  DPE_ERR_STATE = -3,
  DPE_ERR_IO = -4,
  DPE_ERR_TOO_MANY_IMAGES = -5, /* DPE.cpp:762-765 */
  DPE_ERR_NO_DEVICE = -6,
  DPE_ERR_COMM = -7             /* NCCL */
} dpe_status;

/* Per-stage parameters: the fields of PatchMatchParams (main.h:78-106) that
 * RunDPEPipeline (main.cpp:508-567) varies, plus the always-on defaults. */
typedef struct dpe_stage_params {
  int state;            /* DPE_FIRST_INIT / REFINE_INIT / REFINE_ITER        */
  int geom_consistency; /* main.cpp:526,553                                  */
  int use_apd;          /* main.cpp:517,521  (also drives use_edge)          */
  int max_iterations;   /* 3                                                 */
  int top_k;            /* 4   main.h:83                                     */
  int weak_peak_radius; /* 6 | 4,2,2   main.cpp:528,555                     */
  int rotate_time;      /* min(2^i,4)   main.cpp:524,552                     */
  float ransac_threshold; /* 0.01-0.00125 i  main.cpp:523,551                */
  float geom_factor;    /* 0.2 main.h:104                                    */
} dpe_stage_params;

typedef struct dpe_ctx dpe_ctx;

/* --- lifetime (replaces DPE::DPE / ~DPE, DPE.cpp:674-731, and
 *     cudaSetDevice in RunDPEPipeline, main.cpp:478) ----------------------- */
/* One context per GPU at a time: the texture handle of the running scale and the folded cameras of the views in
 * flight live in per-device constant memory, so two contexts must not run stages on the same device concurrently
 * (one after the other is fine).  Device memory comes from the device's stream-ordered pool and stays with the
 * process after dpe_ctx_destroy unless DPE_RELEASE_MEMORY=1. */
DPE_API int dpe_ctx_create(dpe_ctx** out, int gpu_index);
DPE_API void dpe_ctx_destroy(dpe_ctx* ctx);
DPE_API const char* dpe_last_error(const dpe_ctx* ctx);
/* how many CUDA kernels this context has launched so far */
DPE_API long long dpe_kernel_launches(const dpe_ctx* ctx);

/* --- several GPUs (no counterpart in the reference, which processes problems one after the other on one GPU,
 *     main.cpp:509-558, and exchanges depth maps through depths.dmb files, DPE.cpp:826-844) -------------------
 * The path shards by reference view: problems [0, n_problems) are split into n_ranks contiguous, balanced
 * blocks, one context (= one GPU) per rank; images and cameras are replicated, per-view state lives on the
 * owner.  The one exchange is the depth atlas: after every stage each rank's depth maps are all-gathered with
 * NCCL, one in-place collective per view slot, queued on a side stream behind that view's last kernel.  A
 * sharded context WITHOUT a communicator leaves the exchange to its driver (dpe_stage_atlas + dpe_view_slot
 * between dpe_run_stage and dpe_stage_commit): what the single-GPU sharding test does. */
/* block of rank `rank`: first = rank*q + min(rank, r), count = q + (rank < r), q = n / n_ranks, r = n % n_ranks */
DPE_API int dpe_shard_range(int n_problems, int n_ranks, int rank, int* first, int* count);
/* one process per GPU (bench.py under torchrun): rank 0 makes the id, the launcher hands it to every rank */
#define DPE_COMM_ID_BYTES 128
DPE_API int dpe_comm_get_unique_id(void* id, size_t bytes);
DPE_API int dpe_comm_init_rank(dpe_ctx* ctx, const void* id, int n_ranks, int rank);
/* one process, n contexts on n GPUs (dpe_run_pipeline): context i becomes rank i.  The communicators are kept
 * for the life of the process and reused by later calls with the same GPUs; dpe_comm_reset_all aborts them. */
DPE_API int dpe_comm_init_all(dpe_ctx** ctxs, int n);
DPE_API void dpe_comm_reset_all(void);

/* --- scene upload (replaces DPE::InuputInitialization DPE.cpp:733-914,
 *     SupportInitialization :1025-1052, CudaSpaceInitialization :916-1023) -- */
/* n_scales = ComputeRoundNum (main.cpp:390-408); scale index k has size
 * round(W/2^(n_scales-1-k)) x round(H/2^(n_scales-1-k)); k=0 is the coarsest. */
DPE_API int dpe_scene_begin(dpe_ctx* ctx, int n_views, int width, int height, int n_scales);
/* gray: H*W bytes (cv::IMREAD_GRAYSCALE image, DPE.cpp:745), copied asynchronously (keep it alive until
 * dpe_scene_commit; pinned memory makes the copy overlap), or NULL when the image arrives by
 * dpe_scene_broadcast_images; K,R row-major; depth_min/max are the cam-file values (the 0.6/1.2 factors of
 * DPE.cpp:788-789 are applied inside). */
DPE_API int dpe_scene_set_view(dpe_ctx* ctx, int view, const uint8_t* gray, const float K[9],
                       const float R[9], const float t[3], float depth_min, float depth_max);
/* collective (every rank calls it, after dpe_scene_set_shard): the images rank `root` uploaded are broadcast
 * to all ranks over NVLink (ncclBroadcast), so a scene's images cross PCIe once */
DPE_API int dpe_scene_broadcast_images(dpe_ctx* ctx, int root);
/* src_ids index views of this scene (positions in pair.txt order), n_src<=31 */
DPE_API int dpe_scene_set_pairs(dpe_ctx* ctx, int view, const int* src_ids, int n_src);
/* edge: edges_k (uchar 0/255), label: labels_k (int32) of GetProblemEdges (main.cpp:331-388) at scale index
 * `scale` (0 = coarsest); n_elems = width x height of that scale (dpe_get_size), anything else is rejected.
 * May be called before dpe_scene_commit or after it (then it uploads at once, also while a stage that does not
 * use prep is running); a view this context does not own is ignored. */
DPE_API int dpe_scene_set_prep(dpe_ctx* ctx, int view, int scale, const uint8_t* edge,
                       const int32_t* label, size_t n_elems);
/* this context is rank `rank` of n_ranks and owns the block dpe_shard_range gives it; views >= n_problems are
 * sources only.  Without this call the context owns every view. */
DPE_API int dpe_scene_set_shard(dpe_ctx* ctx, int n_problems, int rank, int n_ranks);
/* test / profiling hook: an unsharded context runs its stages over views [first_view, first_view + count) only
 * (the other views still serve as sources) */
DPE_API int dpe_scene_set_active(dpe_ctx* ctx, int first_view, int count);
/* builds pyramids, textures, per-pair constants; after this the scene is
 * resident in HBM. */
DPE_API int dpe_scene_commit(dpe_ctx* ctx);

/* --- the hot path (replaces DPE::RunPatchMatch DPE.cu:3126-3249 plus the
 *     per-view host tail of ProcessProblem main.cpp:423-446, for every view
 *     this context owns) ---------------------------------------------------- */
/* seed: the reference's curand_init seed (clock64() there, DPE.cu:1032): every (view, stage) starts
 * pixel (x, y) from the cuRAND XORWOW state curand_init(seed, y, x). */
DPE_API int dpe_run_stage(dpe_ctx* ctx, int scale_idx, const dpe_stage_params* params, uint64_t seed);
/* the same in two halves: dpe_stage_begin queues every kernel of every owned view (and, with several ranks, the
 * all-gathers) and returns; dpe_stage_wait_view blocks until `view` has finished the stage, after which its maps
 * can be read (dpe_export_view, dpe_get_maps) while later views still run; dpe_stage_end waits for the rest.
 * dpe_run_stage = begin + end. */
DPE_API int dpe_stage_begin(dpe_ctx* ctx, int scale_idx, const dpe_stage_params* params, uint64_t seed);
DPE_API int dpe_stage_wait_view(dpe_ctx* ctx, int view);
DPE_API int dpe_stage_end(dpe_ctx* ctx);
/* device pointer of the depth atlas the last stage wrote, bytes per slot (one depth map) and in total.  A view's
 * map sits in slot dpe_view_slot: local index * n_ranks + owner rank for problems, behind them for source-only
 * views (which stay zero). */
DPE_API int dpe_stage_atlas(dpe_ctx* ctx, void** dev_ptr, size_t* slot_bytes, size_t* total_bytes);
DPE_API int dpe_view_slot(dpe_ctx* ctx, int view, int* slot);
/* publishes the atlas written by the last stage as the source depth maps of
 * the next geometric-consistency stage (the reference does this through
 * depths.dmb files, DPE.cpp:826-844). */
DPE_API int dpe_stage_commit(dpe_ctx* ctx);

/* View order inside a stage.  0 (default): every view reads the depth maps the PREVIOUS stage
 * committed (Jacobi) — independent of scheduling, required when views are sharded.  1: views run one
 * after the other and each publishes its depth map immediately, so view k reads this stage's maps of
 * views < k and the previous stage's of views > k — the order the reference gets from processing
 * problems serially through depths.dmb files (main.cpp:509-558, DPE.cpp:826-844).  Single GPU only. */
DPE_API int dpe_set_view_order(dpe_ctx* ctx, int sequential);

/* fp32 arithmetic of the matching cost.
 * DPE_COST_REFERENCE_EXACT (default): every cost is computed operation by operation as the reference's
 * --use_fast_math build computes it — homography from R, t, K per evaluation (ComputeHomography, DPE.cu:453-513),
 * tap coordinates associated as its SASS does (ComputeCorrespondingPoint, DPE.cu:515-522), bilateral weights from
 * run-time sigmas (DPE.cu:550-555), raw-intensity moments in its summation order (DPE.cu:716-775), geometric
 * consistency through world coordinates (DPE.cu:881-953): the NCC costs are bit-identical to the reference's
 * device outputs on the golden vectors (tests/golden/ref_probe_c1.npz), the geometric costs on 96 % of them.
 * DPE_COST_REFERENCE: the same moments, but the homography constant-folded per view pair (A - b m^T), taps
 * stepped incrementally, one reciprocal per tap: ~8 % faster over a whole scene, 82 % of the costs bit-identical,
 * the rest within ~2e-5 (a tap crossing a 1/256 filter-weight bin).
 * DPE_COST_CENTRED: intensities centred on the centre pixel before the moment sums, costs within 1e-4 of the
 * float64 formula also on low-contrast patches (where the reference's E[x^2]-E[x]^2 loses ~3 digits). */
enum { DPE_COST_CENTRED = 0, DPE_COST_REFERENCE = 1, DPE_COST_REFERENCE_EXACT = 2 };
DPE_API int dpe_set_cost_arithmetic(dpe_ctx* ctx, int mode);
/* Edge-mode propagation, direction 4: 0 (default) samples the other colour like directions 5-7, which
 * makes a sweep race-free and bit-reproducible; 1 samples the reference's positions (own colour, read while
 * the same launch writes it: DPE.cu:1274-1278) for parity runs; 2 samples the reference's positions
 * from a copy of the maps taken before each half-sweep (what a reference thread reads while the pixels up-left
 * of it are still in flight, the usual case): the reference's positions without its nondeterminism. */
DPE_API int dpe_set_reference_race(dpe_ctx* ctx, int on);
/* test hook: run earlier forms of some kernels, for A/B comparisons of results and speed.  Bit 4: the WEAK-only
 * steps (label-boundary walk, nearest strong pixel, anchor search, plane fit) as image-sized launches, one thread
 * per pixel, instead of over the compacted WEAK lists. */
DPE_API int dpe_debug_set_variants(dpe_ctx* ctx, int mask);
/* test hook: scratch arrays of the view that ran last on the first stream.  what: 0 anchors (P x 9 short2),
 * 1 fit planes (P float4), 2 radius (P int32), 3 costs (P float), 4 weak_reliable (P u8), 5 nearest strong
 * (P short2), 6 complexity (P float), 7 plane hypotheses in reference-camera coordinates (P float4), 8 selected
 * views (P uint32), 9 pixel state (P u8), 10 the image of the shard's first view at that scale (P float), 11 after a stage truncated at a strong sweep (step 2, 5, 8):
 * which candidate each pixel took in it (P u8: 0 kept, 1..8 propagation slot + 1, 10..14 refinement hypothesis), 12 nearest edge
 * pixel per direction (P x 8 short2), 13 label boundary per direction (P x 8 short2; defined for WEAK pixels with a label > 0). */
DPE_API int dpe_debug_read(dpe_ctx* ctx, int what, void* out, size_t bytes);
/* test hook: the following stages stop after `step` of every view-stage, in the numbering of
 * DPE::RunPatchMatch's launch sequence (DPE.cu:3126-3249) used by oracle/ref_stage_probe.cu: 0 anchor search,
 * 1 initialisation, 2+3i strong sweeps of iteration i, 3+3i plane fit, 4+3i weak sweeps, 11 the tail;
 * -1 (default) runs everything.  The scratch arrays then hold the state after that step (dpe_debug_read);
 * the carried maps of the view are not valid until a full stage has run.  Calling it also discards a pending
 * (truncated) stage, so that the same stage can be run again from the same inputs without dpe_stage_commit. */
DPE_API int dpe_debug_stop_after(dpe_ctx* ctx, int step);
/* test hook: replaces the carried maps of `view` at scale k (planes4: P x (world normal, depth); state; selected
 * views) and / or its slot of the committed depth atlas, i.e. what the reference reads back from depths.dmb,
 * normals.dmb, weak.bin and selected_views.bin at the start of a view-stage (DPE.cpp:826-914).  Lets a stage
 * be replayed from stored inputs (tests/golden/ref_stage_weak.npz). */
DPE_API int dpe_debug_set_maps(dpe_ctx* ctx, int view, int scale_idx, const float* planes4, const uint8_t* state,
                               const uint32_t* selected, const float* atlas_depth);

/* --- gate-1 hook: bilateral NCC of fixed plane hypotheses ------------------
 * planes: n_pix x (nx,ny,nz,d) in reference-camera coordinates (n.X + d = 0,
 * DPE.cu:337-342); xy: n_pix x (x,y); cost_out: n_pix x n_src floats =
 * ComputeBilateralNCCOld (DPE.cu:692-778) per source view.  mode 0 = hardware
 * bilinear (the product path), 1 = exact fp32 bilinear from 4 point taps. */
DPE_API int dpe_cost_eval(dpe_ctx* ctx, int view, int scale_idx, int n_pix, const int* xy,
                  const float* planes, int mode, float* cost_out);
/* ComputeGeomConsistencyCost (DPE.cu:915-953) against the committed atlas */
DPE_API int dpe_geom_eval(dpe_ctx* ctx, int view, int scale_idx, int n_pix, const int* xy,
                  const float* planes, float* cost_out);

/* --- results (replaces DPE::GetPlaneHypothesis/GetPixelStates/
 *     GetSelectedViews DPE.cpp:1091-1105) -----------------------------------
 * Maps of the last stage run for `view` at that stage's scale. Any pointer may
 * be NULL.  depth: H*W (0 where out of range, main.cpp:431-434); normal: H*W*3
 * world-space; state: H*W PixelState; selected: H*W bitmasks. */
DPE_API int dpe_get_size(dpe_ctx* ctx, int scale_idx, int* width, int* height);
DPE_API int dpe_get_maps(dpe_ctx* ctx, int view, float* depth, float* normal3, uint8_t* state,
                 uint32_t* selected);
/* the payloads of depth.npy / normal.npy / weak.npy of `view` (Write*AsNpy, main.cpp:99-260): depth zeroed where
 * the pixel is UNKNOWN (main.cpp:36-46), normal H*W*3, weak int8 {0 unknown, 1 weak, 2 strong} (main.cpp:190-197);
 * packed on the device, copied on a stream of their own (pinned destinations overlap running stages).  Any pointer
 * may be NULL. */
DPE_API int dpe_export_view(dpe_ctx* ctx, int view, float* depth, float* normal3, int8_t* weak);
/* viz=True (ShowDepthMap / ShowNormalMap / ShowWeakImage, DPE.cpp:384-503, called per view-stage at main.cpp:448-454): the
 * three colour-mapped images of `view`'s current maps, interleaved B,G,R, rendered on the device.  The pointers are
 * DEVICE pointers into a staging buffer of the context, valid until the next call; the size is that of the view's scale. */
DPE_API int dpe_viz_render(dpe_ctx* ctx, int view, void** bgr_depth, void** bgr_normal, void** bgr_weak, int* width, int* height);
/* number of (pixel,hypothesis,view) bilateral-NCC units evaluated so far
 * (NCCOld = 1 unit = 36 taps, NCCNew = taps/36); 0 unless counting is on. */
DPE_API int dpe_set_count_evals(dpe_ctx* ctx, int on);
DPE_API double dpe_eval_units(dpe_ctx* ctx);
/* GPU milliseconds spent inside dpe_run_stage so far (CUDA events). */
DPE_API double dpe_stage_gpu_ms(dpe_ctx* ctx);
/* the part of it this rank spent after its own last view, waiting for the all-gathers to finish (the exposed
 * exchange plus the wait for slower ranks); 0 on one GPU */
DPE_API double dpe_stage_comm_ms(dpe_ctx* ctx);

/* per-kernel-class profile: dpe_set_profile(ctx, n) with n > 0 makes dpe_run_stage bracket every
 * launch of its first n local views with CUDA events (those views run on one stream, the eval-unit
 * counter is read after each launch); the remaining views run as usual so that the stage still sees
 * every view's depth map.  n = 0 switches it off.  Classes index the
 * arrays returned by dpe_get_profile (each DPE_N_KERNEL_CLASSES long). */
enum {
  DPE_K_LOAD = 0, DPE_K_EDGE_INFO, DPE_K_NEAREST, DPE_K_NEIGHBOURS, DPE_K_INIT, DPE_K_STRONG, DPE_K_FIT,
  DPE_K_WEAK, DPE_K_EXTRACT, DPE_K_MEDIAN, DPE_K_CLASSIFY, DPE_K_FINISH, DPE_N_KERNEL_CLASSES
};
DPE_API int dpe_set_profile(dpe_ctx* ctx, int on);
DPE_API int dpe_get_profile(dpe_ctx* ctx, double* ms, double* units, long long* launches);

/* tap-loop study: the candidate-scoring loop of the strong sweep in isolation on `view`'s current
 * maps (n_cand neighbour planes x all source views per pixel of one colour); variant selects how many
 * tap rows are fetched before use and the register budget (dpe_kernels.cu: launch_ncc_bench). */
DPE_API int dpe_bench_ncc(dpe_ctx* ctx, int view, int variant, int n_cand, int reps, double* units_per_s,
                  double* checksum);

/* --- depth-map fusion on the device (replaces DPE::RunFusion, DPE.cpp:1220-1370, and the point list
 *     ExportPointCloud writes, DPE.cpp:532-572) -----------------------------------------------------
 * Fusion reads the final (world normal, depth) map, pixel states and colour image (B,G,R interleaved) of every
 * problem view; cameras and source lists are the scene's.  Two ways to give it the maps:
 *   dpe_fuse_prepare   the maps the last stage left on the GPU(s) — nothing goes through the host; with several
 *                      ranks it is a collective that all-gathers the ranks' blocks with NCCL, after which every
 *                      rank holds the maps of all problem views: one rank can fuse all of them in order (the
 *                      reference's semantics), or every rank its own block against its own marks (faster; marks
 *                      do not cross ranks, so surface seen from two blocks is fused twice);
 *   dpe_fuse_set_view  maps from the host (what the reference reads from depths.dmb / normals.dmb / weak.bin:
 *                      depth zeroed where out of range), unsharded contexts only.
 * Colours: dpe_fuse_set_view, or dpe_fuse_set_color (asynchronous; rank `root` alone may upload and
 * dpe_fuse_broadcast_colors, a collective, hands them to the other ranks).  dpe_fuse_run fuses views
 * [first_view, first_view + count) in index order, starting from a clean set of marks; the cloud is the same from
 * run to run.  dpe_fuse_get copies it out (n_points x 3 each).
 * dpe_fuse_set_block: the reference's optional <dense>/blocks/mask_<id>.jpg gate (DPE.cpp:1242-1268, 1296-1298) — a
 * grey image of the view's size; a pixel of that view whose mask value is below 128 is not fused as a reference pixel
 * (it still serves as a source pixel of other views).  Set it on the context that fuses the view; views without a mask
 * are not gated.  dpe_run_pipeline reads the folder when it exists. */
DPE_API int dpe_fuse_prepare(dpe_ctx* ctx);
DPE_API int dpe_fuse_set_view(dpe_ctx* ctx, int view, const float* depth, const float* normal3, const uint8_t* state,
                      const uint8_t* bgr);
DPE_API int dpe_fuse_set_color(dpe_ctx* ctx, int view, const uint8_t* bgr);
DPE_API int dpe_fuse_set_block(dpe_ctx* ctx, int view, const uint8_t* mask);
DPE_API int dpe_fuse_broadcast_colors(dpe_ctx* ctx, int root);
DPE_API int dpe_fuse_run(dpe_ctx* ctx, int first_view, int count, size_t* n_points);
DPE_API int dpe_fuse_get(dpe_ctx* ctx, float* xyz, uint8_t* bgr);
/* the same cloud as the vertex records of the PLY file ExportPointCloud writes (DPE.cpp:553-569): n_points x 15 bytes,
 * x y z float32 little-endian + blue green red uint8 */
DPE_API int dpe_fuse_get_ply_records(dpe_ctx* ctx, void* records);

/* --- micro-benchmarks used for the roofline denominators ------------------ */
/* filtered tex2D<float> taps per second on a WxH float texture */
DPE_API int dpe_probe_tex_rate(dpe_ctx* ctx, int width, int height, int iters, double* taps_per_s);
/* same with a texel format (0 f32, 1 f16, 2 u8 normalised), a lane layout (0 row of 32, 1 red/black
 * zig-zag, 2 8x4 block, 3 one colour of 8x8, 4 16x2) and a 2x2 reference->source map m */
DPE_API int dpe_probe_tex_pattern(dpe_ctx* ctx, int fmt, int layout, int width, int height, int iters, const float m[4],
                          int threads, int blocks_per_sm, double* taps_per_s);
DPE_API int dpe_probe_fma_rate(dpe_ctx* ctx, int iters, double* fma_per_s);
/* weights[i] = hardware bilinear result at fractional offset i/n between a
 * texel holding 0 and a texel holding 1 (characterises the 1.8 fixed-point
 * interpolation weights). */
DPE_API int dpe_probe_tex_weights(dpe_ctx* ctx, int n, float* weights);

/* --- whole pipeline (replaces RunDPEPipeline main.cpp:474-600); same
 *     argument meaning as the Python / CLI surface --------------------------- */
DPE_API int dpe_run_pipeline(const char* dense_folder, int gpu_index, int verbose, int fusion, int viz,
                     int depth, int normal, int weak, int edge);

#ifdef __cplusplus
}
#endif
#endif /* DPE_B200_H_ */
