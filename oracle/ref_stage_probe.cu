// TEST INFRASTRUCTURE — per-kernel differential oracle.  Compiles the reference's own kernels
// (#include of /root/reference/csrc/DPE-MVS/DPE.cu where it lies; nothing is copied) and runs ONE
// (view, stage) exactly as DPE::RunPatchMatch does (DPE.cu:3126-3249: same kernels, same launch
// geometry, same order) on inputs given as raw arrays, with buffers set up as
// DPE::CudaSpaceInitialization does (DPE.cpp:916-1023), dumping the state after selected steps.
// oracle/make_stage_golden.py drives it on a GPU box and commits inputs + dumps under tests/golden/;
// tests/test_stage_golden.py replays the inputs through the CPU logic simulator, step by step.
// The RNG seed is pinned by oracle/cvshim (clock64() -> DPE_REF_SEED).
//
// input (little endian): int32 W, H, N (images incl. ref), low_w, low_h, state, geom, use_apd,
//   max_iterations, weak_peak_radius, rotate_time; float ransac_threshold, depth_min, depth_max;
//   float images[N][H][W]; float depths[N][H][W] (only if geom); Camera cams[N] (main.h:50-59);
//   float4 planes[H][W] (world normal, depth); uint8 weak[H][W]; uint32 selected[H][W];
//   uint8 edge[H][W]; uint8 edge_low[low_h][low_w]; int32 label[H][W]
// output: for each step s in {0: anchors, 1: init, 2+3i: strong sweeps of iteration i, 3+3i: fit plane,
//   4+3i: weak sweeps, 11: final}: int32 s, then
//   float4 planes[P]; float costs[P]; uint32 selected[P]; uint8 weak[P]; float4 fit[P]; int32 radius[P];
//   short2 neighbours[P][9] (expanded through neighbours_map; -2 for non-WEAK pixels); uint8 reliable[P]
#include "/root/reference/csrc/DPE-MVS/DPE.cu"

#include <cstdio>
#include <vector>

static cudaTextureObject_t make_tex(const float* host, int W, int H) {
  cudaArray* arr;
  cudaChannelFormatDesc cd = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
  cudaMallocArray(&arr, &cd, W, H);
  cudaMemcpy2DToArray(arr, 0, 0, host, W * sizeof(float), W * sizeof(float), H, cudaMemcpyHostToDevice);
  cudaResourceDesc rd;
  memset(&rd, 0, sizeof(rd));
  rd.resType = cudaResourceTypeArray;
  rd.res.array.array = arr;
  cudaTextureDesc td;
  memset(&td, 0, sizeof(td));
  td.addressMode[0] = cudaAddressModeWrap;  // as written in DPE.cpp:929-933
  td.addressMode[1] = cudaAddressModeWrap;
  td.filterMode = cudaFilterModeLinear;
  td.readMode = cudaReadModeElementType;
  td.normalizedCoords = 0;
  cudaTextureObject_t t = 0;
  cudaCreateTextureObject(&t, &rd, &td, NULL);
  return t;
}

template <class T>
static bool rd(FILE* f, std::vector<T>& v, size_t n) {
  v.resize(n);
  return fread(v.data(), sizeof(T), n, f) == n;
}
template <class T>
static T* up(const std::vector<T>& v) {
  T* d = nullptr;
  cudaMalloc(&d, v.size() * sizeof(T));
  cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  return d;
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: ref_stage_probe in.bin out.bin\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  int hi[11];
  float hf[3];
  if (fread(hi, 4, 11, f) != 11 || fread(hf, 4, 3, f) != 3) return 2;
  const int W = hi[0], H = hi[1], N = hi[2], low_w = hi[3], low_h = hi[4];
  const size_t P = (size_t)W * H;
  PatchMatchParams params;  // defaults of main.h:78-106
  params.state = (RunState)hi[5];
  params.geom_consistency = hi[6] != 0;
  params.use_APD = hi[7] != 0;
  params.use_edge = params.use_APD;  // main.cpp:518,522
  params.max_iterations = hi[8];
  params.weak_peak_radius = hi[9];
  params.rotate_time = hi[10];
  params.ransac_threshold = hf[0];
  params.depth_min = hf[1];
  params.depth_max = hf[2];
  params.num_images = N;
  std::vector<float> img, dep;
  std::vector<Camera> cams;
  std::vector<float4> planes;
  std::vector<uchar> weak, edge, edge_low;
  std::vector<unsigned int> selected;
  std::vector<int> label;
  bool ok = rd(f, img, (size_t)N * P);
  if (params.geom_consistency) ok = ok && rd(f, dep, (size_t)N * P);
  ok = ok && rd(f, cams, N) && rd(f, planes, P) && rd(f, weak, P) && rd(f, selected, P) && rd(f, edge, P) &&
       rd(f, edge_low, (size_t)low_w * low_h) && rd(f, label, P);
  fclose(f);
  if (!ok) { fprintf(stderr, "short input\n"); return 2; }

  // neighbours_map / weak_count (DPE.cpp:859-881)
  std::vector<int> nmap(P, 0);
  int weak_count = 0;
  if (params.use_APD) {
    for (size_t i = 0; i < P; ++i)
      if (weak[i] == WEAK) nmap[i] = weak_count++;
  } else {
    for (size_t i = 0; i < P; ++i) weak[i] = STRONG;
  }

  cudaTextureObjects tex_h, dep_h;
  memset(&tex_h, 0, sizeof(tex_h));
  memset(&dep_h, 0, sizeof(dep_h));
  for (int i = 0; i < N; ++i) {
    tex_h.images[i] = make_tex(&img[(size_t)i * P], W, H);
    if (params.geom_consistency) dep_h.images[i] = make_tex(&dep[(size_t)i * P], W, H);
  }
  DataPassHelper h;
  memset(&h, 0, sizeof(h));
  h.width = W; h.height = H; h.low_width = low_w; h.low_height = low_h;
  cudaMalloc(&h.texture_objects_cuda, sizeof(tex_h)); cudaMemcpy(h.texture_objects_cuda, &tex_h, sizeof(tex_h), cudaMemcpyHostToDevice);
  cudaMalloc(&h.texture_depths_cuda, sizeof(dep_h)); cudaMemcpy(h.texture_depths_cuda, &dep_h, sizeof(dep_h), cudaMemcpyHostToDevice);
  h.cameras_cuda = up(cams);
  cudaMalloc(&h.costs_cuda, 4 * P); cudaMemset(h.costs_cuda, 0, 4 * P);
  cudaMalloc(&h.rand_states_cuda, sizeof(curandState) * P);
  h.selected_views_cuda = up(selected);
  cudaMalloc(&h.view_weight_cuda, P * MAX_IMAGES); cudaMemset(h.view_weight_cuda, 0, P * MAX_IMAGES);
  h.plane_hypotheses_cuda = up(planes);
  cudaMalloc(&h.fit_plane_hypotheses_cuda, 16 * P); cudaMemset(h.fit_plane_hypotheses_cuda, 0, 16 * P);
  h.edge_cuda = up(edge);
  h.edge_low_res_cuda = up(edge_low);
  cudaMalloc(&h.edge_neigh_cuda, P * 8 * sizeof(short2)); cudaMemset(h.edge_neigh_cuda, 0xFF, P * 8 * sizeof(short2));
  cudaMalloc(&h.complex_cuda, 4 * P); cudaMemset(h.complex_cuda, 0, 4 * P);
  h.label_cuda = up(label);
  cudaMalloc(&h.label_boundary_cuda, (size_t)(weak_count + 1) * 8 * sizeof(short2));
  cudaMemset(h.label_boundary_cuda, 0xFF, (size_t)(weak_count + 1) * 8 * sizeof(short2));
  cudaMalloc(&h.radius_cuda, 4 * P); cudaMemset(h.radius_cuda, 0, 4 * P);
  h.weak_info_cuda = up(weak);
  cudaMalloc(&h.weak_reliable_cuda, P); cudaMemset(h.weak_reliable_cuda, 0, P);
  cudaMalloc(&h.weak_nearest_strong, P * sizeof(short2)); cudaMemset(h.weak_nearest_strong, 0xFF, P * sizeof(short2));
  h.neighbours_map_cuda = up(nmap);
  cudaMalloc(&h.neighbours_cuda, (size_t)(weak_count + 1) * NEIGHBOUR_NUM * sizeof(short2));
  cudaMemset(h.neighbours_cuda, 0xFF, (size_t)(weak_count + 1) * NEIGHBOUR_NUM * sizeof(short2));
  cudaMalloc(&h.params, sizeof(params)); cudaMemcpy(h.params, &params, sizeof(params), cudaMemcpyHostToDevice);
  DataPassHelper* hd;
  cudaMalloc(&hd, sizeof(h)); cudaMemcpy(hd, &h, sizeof(h), cudaMemcpyHostToDevice);

  FILE* out = fopen(argv[2], "wb");
  if (!out) return 2;
  auto dump = [&](int step) {
    if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "step %d failed: %s\n", step, cudaGetErrorString(cudaGetLastError())); exit(3); }
    std::vector<float4> pl(P), fit(P);
    std::vector<float> co(P);
    std::vector<unsigned int> se(P);
    std::vector<uchar> wk(P), rel(P);
    std::vector<int> ra(P);
    std::vector<short2> nb((size_t)(weak_count + 1) * NEIGHBOUR_NUM), nbx(P * NEIGHBOUR_NUM);
    cudaMemcpy(pl.data(), h.plane_hypotheses_cuda, 16 * P, cudaMemcpyDeviceToHost);
    cudaMemcpy(co.data(), h.costs_cuda, 4 * P, cudaMemcpyDeviceToHost);
    cudaMemcpy(se.data(), h.selected_views_cuda, 4 * P, cudaMemcpyDeviceToHost);
    cudaMemcpy(wk.data(), h.weak_info_cuda, P, cudaMemcpyDeviceToHost);
    cudaMemcpy(fit.data(), h.fit_plane_hypotheses_cuda, 16 * P, cudaMemcpyDeviceToHost);
    cudaMemcpy(ra.data(), h.radius_cuda, 4 * P, cudaMemcpyDeviceToHost);
    cudaMemcpy(rel.data(), h.weak_reliable_cuda, P, cudaMemcpyDeviceToHost);
    cudaMemcpy(nb.data(), h.neighbours_cuda, nb.size() * sizeof(short2), cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < P; ++i)
      for (int k = 0; k < NEIGHBOUR_NUM; ++k)
        nbx[i * NEIGHBOUR_NUM + k] = (params.use_APD && weak[i] == WEAK) ? nb[(size_t)nmap[i] * NEIGHBOUR_NUM + k] : make_short2(-2, -2);
    fwrite(&step, 4, 1, out);
    fwrite(pl.data(), 16, P, out); fwrite(co.data(), 4, P, out); fwrite(se.data(), 4, P, out); fwrite(wk.data(), 1, P, out);
    fwrite(fit.data(), 16, P, out); fwrite(ra.data(), 4, P, out); fwrite(nbx.data(), sizeof(short2), nbx.size(), out);
    fwrite(rel.data(), 1, P, out);
  };

  // ---- DPE::RunPatchMatch, DPE.cu:3126-3249
  dim3 gf((W + 15) / 16, (H + 15) / 16, 1), bf(16, 16, 1);
  dim3 gh((W + 31) / 32, ((H / 2) + 15) / 16, 1), bh(32, 16, 1);
  InitRandomStates<<<gf, bf>>>(hd);
  GenEdgeInform<<<gf, bf>>>(hd);
  FindNearestStrongPoint<<<gf, bf>>>(hd);
  GenNeighbours<<<gf, bf>>>(hd);
  NeigbourUpdate<<<gf, bf>>>(hd);
  dump(0);
  if (argc > 3) {
    // optional: the outputs of GenEdgeInform / FindNearestStrongPoint themselves (DPE.cu:2483-2591, 2855-2889), per pixel:
    // edge_neigh (8 x short2), complex (float), label_boundary (8 x short2, (-2,-2) where the pixel is not WEAK),
    // weak_nearest_strong (short2)
    std::vector<short2> en(P * 8), lbc((size_t)(weak_count + 1) * 8), lbx(P * 8), ns(P);
    std::vector<float> cx(P);
    cudaMemcpy(en.data(), h.edge_neigh_cuda, en.size() * sizeof(short2), cudaMemcpyDeviceToHost);
    cudaMemcpy(cx.data(), h.complex_cuda, 4 * P, cudaMemcpyDeviceToHost);
    cudaMemcpy(lbc.data(), h.label_boundary_cuda, lbc.size() * sizeof(short2), cudaMemcpyDeviceToHost);
    cudaMemcpy(ns.data(), h.weak_nearest_strong, P * sizeof(short2), cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < P; ++i)
      for (int k = 0; k < 8; ++k) lbx[i * 8 + k] = (params.use_APD && weak[i] == WEAK) ? lbc[(size_t)nmap[i] * 8 + k] : make_short2(-2, -2);
    FILE* ex = fopen(argv[3], "wb");
    if (!ex) return 2;
    fwrite(en.data(), sizeof(short2), en.size(), ex); fwrite(cx.data(), 4, P, ex);
    fwrite(lbx.data(), sizeof(short2), lbx.size(), ex); fwrite(ns.data(), sizeof(short2), P, ex);
    fclose(ex);
  }
  RandomInitialization<<<gf, bf>>>(hd);
  dump(1);
  for (int i = 0; i < params.max_iterations; ++i) {
    BlackPixelUpdateStrong<<<gh, bh>>>(i, hd);
    cudaDeviceSynchronize();
    RedPixelUpdateStrong<<<gh, bh>>>(i, hd);
    dump(2 + 3 * i);
    RANSACToGetFitPlane<<<gf, bf>>>(hd);
    dump(3 + 3 * i);
    BlackPixelUpdateWeak<<<gh, bh>>>(i, hd);
    cudaDeviceSynchronize();
    RedPixelUpdateWeak<<<gh, bh>>>(i, hd);
    dump(4 + 3 * i);
  }
  GetDepthandNormal<<<gf, bf>>>(hd);
  cudaDeviceSynchronize();
  BlackPixelFilterStrong<<<gh, bh>>>(hd);
  cudaDeviceSynchronize();
  RedPixelFilterStrong<<<gh, bh>>>(hd);
  cudaDeviceSynchronize();
  DepthToWeak<<<gf, bf>>>(hd);
  cudaDeviceSynchronize();
  LocalRefine<<<gf, bf>>>(hd);
  dump(11);
  fclose(out);
  return 0;
}
