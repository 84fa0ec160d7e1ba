// TEST INFRASTRUCTURE — golden-vector generator.  Compiles the reference's own device code
// (#include of /root/reference/csrc/DPE-MVS/DPE.cu where it lies; nothing is copied) and
// calls ComputeBilateralNCCOld (DPE.cu:692-778) and ComputeGeomConsistencyCost
// (DPE.cu:915-953) for given pixels / plane hypotheses on the GPU, with textures set up
// exactly as DPE::CudaSpaceInitialization does (DPE.cpp:919-961).  oracle/make_golden.py
// drives it and commits the outputs under tests/golden/ to pin oracle/ncc_oracle.py.
//
// input file (little endian): int32 W, H, N(images incl. ref), n_pts;
//   N x W*H float images; N x W*H float depth maps; N x Camera (main.h:50-59, 112 bytes);
//   n_pts x int2 pixels; n_pts x float4 planes
// output file: n_pts x (N-1) float NCC costs, then n_pts x (N-1) float geometric costs
#include "/root/reference/csrc/DPE-MVS/DPE.cu"

#include <cstdio>
#include <vector>

__global__ void ProbeKernel(DataPassHelper* helper, int n_pts, const int2* pts, const float4* planes, float* ncc,
                            float* geom) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pts) return;
  const int n_src = helper->params->num_images - 1;
  for (int s = 1; s <= n_src; ++s) {
    ncc[i * n_src + s - 1] = ComputeBilateralNCCOld(pts[i], s, planes[i], helper);
    geom[i * n_src + s - 1] = ComputeGeomConsistencyCost(pts[i], s, planes[i], helper);
  }
}

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

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: ref_probe in.bin out.bin\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  int hdr[4];
  if (fread(hdr, 4, 4, f) != 4) return 2;
  const int W = hdr[0], H = hdr[1], N = hdr[2], n_pts = hdr[3];
  std::vector<float> img((size_t)N * W * H), dep((size_t)N * W * H);
  std::vector<Camera> cams(N);
  std::vector<int2> pts(n_pts);
  std::vector<float4> planes(n_pts);
  bool ok = fread(img.data(), 4, img.size(), f) == img.size() && fread(dep.data(), 4, dep.size(), f) == dep.size() &&
            fread(cams.data(), sizeof(Camera), N, f) == (size_t)N && fread(pts.data(), sizeof(int2), n_pts, f) == (size_t)n_pts &&
            fread(planes.data(), sizeof(float4), n_pts, f) == (size_t)n_pts;
  fclose(f);
  if (!ok) { fprintf(stderr, "short input\n"); return 2; }

  cudaTextureObjects tex_h, dep_h;
  memset(&tex_h, 0, sizeof(tex_h));
  memset(&dep_h, 0, sizeof(dep_h));
  for (int i = 0; i < N; ++i) {
    tex_h.images[i] = make_tex(&img[(size_t)i * W * H], W, H);
    dep_h.images[i] = make_tex(&dep[(size_t)i * W * H], W, H);
  }
  cudaTextureObjects *tex_d, *dep_d;
  cudaMalloc(&tex_d, sizeof(tex_h)); cudaMemcpy(tex_d, &tex_h, sizeof(tex_h), cudaMemcpyHostToDevice);
  cudaMalloc(&dep_d, sizeof(dep_h)); cudaMemcpy(dep_d, &dep_h, sizeof(dep_h), cudaMemcpyHostToDevice);
  Camera* cams_d;
  cudaMalloc(&cams_d, sizeof(Camera) * N); cudaMemcpy(cams_d, cams.data(), sizeof(Camera) * N, cudaMemcpyHostToDevice);
  PatchMatchParams params;  // defaults of main.h:78-106
  params.num_images = N;
  PatchMatchParams* params_d;
  cudaMalloc(&params_d, sizeof(params)); cudaMemcpy(params_d, &params, sizeof(params), cudaMemcpyHostToDevice);
  DataPassHelper helper;
  memset(&helper, 0, sizeof(helper));
  helper.width = W; helper.height = H;
  helper.texture_objects_cuda = tex_d; helper.texture_depths_cuda = dep_d; helper.cameras_cuda = cams_d;
  helper.params = params_d;
  DataPassHelper* helper_d;
  cudaMalloc(&helper_d, sizeof(helper)); cudaMemcpy(helper_d, &helper, sizeof(helper), cudaMemcpyHostToDevice);
  int2* pts_d; float4* planes_d; float *ncc_d, *geom_d;
  cudaMalloc(&pts_d, sizeof(int2) * n_pts); cudaMemcpy(pts_d, pts.data(), sizeof(int2) * n_pts, cudaMemcpyHostToDevice);
  cudaMalloc(&planes_d, sizeof(float4) * n_pts); cudaMemcpy(planes_d, planes.data(), sizeof(float4) * n_pts, cudaMemcpyHostToDevice);
  const size_t n_out = (size_t)n_pts * (N - 1);
  cudaMalloc(&ncc_d, 4 * n_out); cudaMalloc(&geom_d, 4 * n_out);
  ProbeKernel<<<(n_pts + 63) / 64, 64>>>(helper_d, n_pts, pts_d, planes_d, ncc_d, geom_d);
  if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 3; }
  std::vector<float> ncc(n_out), geom(n_out);
  cudaMemcpy(ncc.data(), ncc_d, 4 * n_out, cudaMemcpyDeviceToHost);
  cudaMemcpy(geom.data(), geom_d, 4 * n_out, cudaMemcpyDeviceToHost);
  f = fopen(argv[2], "wb");
  fwrite(ncc.data(), 4, n_out, f);
  fwrite(geom.data(), 4, n_out, f);
  fclose(f);
  return 0;
}
