// Dev probe: does a 2-D TMA tile load work when the box exceeds the tensor / starts at negative coords?
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tma_probe tools/tma_probe.cu -lcuda && /tmp/tma_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#define OVRFSR_MODE_NS probe
#include "../openvr_fsr_b200/csrc/tma_utils.cuh"
using namespace ovrfsr;

template <int BW, int BH>
__global__ void k(const __grid_constant__ CUtensorMap map, uint32_t *out, int x0, int y0) {
  __shared__ __align__(128) uint32_t tile[BW * BH];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(&bar, BW * BH * 4);
    tma_load_2d(tile, &map, x0, y0, &bar);
  }
  __syncthreads();
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = tile[i];
}

template <int BW, int BH>
int run(int W, int H, int pitchElems, int x0, int y0) {
  std::vector<uint32_t> h((size_t)pitchElems * H);
  for (int y = 0; y < H; ++y) for (int x = 0; x < pitchElems; ++x) h[(size_t)y * pitchElems + x] = (y << 16) | x | 0x80000000u;
  uint32_t *d, *o;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, BW * BH * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}; cuuint64_t strides[1] = {(cuuint64_t)pitchElems * 4};
  cuuint32_t box[2] = {BW, BH}; cuuint32_t es[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("W=%d H=%d box %dx%d: encode failed %d\n", W, H, BW, BH, (int)r); return 1; }
  k<BW, BH><<<1, 128>>>(map, o, x0, y0);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("W=%d H=%d box %dx%d at (%d,%d): kernel error %s\n", W, H, BW, BH, x0, y0, cudaGetErrorString(e)); return 2; }
  std::vector<uint32_t> t(BW * BH);
  cudaMemcpy(t.data(), o, t.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int y = 0; y < BH; ++y) for (int x = 0; x < BW; ++x) {
    int gx = x0 + x, gy = y0 + y;
    uint32_t want = (gx < 0 || gy < 0 || gx >= W || gy >= H) ? 0u : (((uint32_t)gy << 16) | gx | 0x80000000u);
    bad += t[y * BW + x] != want;
  }
  printf("W=%d H=%d pitch=%d box %dx%d at (%d,%d): ok, %d mismatches\n", W, H, pitchElems * 4, BW, BH, x0, y0, bad);
  return 0;
}
int main(int argc, char **argv) {
  cuInit(0); cudaFree(0);
  int v = argc > 1 ? atoi(argv[1]) : 0;  // one variant per process: an illegal instruction is sticky
  switch (v) {
    case 0: return run<72, 34>(256, 256, 256, -4, -1);
    case 1: return run<72, 34>(256, 256, 256, 60, 31);
    case 2: return run<60, 29>(1683, 1869, 1684, -4, -2);
    case 3: return run<60, 29>(1683, 1869, 1684, 1640, 1850);
    case 4: return run<44, 21>(16, 16, 16, -4, -2);
    case 5: return run<72, 34>(2244, 2492, 2244, 2236, 2463);
  }
  return 0;
}
