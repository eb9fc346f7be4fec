// Dev probe: what copy bandwidth do the candidate store paths reach on one image of the C2 output size?
//   A  TMA box load -> TMA box store, one elected thread per CTA drives an S-stage ring (no thread touches the data)
//   B  TMA box load -> LDS.128 -> STG.128 by 256 threads (what a kernel that transforms the tile would do), 2 stages
//   C  persistent grid-stride LDG.128 -> STG.128, no shared memory
//   M  cudaMemcpy2DAsync device -> device
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/copy_probe tools/copy_probe.cu && gpurun_out/copy_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#define OVRFSR_MODE_NS probe
#include "../openvr_fsr_b200/csrc/tma_utils.cuh"
using namespace ovrfsr;

__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int x, int y, const void *src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(x), "r"(y),
               "r"(smem_u32(src))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// ---- A: pure TMA ring -----------------------------------------------------------------------------------
template <int TW, int TH, int S, int D>
__global__ void __launch_bounds__(32) copyA(const __grid_constant__ CUtensorMap src, const __grid_constant__ CUtensorMap dst, int w, int h) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[S];
  const int tilesX = (w + TW - 1) / TW, tilesY = (h + TH - 1) / TH, n = tilesX * tilesY;
  if (threadIdx.x != 0) return;
  for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
  fence_barrier_init();
  int cnt = 0;
  for (int t = blockIdx.x; t < n; t += gridDim.x) ++cnt;
  for (int i = 0; i < cnt + D; ++i) {
    const int j = i - D;
    if (j >= 0) {
      const int t = blockIdx.x + j * gridDim.x, s = j % S;
      mbar_wait(&full[s], (j / S) & 1);
      tma_store_2d(&dst, (t % tilesX) * TW, (t / tilesX) * TH, smem + s * (TW * TH * 4));
      bulk_commit();
    }
    if (i < cnt) {
      const int t = blockIdx.x + i * gridDim.x, s = i % S;
      if (i >= S) bulk_wait_read<S - D>(); // the store that last read stage s is done with it
      mbar_arrive_expect_tx(&full[s], TW * TH * 4);
      tma_load_2d(smem + s * (TW * TH * 4), &src, (t % tilesX) * TW, (t / tilesX) * TH, &full[s]);
    }
  }
  bulk_wait_all<0>();
}

// ---- B: TMA load, thread stores ---------------------------------------------------------------------------
template <int TW, int TH>
__global__ void __launch_bounds__(256) copyB(const __grid_constant__ CUtensorMap src, uint8_t *dst, size_t pitch, int w, int h) {
  __shared__ __align__(128) uint8_t smem[2][TW * TH * 4];
  __shared__ uint64_t full[2];
  const int tilesX = (w + TW - 1) / TW, tilesY = (h + TH - 1) / TH, n = tilesX * tilesY;
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&full[0], 1); mbar_init(&full[1], 1);
    fence_barrier_init();
    if ((int)blockIdx.x < n) {
      mbar_arrive_expect_tx(&full[0], TW * TH * 4);
      tma_load_2d(smem[0], &src, (blockIdx.x % tilesX) * TW, (blockIdx.x / tilesX) * TH, &full[0]);
    }
  }
  __syncthreads();
  int i = 0;
  for (int t = blockIdx.x; t < n; t += gridDim.x, ++i) {
    const int s = i & 1, tn = t + gridDim.x;
    if (tid == 0 && tn < n) {
      mbar_arrive_expect_tx(&full[s ^ 1], TW * TH * 4);
      tma_load_2d(smem[s ^ 1], &src, (tn % tilesX) * TW, (tn / tilesX) * TH, &full[s ^ 1]);
    }
    mbar_wait(&full[s], (i >> 1) & 1);
    const int ox = (t % tilesX) * TW, oy = (t / tilesX) * TH;
    constexpr int VPR = TW / 4; // 16-byte vectors per tile row
    for (int q = tid; q < VPR * TH; q += 256) {
      const int r = q / VPR, c = q % VPR;
      const uint4 v = *reinterpret_cast<const uint4 *>(smem[s] + (size_t)q * 16);
      if (oy + r < h && ox + c * 4 + 3 < w) *reinterpret_cast<uint4 *>(dst + (size_t)(oy + r) * pitch + (size_t)(ox + c * 4) * 4) = v;
    }
    __syncthreads(); // everyone is done reading stage s before it is refilled (next iteration's prefetch targets s)
  }
}

// ---- C: plain vector copy ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) copyC(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, size_t pitch, int w, int h) {
  const int vpr = w / 4; // 16-byte vectors per row (w % 4 == 0 here)
  const long total = (long)vpr * h;
  for (long q = (long)blockIdx.x * 256 + threadIdx.x; q < total; q += (long)gridDim.x * 256) {
    const int r = (int)(q / vpr), c = (int)(q % vpr);
    *reinterpret_cast<uint4 *>(dst + (size_t)r * pitch + (size_t)c * 16) = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)r * pitch + (size_t)c * 16));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn enc;
static CUtensorMap make_map(void *p, int w, int h, size_t pitch, int bw, int bh) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h}, strides[1] = {(cuuint64_t)pitch};
  cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

constexpr int W = 2244, H = 2492, NB = 8;
static size_t pitch = ((size_t)W * 4 + 255) / 256 * 256;
static uint8_t *srcs[NB], *dsts[NB];

template <typename F>
static void timeit(const char *name, F launch) {
  for (int i = 0; i < NB; ++i) launch(i);
  cudaDeviceSynchronize();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("%-40s ERROR %s\n", name, cudaGetErrorString(e)); exit(2); }
  // verify buffer 0
  std::vector<uint8_t> a(pitch * H), b(pitch * H);
  cudaMemcpy(a.data(), srcs[0], pitch * H, cudaMemcpyDeviceToHost); cudaMemcpy(b.data(), dsts[0], pitch * H, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int y = 0; y < H; ++y) bad += memcmp(a.data() + y * pitch, b.data() + y * pitch, (size_t)W * 4) != 0;
  for (int i = 0; i < NB; ++i) cudaMemset(dsts[i], 0, pitch * H);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = 5;
  cudaEventRecord(e0);
  for (int r = 0; r < reps; ++r) for (int i = 0; i < NB; ++i) launch(i);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double us = ms * 1e3 / (reps * NB), gbs = 2.0 * W * H * 4 / us / 1e3;
  printf("%-40s %7.2f us  %7.0f GB/s  (%ld bad rows)\n", name, us, gbs, bad);
}

int main() {
  cudaFree(0);
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  enc = (EncodeTiledFn)p;
  std::vector<uint8_t> hsrc(pitch * H);
  for (size_t i = 0; i < hsrc.size(); ++i) hsrc[i] = (uint8_t)(i * 2654435761u >> 13);
  for (int i = 0; i < NB; ++i) {
    cudaMalloc(&srcs[i], pitch * H); cudaMalloc(&dsts[i], pitch * H);
    cudaMemcpy(srcs[i], hsrc.data(), pitch * H, cudaMemcpyHostToDevice);
  }
  timeit("M cudaMemcpy2DAsync", [&](int i) { cudaMemcpy2DAsync(dsts[i], pitch, srcs[i], pitch, (size_t)W * 4, H, cudaMemcpyDeviceToDevice, 0); });
#define RUN_A(TW, TH, S, D, CPS)                                                                                        \
  {                                                                                                                     \
    static CUtensorMap ms[NB], md[NB];                                                                                  \
    for (int i = 0; i < NB; ++i) { ms[i] = make_map(srcs[i], W, H, pitch, TW, TH); md[i] = make_map(dsts[i], W, H, pitch, TW, TH); } \
    cudaFuncSetAttribute(copyA<TW, TH, S, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, S * TW * TH * 4);              \
    timeit("A tma->tma " #TW "x" #TH " S" #S " D" #D " cta/sm " #CPS,                                                  \
           [&](int i) { copyA<TW, TH, S, D><<<148 * CPS, 32, S * TW * TH * 4>>>(ms[i], md[i], W, H); });                   \
  }
  RUN_A(64, 32, 4, 2, 1) RUN_A(64, 32, 4, 2, 2) RUN_A(64, 32, 4, 2, 4) RUN_A(64, 32, 8, 4, 2) RUN_A(64, 32, 8, 4, 3)
  RUN_A(128, 32, 4, 2, 2) RUN_A(128, 32, 6, 3, 2) RUN_A(256, 16, 4, 2, 2) RUN_A(64, 64, 4, 2, 2) RUN_A(64, 32, 2, 1, 4) RUN_A(64, 32, 2, 1, 8)
#define RUN_B(TW, TH, CPS)                                                                                              \
  {                                                                                                                     \
    static CUtensorMap ms[NB];                                                                                          \
    for (int i = 0; i < NB; ++i) ms[i] = make_map(srcs[i], W, H, pitch, TW, TH);                                        \
    timeit("B tma->lds->stg128 " #TW "x" #TH " cta/sm " #CPS, [&](int i) { copyB<TW, TH><<<148 * CPS, 256>>>(ms[i], dsts[i], pitch, W, H); }); \
  }
  RUN_B(64, 32, 2) RUN_B(64, 32, 3) RUN_B(64, 32, 4) RUN_B(128, 16, 3)
  timeit("C ldg128->stg128 grid 148x4", [&](int i) { copyC<<<148 * 4, 256>>>(srcs[i], dsts[i], pitch, W, H); });
  timeit("C ldg128->stg128 grid 148x8", [&](int i) { copyC<<<148 * 8, 256>>>(srcs[i], dsts[i], pitch, W, H); });
  return 0;
}
