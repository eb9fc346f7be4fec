// Dev microbenchmark: issue/throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) and a 50/50 mix with
// FMNMX on sm_100a.  Build+run under gpurun:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ub tools/ubench_fp32x2.cu && /tmp/ub
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) bench(float *out, int iters, float seed) {
  float a[8], b = seed, c = 0.5f;
  float2 p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = make_float2(a[i], a[i] + 1.f); }
  const float2 b2 = make_float2(b, b * 1.01f), c2 = make_float2(c, c);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { a[i] = fmaf(a[i], b, c); }
      else if (MODE == 1) { p[i] = __ffma2_rn(p[i], b2, c2); }
      else if (MODE == 2) { a[i] = fmaf(a[i], b, c); a[i] = fminf(a[i], 1e30f); }   // FFMA + FMNMX
      else if (MODE == 3) { p[i] = __ffma2_rn(p[i], b2, c2); p[i].x = fminf(p[i].x, 1e30f); p[i].y = fminf(p[i].y, 1e30f); }
      else if (MODE == 4) { p[i] = __fmul2_rn(p[i], b2); }
      else if (MODE == 5) { p[i] = __fadd2_rn(p[i], c2); }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + p[i].x + p[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, double lanes_per_op, float *d) {
  const int iters = 4096, blocks = 148 * 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<MODE><<<blocks, 256>>>(d, iters, 1.0001f);
  cudaEventRecord(e0);
  bench<MODE><<<blocks, 256>>>(d, iters, 1.0001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)blocks * 256 * iters * 8;
  printf("%-28s %8.3f ms  %7.1f Gop/s (thread-instr)  %7.1f G lane-FMA/s  -> %.1f lane-ops/clk/SM @1.9GHz\n", name, ms,
         ops / ms * 1e-6, ops * lanes_per_op / ms * 1e-6, ops * lanes_per_op / ms * 1e-6 / 148 / 1.9);
}

int main() {
  float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  run<0>("FFMA scalar", 1, d);
  run<1>("FFMA2 packed", 2, d);
  run<2>("FFMA + FMNMX", 1, d);
  run<3>("FFMA2 + 2 FMNMX", 2, d);
  run<4>("FMUL2", 2, d);
  run<5>("FADD2", 2, d);
  return 0;
}
