// FP32 issue-rate microbenchmark for B200: scalar FFMA vs packed FFMA2 (fma.rn.f32x2), and an FFMA + LOP3/IMAD mix.
// Prints the measured FP32 FMA peak (TFLOP/s) that BASELINE.md asks the builder to measure.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/microbench_fp32.cu -o tools/microbench_fp32 && tools/microbench_fp32
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, int iters, float a, float b) {
  float v[16];
  float2 w[8];
  unsigned u[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 0.001f + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) { w[i] = make_float2(v[2 * i], v[2 * i + 1]); u[i] = threadIdx.x + i; }
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], a, b);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = fma2(w[i], make_float2(a, a), make_float2(b, b));
    } else if (MODE == 2) {   // 8 FFMA2 (=16 flop-lanes) + 8 LOP3: does the packed form leave issue slots for integer work?
#pragma unroll
      for (int i = 0; i < 8; ++i) { w[i] = fma2(w[i], make_float2(a, a), make_float2(b, b)); u[i] = (u[i] ^ (u[(i + 1) & 7] >> 3)) + 0x9E3779B9u; }
    } else {                  // 16 FFMA + 8 LOP3/IADD
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], a, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = (u[i] ^ (u[(i + 1) & 7] >> 3)) + 0x9E3779B9u;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += w[i].x + w[i].y + (float)u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(const char* name, float* out, int grid, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<MODE><<<grid, 256>>>(out, iters, 1.0001f, 0.5f);
  cudaEventRecord(e0);
  bench<MODE><<<grid, 256>>>(out, iters, 1.0001f, 0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fma_lanes = (double)grid * 256 * iters * 16;
  printf("%-28s %8.3f ms  %7.2f TFLOP/s (fp32 FMA = 2 flop)\n", name, ms, 2.0 * fma_lanes / ms / 1e9);
  return ms;
}

int main() {
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * 8, iters = 1 << 16;
  float* out; cudaMalloc(&out, (size_t)grid * 256 * sizeof(float));
  run<0>("FFMA  (16 scalar)", out, grid, iters);
  run<1>("FFMA2 (8 packed)", out, grid, iters);
  run<3>("16 FFMA + 8 int ops", out, grid, iters);
  run<2>("8 FFMA2 + 8 int ops", out, grid, iters);
  return 0;
}
