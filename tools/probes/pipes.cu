// Issue / pipe rates of the instructions the MALA kernel is made of, on one SM sub-partition:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/probes/pipes.cu -o tools/probes/pipes && tools/probes/pipes
// Each mode runs 8 independent dependency chains per thread; printed = SM cycles per warp-instruction per scheduler
// (1.0 = one instruction every cycle) at 1, 2, 4 and 8 warps per scheduler.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long mulwide(uint32_t a, uint32_t b) {
  unsigned long long d;
  asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ float ffma(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float mufu_lg2(float a) {
  float d;
  asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a));
  return d;
}
__device__ __forceinline__ float mufu_sin(float a) {
  float d;
  asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a));
  return d;
}
__device__ __forceinline__ float mufu_tanh(float a) {
  float d;
  asm volatile("tanh.approx.f32 %0, %1;" : "=f"(d) : "f"(a));
  return d;
}
__device__ __forceinline__ uint32_t mufu_tanh_h2(uint32_t a) {
  uint32_t d;
  asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ uint32_t mufu_tanh_b2(uint32_t a) {
  uint32_t d;
  asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ float fsel(float a, float b, int p) {
  float d;
  asm volatile("{.reg .pred q; setp.ne.s32 q, %3, 0; selp.f32 %0, %1, %2, q;}" : "=f"(d) : "f"(a), "f"(b), "r"(p));
  return d;
}

constexpr int ILP = 8;
// MODE: 0 FFMA, 1 FFMA2, 2 IMAD.WIDE, 3 LOP3, 4 IMAD.WIDE+LOP3 (Philox round shape), 5 FFMA2+LOP3, 6 FFMA2+IMAD.WIDE,
//       7 MUFU.LG2, 8 FFMA+IMAD.WIDE, 9 MUFU.SIN + 4 FFMA2, 10 selp, 11 FFMA + FFMA2, 12 LDS.128+STS.128 + 2 FFMA2
template <int MODE>
__global__ void __launch_bounds__(1024) bench(float* out, int iters, uint32_t ka, float fa, long long* cyc) {
  extern __shared__ float4 sm[];
  float4* vsm = sm + (iters & 1);
  uint32_t u[ILP], w[ILP];
  float f[ILP];
  unsigned long long p[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    u[i] = threadIdx.x * 7 + i; w[i] = threadIdx.x * 13 + i; f[i] = 1.f + 0.001f * (threadIdx.x + i);
    p[i] = ((unsigned long long)__float_as_uint(f[i]) << 32) | __float_as_uint(f[i]);
  }
  const unsigned long long pa = ((unsigned long long)__float_as_uint(fa) << 32) | __float_as_uint(fa);
  sm[threadIdx.x] = make_float4(0, 0, 0, 0);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) f[i] = ffma(f[i], fa, fa);
      if (MODE == 1) p[i] = fma2(p[i], pa, pa);
      if (MODE == 2) { const unsigned long long m = mulwide(u[i], ka); u[i] = (uint32_t)(m >> 32) + (uint32_t)m; }
      if (MODE == 3) u[i] = xor3(u[i], w[i], ka);
      if (MODE == 4) { const unsigned long long m = mulwide(u[i], ka); u[i] = xor3((uint32_t)(m >> 32), w[i], ka); w[i] = (uint32_t)m; }
      if (MODE == 5) { p[i] = fma2(p[i], pa, pa); u[i] = xor3(u[i], w[i], ka); }
      if (MODE == 6) { p[i] = fma2(p[i], pa, pa); const unsigned long long m = mulwide(u[i], ka); u[i] = (uint32_t)(m >> 32); w[i] ^= (uint32_t)m; }
      if (MODE == 7) f[i] = mufu_lg2(f[i]);
      if (MODE == 8) { f[i] = ffma(f[i], fa, fa); const unsigned long long m = mulwide(u[i], ka); u[i] = (uint32_t)(m >> 32); w[i] ^= (uint32_t)m; }
      if (MODE == 9) { if (i < 2) f[i] = mufu_sin(f[i]); p[i] = fma2(p[i], pa, pa); }
      if (MODE == 10) f[i] = fsel(f[i], fa, (int)(u[i] & 1));
      if (MODE == 11) { f[i] = ffma(f[i], fa, fa); p[i] = fma2(p[i], pa, pa); }
      if (MODE == 13) f[i] = mufu_tanh(f[i]);
      if (MODE == 14) u[i] = mufu_tanh_h2(u[i]);
      if (MODE == 15) u[i] = mufu_tanh_b2(u[i]);
      if (MODE == 12) {
        if (i < 4) {
          float4 m = vsm[threadIdx.x];
          unsigned long long a = ((unsigned long long)__float_as_uint(m.y) << 32) | __float_as_uint(m.x);
          unsigned long long b = ((unsigned long long)__float_as_uint(m.w) << 32) | __float_as_uint(m.z);
          a = fma2(p[i], pa, a); b = fma2(p[i], p[i], b);
          vsm[threadIdx.x] = make_float4(__uint_as_float((uint32_t)a), __uint_as_float((uint32_t)(a >> 32)),
                                        __uint_as_float((uint32_t)b), __uint_as_float((uint32_t)(b >> 32)));
        }
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += f[i] + (float)u[i] + (float)w[i] + __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + sm[threadIdx.x].x;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter, float* out, long long* cyc) {
  printf("%-34s", name);
  for (int warps_per_sched : {1, 2, 4, 8}) {
    const int threads = warps_per_sched * 4 * 32, iters = 4096;
    bench<MODE><<<148, threads, threads * sizeof(float4)>>>(out, iters, 0xD2511F53u, 1.0001f, cyc);
    bench<MODE><<<148, threads, threads * sizeof(float4)>>>(out, iters, 0xD2511F53u, 1.0001f, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double instr_per_sched = (double)iters * per_iter * warps_per_sched;
    printf("  w%d: %5.2f", warps_per_sched, c / instr_per_sched);
  }
  printf("   cycles / warp-instr / scheduler\n");
}

int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * sizeof(float));
  long long* cyc; cudaMalloc(&cyc, sizeof(long long));
  run<0>("FFMA", ILP, out, cyc);
  run<1>("FFMA2", ILP, out, cyc);
  run<2>("IMAD.WIDE (+IADD)", 2 * ILP, out, cyc);
  run<3>("LOP3", ILP, out, cyc);
  run<4>("IMAD.WIDE + LOP3 (Philox)", 2 * ILP, out, cyc);
  run<5>("FFMA2 + LOP3", 2 * ILP, out, cyc);
  run<6>("FFMA2 + IMAD.WIDE (+LOP)", 3 * ILP, out, cyc);
  run<8>("FFMA + IMAD.WIDE (+LOP)", 3 * ILP, out, cyc);
  run<7>("MUFU.LG2", ILP, out, cyc);
  run<9>("2 MUFU.SIN + 8 FFMA2", ILP + 2, out, cyc);
  run<10>("SELP (+LOP)", 3 * ILP, out, cyc);
  run<11>("FFMA + FFMA2", 2 * ILP, out, cyc);
  run<12>("4x(LDS128+2FFMA2+STS128)", 16, out, cyc);
  run<13>("MUFU.TANH f32", ILP, out, cyc);
  run<14>("tanh.approx.f16x2 (per pair)", ILP, out, cyc);
  run<15>("tanh.approx.bf16x2 (per pair)", ILP, out, cyc);
  return 0;
}
