// Probe: per-SM latency / throughput of 1-D TMA bulk copies (cp.async.bulk) with all 148 SMs active.
//   store: smem -> global, `bytes` per copy, wait_group.read 0 after each (serial) or after all (pipelined)
//   load : global -> smem with an mbarrier, serial
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(float* g, long long stride_floats, int bytes, int reps, int mode, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  const uint32_t b = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    float* base = g + (long long)blockIdx.x * stride_floats;
    const long long t0 = clock64();
    uint32_t ph = 0;
    for (int r = 0; r < reps; ++r) {
      float* p = base + (long long)(r % 8) * (bytes / 4);
      if (mode == 0 || mode == 1) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p), "r"(smem_u32(smem)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (mode == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      } else {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)), "l"(p), "r"(bytes), "r"(b) : "memory");
        uint32_t done = 0;
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b), "r"(ph) : "memory");
        ph ^= 1;
      }
    }
    if (mode == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    const long long t1 = clock64();
    if (mode <= 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    const long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
}
int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 148;
  const long long stride = 1 << 20;   // 4 MB per CTA
  float* g; cudaMalloc(&g, (size_t)grid * stride * 4);
  long long* out; cudaMalloc(&out, grid * 16);
  static long long h[2 * 1024];
  const int sizes[4] = {4096, 25600, 51200, 102400};
  const char* names[3] = {"store serial (wait_group.read each)", "store pipelined", "load serial"};
  for (int mode = 0; mode < 3; ++mode)
    for (int si = 0; si < 4; ++si) {
      const int bytes = sizes[si], reps = 32;
      cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
      probe<<<grid, 128, 110 * 1024>>>(g, stride, bytes, 4, mode, out);   // warm
      probe<<<grid, 128, 110 * 1024>>>(g, stride, bytes, reps, mode, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, out, grid * 16, cudaMemcpyDeviceToHost);
      double a = 0, b = 0;
      for (int i = 0; i < grid; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
      a /= grid; b /= grid;
      printf("grid %d %-38s %6d B: %8.0f cycles/copy (%.1f B/clk/SM), to full completion %8.0f cycles/copy\n", grid, names[mode], bytes, a / reps,
             bytes / (a / reps), b / reps);
    }
  return 0;
}
