// Probe: column-alignment rules of tcgen05.ld shapes.  mode: 0 = 32x32b.x16 @col 50, 1 = 32x32b.x8 @col 2, 2 = 32x32b.x8 @col 1,
// 3 = 16x256b.x1 @col 4, 4 = 16x256b.x1 @col 2, 5 = 32x32b.x2 @ col 1
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(int* out, int mode) {
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  const uint32_t trow = base + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < 128; c0 += 16) {
    uint32_t v[16];
    for (int c = 0; c < 16; ++c) v[c] = tid * 1000 + c0 + c;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(trow + c0),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                 "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[16];
  for (int i = 0; i < 16; ++i) r[i] = 0xffffffff;
  if (mode == 0)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(trow + 50) : "memory");
  if (mode == 1 || mode == 2)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(trow + (mode == 1 ? 2 : 1)) : "memory");
  if (mode == 3 || mode == 4)
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(trow + (mode == 3 ? 4 : 2)) : "memory");
  if (mode == 5)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(trow + 1) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 16; ++i) out[tid * 16 + i] = r[i];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(128u) : "memory");
}
int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  int* d; cudaMalloc(&d, 128 * 16 * 4);
  probe<<<1, 128>>>(d, mode);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("mode %d: error %s\n", mode, cudaGetErrorString(e)); return 1; }
  static int h[128 * 16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("mode %d ok:", mode);
  for (int t = 0; t < 6; t += 5) { printf(" [t%d:", t); for (int i = 0; i < 16; ++i) printf(" %d", h[t * 16 + i]); printf("]"); }
  printf("\n");
  return 0;
}
