// Probe: which (lane, column) of tensor memory lands in which (thread, register) for the tcgen05.ld / st shapes.
// 128 threads write value = lane * 1000 + column with the 32x32b shape (thread t <-> lane t); every warp then reads its
// lane quarter back with the 16-lane shapes and prints the mapping of warp 0 and warp 1.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(int* out, int mode) {
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  const uint32_t trow = base + ((uint32_t)(warp * 32) << 16);
  uint32_t v[16];
  for (int c = 0; c < 16; ++c) v[c] = tid * 1000 + c;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(trow),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
               "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[8];
  if (mode == 0) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(trow) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 4; ++i) out[((0 * 128 + tid) * 8) + i] = r[i];
  }
  if (mode == 1) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(trow + (16u << 16)) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 4; ++i) out[((1 * 128 + tid) * 8) + i] = r[i];
  }
  if (mode == 2) {
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(trow) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 2; ++i) out[((2 * 128 + tid) * 8) + i] = r[i];
  }
  if (mode == 3) {
  asm volatile("tcgen05.ld.sync.aligned.16x64b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(trow) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  out[((3 * 128 + tid) * 8)] = r[0];
  }
  if (mode == 4) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(trow) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 8; ++i) out[((4 * 128 + tid) * 8) + i] = r[i];
  }
  if (mode == 5) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(trow + 3) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 4; ++i) out[((5 * 128 + tid) * 8) + i] = r[i];
  }
  if (mode == 6) {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1,%2,%3,%4};" ::"r"(trow + 8), "r"(7000000 + tid * 10 + 0), "r"(7000000 + tid * 10 + 1),
               "r"(7000000 + tid * 10 + 2), "r"(7000000 + tid * 10 + 3) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(trow + 8) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 8; ++i) out[((6 * 128 + tid) * 8) + i] = r[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(32u) : "memory");
}
int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  int* d; cudaMalloc(&d, 7 * 128 * 8 * 4); cudaMemset(d, 0xff, 7 * 128 * 8 * 4);
  probe<<<1, 128>>>(d, mode);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  static int h[7 * 128 * 8]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[7] = {"ld16x256b.x1 @lane0", "ld16x256b.x1 @lane16", "ld16x128b.x1", "ld16x64b.x1", "ld16x256b.x2", "ld16x256b.x1 @col3", "st16x256b.x1@col8 -> ld32x32b.x8"};
  const int nreg[7] = {4, 4, 2, 1, 8, 4, 8};
  for (int s = mode; s <= mode; ++s) {
    printf("== %s (value = lane*1000+col)\n", names[s]);
    for (int t = 0; t < 64; ++t) {
      if (t >= 34 && s != 6) break;
      printf("t%02d:", t);
      for (int i = 0; i < nreg[s]; ++i) printf(" %d", h[(s * 128 + t) * 8 + i]);
      printf("\n");
    }
  }
  return 0;
}
