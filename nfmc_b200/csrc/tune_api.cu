// tune_api.cu -- warm-up adaptation on the device (reference: MetropolisSampler.update_kernel, mcmc/base.py:142-161).
//
// The reference recomputes, after every warm-up iteration, the across-chain variance of every coordinate
// (torch.var(x.flatten(1,-1), dim=0), unbiased) and moves the inverse-mass diagonal towards it,
//     imd <- c * var + (1 - c) * imd            (c = params.imd_adjustment),
// and feeds the iteration's mean acceptance to the dual-averaging step-size rule.  Here that is two small kernels with
// the collective of SURVEY 8(e)-3 between them:
//   nfmc_chain_sums      sums[0:d] += sum_x, sums[d:2d] += sum_x2 (fp64), sums[2d] = accepted count so far, sums[2d+1] = n
//   (multi-GPU: one all-reduce(SUM) of the 2d+2 doubles, issued by the host mirror over NCCL)
//   nfmc_tune_inv_mass   var from the pooled sums -> EMA into inv_mass_diag (device, fp32)
// so every rank ends up with the same inverse mass and, reading sums[2d] / sums[2d+1], the same acceptance rate.
#include "host_common.cuh"

namespace nfmc {

constexpr int kSumThreads = 256;

// thread c of a CTA owns columns c, c + 256, ...; consecutive threads read consecutive floats of a row (coalesced);
// each CTA takes a contiguous slab of rows, accumulates in fp32 over short runs and in fp64 across them
__global__ void __launch_bounds__(kSumThreads) chain_sums_kernel(const float* __restrict__ x, long long n, int d, double* __restrict__ sums,
                                                                 const unsigned long long* __restrict__ counts) {
  const long long rows_per = (n + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * rows_per, r1 = min(n, r0 + rows_per);
  for (int c = threadIdx.x; c < d; c += kSumThreads) {
    double s1 = 0.0, s2 = 0.0;
    for (long long r = r0; r < r1; r += 64) {
      float a = 0.f, b = 0.f;
      const long long re = min(r1, r + 64);
      for (long long rr = r; rr < re; ++rr) {
        const float v = __ldg(x + rr * d + c);
        a += v;
        b = fmaf(v, v, b);
      }
      s1 += (double)a;
      s2 += (double)b;
    }
    if (r1 > r0) {
      atomicAdd(sums + c, s1);
      atomicAdd(sums + d + c, s2);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    sums[2 * d] = counts ? (double)counts[0] : 0.0;
    sums[2 * d + 1] = (double)n;
  }
}

__global__ void tune_inv_mass_kernel(const double* __restrict__ sums, int d, float c, float* __restrict__ imd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d) return;
  const double N = sums[2 * d + 1];
  if (N < 2.0) return;                                            // mcmc/base.py:150: only with more than one chain
  const double mean = sums[i] / N;
  double var = (sums[d + i] - N * mean * mean) / (N - 1.0);       // unbiased, as torch.var
  if (var < 0.0) var = 0.0;
  imd[i] = (float)((double)c * var + (1.0 - (double)c) * (double)imd[i]);
}

}  // namespace nfmc

using namespace nfmc;

extern "C" int nfmc_chain_sums(const float* x, int64_t n, int32_t d, double* sums, const uint64_t* counts, void* stream) {
  if (!x || !sums || n < 1 || d < 1 || d > NFMC_MAX_DIM) return set_error("chain_sums: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  if (int e = check_cuda(cudaMemsetAsync(sums, 0, (size_t)(2 * d + 2) * sizeof(double), s), "chain_sums memset")) return e;
  long long grid = (n + 255) / 256;
  const long long cap = 4ll * sm_count();
  if (grid > cap) grid = cap;
  chain_sums_kernel<<<(int)grid, kSumThreads, 0, s>>>(x, n, d, sums, reinterpret_cast<const unsigned long long*>(counts));
  return check_cuda(cudaGetLastError(), "chain_sums_kernel launch");
}

extern "C" int nfmc_tune_inv_mass(const double* sums, int32_t d, float imd_adjustment, float* inv_mass_diag, void* stream) {
  if (!sums || !inv_mass_diag || d < 1 || d > NFMC_MAX_DIM) return set_error("tune_inv_mass: bad arguments");
  tune_inv_mass_kernel<<<(d + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, d, imd_adjustment, inv_mass_diag);
  return check_cuda(cudaGetLastError(), "tune_inv_mass_kernel launch");
}
