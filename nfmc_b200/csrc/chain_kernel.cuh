// chain_kernel.cuh -- pieces shared by every per-chain kernel: launch arguments, CTA statistics, sample sink.
#pragma once
#include "common.cuh"
#include "potentials.cuh"

namespace nfmc {

struct RngArgs {
  uint64_t seed;
  uint64_t step0;
  const float* normals;   // [steps, n, d] or nullptr
  const float* uniforms;  // [steps, n] or nullptr
};
struct StatsArgs {
  double* sum_x;
  double* sum_x2;
  unsigned long long* counts;
};
struct SinkArgs {
  float* samples;  // [rows, n, d] or nullptr
  long long seen0;
  int thinning;
};

struct ChainArgs {
  PotParams pot;
  float* x;
  long long n;
  long long chain0;
  int d, gs;
  int n_steps;
  RngArgs rng;
  StatsArgs stats;
  SinkArgs sink;
};

// dynamic shared memory carve-up: [double sx[d]] [double sx2[d]] [unsigned long long cnt[4]] [kernel-specific ...]
struct CtaStats {
  double* sx;
  double* sx2;
  unsigned long long* cnt;
};
__device__ __forceinline__ size_t cta_stats_bytes(int d) { return (size_t)(2 * d) * sizeof(double) + 4 * sizeof(unsigned long long); }
inline size_t cta_stats_bytes_host(int d) { return (size_t)(2 * d) * sizeof(double) + 4 * sizeof(unsigned long long); }

__device__ __forceinline__ CtaStats cta_stats_init(unsigned char* smem, int d) {
  CtaStats s;
  s.sx = reinterpret_cast<double*>(smem);
  s.sx2 = s.sx + d;
  s.cnt = reinterpret_cast<unsigned long long*>(s.sx2 + d);
  for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) s.sx[i] = 0.0;
  if (threadIdx.x < 4) s.cnt[threadIdx.x] = 0ull;
  __syncthreads();
  return s;
}
__device__ __forceinline__ void cta_stats_finish(const CtaStats& s, const StatsArgs& out, int d) {
  __syncthreads();
  if (out.sum_x && out.sum_x2) {
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
      atomicAdd(out.sum_x + i, s.sx[i]);
      atomicAdd(out.sum_x2 + i, s.sx2[i]);
    }
  }
  if (out.counts && threadIdx.x < 4 && s.cnt[threadIdx.x]) atomicAdd(out.counts + threadIdx.x, s.cnt[threadIdx.x]);
}

// write the state as sample row for step k of this launch if the thinning rule keeps it
template <int E>
__device__ __forceinline__ void sink_store(const SinkArgs& sk, const Geom& g, long long n, long long row_chain, int k,
                                           const float (&lo)[E], const float (&hi)[E]) {
  const long long idx = sk.seen0 + k;
  if (idx % sk.thinning != 0) return;
  const long long first = (sk.seen0 + sk.thinning - 1) / sk.thinning;
  const long long r = idx / sk.thinning - first;
  store_chain(sk.samples + (r * n + row_chain) * (long long)g.d, g, lo, hi);
}

template <int E>
__device__ __forceinline__ void accumulate_moments(const float (&lo)[E], const float (&hi)[E], float (&m1lo)[E],
                                                   float (&m1hi)[E], float (&m2lo)[E], float (&m2hi)[E]) {
#pragma unroll
  for (int e = 0; e < E; ++e) {
    m1lo[e] += lo[e];
    m1hi[e] += hi[e];
    m2lo[e] = fmaf(lo[e], lo[e], m2lo[e]);
    m2hi[e] = fmaf(hi[e], hi[e], m2hi[e]);
  }
}

}  // namespace nfmc
