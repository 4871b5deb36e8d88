// flow_args.cuh -- launch arguments, shared-memory plan and base-draw helper shared by the flow kernels.
#pragma once
#include "chain_kernel.cuh"
#include "flow.cuh"
#include "host_common.cuh"

namespace nfmc {

struct FlowArgs {
  int d, gs, Lc, M, H;
  const float* blob;
  long long blob_floats;
  int stage_blob;  // copy the blob to shared memory first
  int exact;       // exact lane layout (host_common.cuh: Layout::exact)
  int mom_slots;   // per-lane running-moment slots (float4 each) carved after the scratch, 0 if none
};

// shared-memory plan of the flow kernels: [cta stats] [blob (optional)] [scratch per group]
struct FlowSmem {
  CtaStats st;
  FlowDesc F;
  float* scr;
  float4* mom;   // this lane's running moments, slot e at mom[e * kThreads]
};
template <bool SB>
__device__ __forceinline__ FlowSmem flow_smem_init(unsigned char* smem, const FlowArgs& A, bool with_stats) {
  FlowSmem S;
  size_t off = 0;
  if (with_stats) {
    S.st = cta_stats_init(smem, A.d);
    off = (cta_stats_bytes(A.d) + 15) & ~size_t(15);
  }
  float* fbase = reinterpret_cast<float*>(smem + off);
  const float* blob = A.blob;
  if (SB) {
    // 128-bit copies, four in flight per thread: the scalar load -> store loop it replaces waited one global-memory latency per
    // element (30 in a row for a d = 100 default flow) and was the largest single stall site of the jump kernels (ncu source
    // page, profiles/jpa_r02_*: 8-11 % of all stall samples on its STS)
    const int nf = (int)A.blob_floats;
    int done = 0;
    if ((reinterpret_cast<uintptr_t>(A.blob) & 15) == 0) {
      const float4* src = reinterpret_cast<const float4*>(A.blob);
      float4* dst = reinterpret_cast<float4*>(fbase);
      const int n4 = nf >> 2;
#pragma unroll 4
      for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = __ldg(src + i);
      done = n4 << 2;
    }
    for (int i = done + threadIdx.x; i < nf; i += blockDim.x) fbase[i] = __ldg(A.blob + i);
    blob = fbase;
    fbase += ((int)A.blob_floats + 3) & ~3;
    __syncthreads();
  }
  S.F = make_flow_desc(blob, A.d, A.Lc, A.M, A.H);
  if (SB) S.F.sbase = (unsigned)__cvta_generic_to_shared(blob);
  S.scr = fbase + (size_t)(threadIdx.x / A.gs) * S.F.scratch;
  float* after = fbase + (size_t)(kThreads / A.gs) * S.F.scratch;
  after = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(after) + 15) & ~uintptr_t(15));
  S.mom = reinterpret_cast<float4*>(after) + threadIdx.x;
  return S;
}

// draw the base sample for (chain, step) into (lo, hi) in PHYSICAL order; returns the accept-uniform bits.
// Injected normals are the LOGICAL z the reference's flow.sample would have drawn (flipped if Lc is odd).
template <int E>
__device__ __forceinline__ uint32_t draw_base(const RngArgs& R, const Geom& g, bool flip, long long n, long long chain,
                                              long long chain0, int k, float (&lo)[E], float (&hi)[E]) {
  if (R.normals) {
    const float* nr = R.normals + ((long long)k * n + chain) * (long long)g.d;
    if (flip) load_chain_flipped(nr, g, lo, hi);
    else load_chain(nr, g, lo, hi);
    return 0u;
  }
  StepNoise<E> nz;
  const RngKey key = make_rng_key(R.seed, 1u, R.step0 + (uint64_t)k, (uint64_t)(chain0 + chain));
  draw_step_noise<E>(key, g.j, nz);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int kk = g.j + g.gs * e;
    lo[e] = kk < g.da ? nz.lo[e] : 0.f;
    hi[e] = kk < g.db ? nz.hi[e] : 0.f;
  }
  return nz.ubits;
}

// ---- host side ----------------------------------------------------------------------------------------------
inline int validate_flow(const nfmc_realnvp* f) {
  if (!f || !f->blob) return set_error("flow is NULL");
  if (f->d < 2 || f->d > NFMC_MAX_DIM) return set_error("flow: d out of range [2, 1024]");
  if (f->n_coupling < 0 || f->n_linear < 1 || f->hidden < 1) return set_error("flow: bad n_coupling / n_linear / hidden");
  if (f->blob_floats != flow_blob_floats(f->d, f->n_coupling, f->n_linear, f->hidden))
    return set_error("flow: blob_floats does not match (d, n_coupling, n_linear, hidden)");
  return 0;
}

// decide shared-memory plan; returns bytes, sets A.stage_blob
inline size_t plan_flow_smem(FlowArgs& A, const nfmc_realnvp* f, const Layout& L, bool with_stats, bool with_moments = false) {
  A.d = f->d; A.gs = L.gs; A.Lc = f->n_coupling; A.M = f->n_linear; A.H = f->hidden;
  A.blob = f->blob; A.blob_floats = f->blob_floats;
  const FlowDesc F = make_flow_desc(nullptr, A.d, A.Lc, A.M, A.H);
  size_t base = with_stats ? ((cta_stats_bytes_host(A.d) + 15) & ~size_t(15)) : 0;
  const size_t scratch = (size_t)(kThreads / L.gs) * F.scratch * sizeof(float);
  const size_t blob_b = (size_t)((A.blob_floats + 3) & ~3ll) * sizeof(float);
  A.exact = L.exact ? 1 : 0;
  A.mom_slots = with_moments ? L.E : 0;
  const size_t mom_b = with_moments ? (size_t)L.E * kThreads * sizeof(float4) + 16 : 0;
  A.stage_blob = (base + scratch + blob_b + mom_b <= 72 * 1024) ? 1 : 0;
  return base + scratch + (A.stage_blob ? blob_b : 0) + mom_b;
}


}  // namespace nfmc
