// local_api.cu -- C-ABI entry points of the local kernels + potential evaluation + RNG dump kernels.
#include "launchers.cuh"

namespace nfmc {

// ---------------------------------------------------------------------------------------------------------
template <int POT, int E>
__global__ void __launch_bounds__(kThreads) potential_kernel(PotParams P, const float* __restrict__ x, float* __restrict__ u,
                                                            float* __restrict__ grad, long long n, int d, int gs) {
  const Geom g = make_geom(d, gs);
  const int cpc = kThreads / gs;
  const long long tiles = (n + cpc - 1) / cpc;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long chain_raw = tile * cpc + threadIdx.x / gs;
    const bool active = chain_raw < n;
    const long long chain = active ? chain_raw : n - 1;
    float lo[E], hi[E];
    load_chain(x + chain * (long long)d, g, lo, hi);
    const PotCtx c = pot_prepare<POT, E>(P, g, lo, hi);
    if (active && g.j == 0) u[chain] = c.u;
    if (grad && active) {
      float glo[E], ghi[E];
#pragma unroll
      for (int e = 0; e < E; ++e) pot_grad<POT>(P, c, g, g.j + g.gs * e, lo[e], hi[e], glo[e], ghi[e]);
      store_chain(grad + chain * (long long)d, g, glo, ghi);
    }
  }
}

template <int E>
__global__ void __launch_bounds__(kThreads) rng_fill_kernel(RngArgs R, unsigned stream_id, long long chain0, int d, int gs,
                                                           long long n, int n_steps, float* normals, float* uniforms) {
  const Geom g = make_geom(d, gs);
  const int cpc = kThreads / gs;
  const long long tiles = (n + cpc - 1) / cpc;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long chain_raw = tile * cpc + threadIdx.x / gs;
    const bool active = chain_raw < n;
    const long long chain = active ? chain_raw : n - 1;
    for (int k = 0; k < n_steps; ++k) {
      StepNoise<E> nz;
      const RngKey key = make_rng_key(R.seed, stream_id, R.step0 + (uint64_t)k, (uint64_t)(chain0 + chain));
      draw_step_noise<E>(key, g.j, nz);
      if (active) {
        if (normals) store_chain(normals + ((long long)k * n + chain) * (long long)d, g, nz.lo, nz.hi);
        if (uniforms && g.j == 0) uniforms[(long long)k * n + chain] = uniform_from_bits(nz.ubits);
      }
    }
  }
}

static size_t local_smem_bytes(int d, bool with_mass, size_t per_dim, int E) {
  size_t b = (cta_stats_bytes_host(d) + 15) & ~size_t(15);
  if (with_mass) b += (size_t)d * per_dim;
  b = (b + 15) & ~size_t(15);
  b += (size_t)E * kThreads * sizeof(float4);  // per-lane running moments
  return b;
}

// Injected random numbers: an ADJUSTED kernel consumes normals [steps, n, d] and uniforms [steps, n] together -- with only
// the normals injected the accept uniform would silently be 0 (always accept), with only the uniforms the proposal noise
// would come from Philox against the caller's intent.  Unadjusted kernels take normals alone.
int validate_injected(const nfmc_rng* rng, int adjusted, const char* who) {
  if (!rng) return 0;
  const bool zn = rng->normals != nullptr, un = rng->uniforms != nullptr;
  if (adjusted ? (zn != un) : (un && !zn))
    return set_error(std::string(who) + ": inject both normals [steps,n,d] and uniforms [steps,n], or neither");
  return 0;
}

static int fill_chain_args(ChainArgs& C, const nfmc_potential* pot, float* x, int64_t n, int32_t n_steps,
                           const nfmc_rng* rng, int64_t chain0, const nfmc_stats* stats, const nfmc_sink* sink,
                           const Layout& L) {
  C.pot = pot_params(pot);
  C.x = x; C.n = n; C.chain0 = chain0; C.d = pot->d; C.gs = L.gs; C.n_steps = n_steps;
  C.rng.seed = rng ? rng->seed : 0; C.rng.step0 = rng ? rng->step0 : 0;
  C.rng.normals = rng ? rng->normals : nullptr; C.rng.uniforms = rng ? rng->uniforms : nullptr;
  C.stats.sum_x = stats ? stats->sum_x : nullptr; C.stats.sum_x2 = stats ? stats->sum_x2 : nullptr;
  C.stats.counts = stats ? stats->counts : nullptr;
  C.sink.samples = sink ? sink->samples : nullptr; C.sink.seen0 = sink ? sink->seen0 : 0;
  C.sink.thinning = (sink && sink->thinning > 0) ? sink->thinning : 1;
  return 0;
}

}  // namespace nfmc

using namespace nfmc;


extern "C" int nfmc_mala_steps(const nfmc_potential* pot, float* x, int64_t n, int32_t n_steps, float step_size,
                               const float* inv_mass_diag, int32_t adjusted, const nfmc_rng* rng, int64_t chain0,
                               const nfmc_stats* stats, const nfmc_sink* sink, void* stream) {
  if (int e = validate_pot(pot)) return e;
  if (!x || n < 1 || n_steps < 0) return set_error("mala_steps: bad x/n/n_steps");
  if (int e = validate_injected(rng, adjusted, "mala_steps")) return e;
  if (!(step_size > 0.f)) return set_error("mala_steps: step_size must be positive");
  if (n_steps == 0) return 0;
  Layout L;
  if (!layout_for_dim(pot->d, L)) return set_error("mala_steps: unsupported event size");
  LocalArgs A;
  fill_chain_args(A.c, pot, x, n, n_steps, rng, chain0, stats, sink, L);
  A.tau = step_size; A.sqrt_2tau = (float)sqrt(2.0 * (double)step_size); A.imd = inv_mass_diag;
  A.adjusted = adjusted; A.n_leapfrog = 0; A.random_walk = 0;
  const size_t smem = local_smem_bytes(pot->d, inv_mass_diag != nullptr, sizeof(float4), L.E);
  const int grid = grid_for(n, L.gs, 4);
  cudaStream_t s = (cudaStream_t)stream;
  NFMC_DISPATCH_E(L.E, { return launch_mala<E>(pot->kind, L.exact, A, grid, smem, s); });
  return 0;
}

extern "C" int nfmc_mh_steps(const nfmc_potential* pot, float* x, int64_t n, int32_t n_steps, const float* inv_mass_diag,
                             int32_t adjusted, const nfmc_rng* rng, int64_t chain0, const nfmc_stats* stats,
                             const nfmc_sink* sink, void* stream) {
  if (int e = validate_pot(pot)) return e;
  if (!x || n < 1 || n_steps < 0) return set_error("mh_steps: bad x/n/n_steps");
  if (int e = validate_injected(rng, adjusted, "mh_steps")) return e;
  if (n_steps == 0) return 0;
  Layout L;
  if (!layout_for_dim(pot->d, L)) return set_error("mh_steps: unsupported event size");
  LocalArgs A;
  fill_chain_args(A.c, pot, x, n, n_steps, rng, chain0, stats, sink, L);
  A.tau = 1.f; A.sqrt_2tau = 1.f; A.imd = inv_mass_diag; A.adjusted = adjusted; A.n_leapfrog = 0; A.random_walk = 1;
  const size_t smem = local_smem_bytes(pot->d, inv_mass_diag != nullptr, sizeof(float4), L.E);
  const int grid = grid_for(n, L.gs, 4);
  cudaStream_t s = (cudaStream_t)stream;
  NFMC_DISPATCH_E(L.E, { return launch_mala<E>(pot->kind, L.exact, A, grid, smem, s); });
  return 0;
}

extern "C" int nfmc_ess_steps(const nfmc_potential* nll, float* x, int64_t n, int32_t n_steps, int32_t max_iterations,
                              const nfmc_rng* rng, int64_t chain0, const nfmc_stats* stats, const nfmc_sink* sink,
                              void* stream) {
  if (int e = validate_pot(nll)) return e;
  if (!x || n < 1 || n_steps < 0 || max_iterations < 0) return set_error("ess_steps: bad x/n/n_steps/max_iterations");
  if (rng && ((rng->normals == nullptr) != (rng->uniforms == nullptr)))
    return set_error("ess_steps: inject both normals [steps,n,d] and uniforms [steps,n,2+max_iterations], or neither");
  if (n_steps == 0) return 0;
  Layout L;
  if (!layout_for_dim(nll->d, L)) return set_error("ess_steps: unsupported event size");
  EssArgs A;
  fill_chain_args(A.c, nll, x, n, n_steps, rng, chain0, stats, sink, L);
  A.max_iterations = max_iterations;
  const size_t smem = local_smem_bytes(nll->d, false, 0, L.E);
  const int grid = grid_for(n, L.gs, 4);
  cudaStream_t s = (cudaStream_t)stream;
  NFMC_DISPATCH_E(L.E, { return launch_ess<E>(nll->kind, L.exact, A, grid, smem, s); });
  return 0;
}

extern "C" int nfmc_hmc_steps(const nfmc_potential* pot, float* x, int64_t n, int32_t n_steps, float step_size,
                              int32_t n_leapfrog, const float* inv_mass_diag, int32_t adjusted, const nfmc_rng* rng,
                              int64_t chain0, const nfmc_stats* stats, const nfmc_sink* sink, void* stream) {
  if (int e = validate_pot(pot)) return e;
  if (!x || n < 1 || n_steps < 0 || n_leapfrog < 0) return set_error("hmc_steps: bad x/n/n_steps/n_leapfrog");
  if (int e = validate_injected(rng, adjusted, "hmc_steps")) return e;
  if (n_steps == 0) return 0;
  Layout L;
  if (!layout_for_dim(pot->d, L)) return set_error("hmc_steps: unsupported event size");
  LocalArgs A;
  fill_chain_args(A.c, pot, x, n, n_steps, rng, chain0, stats, sink, L);
  A.tau = step_size; A.sqrt_2tau = 0.f; A.imd = inv_mass_diag; A.adjusted = adjusted; A.n_leapfrog = n_leapfrog; A.random_walk = 0;
  size_t smem = local_smem_bytes(pot->d, inv_mass_diag != nullptr, sizeof(float4), L.E);
  // the packed kernel stages a diagonal Gaussian's per-lane precisions and negated means: 2 x [E][32] float2
  if (pot->kind == NFMC_POT_DIAG_GAUSSIAN) smem += (size_t)2 * L.E * 32 * sizeof(float2);
  const int grid = grid_for(n, L.gs, 4);
  cudaStream_t s = (cudaStream_t)stream;
  NFMC_DISPATCH_E(L.E, { return launch_hmc<E>(pot->kind, L.exact, A, grid, smem, s); });
  return 0;
}

extern "C" int nfmc_potential_eval(const nfmc_potential* pot, const float* x, float* u, float* grad, int64_t n, void* stream) {
  if (int e = validate_pot(pot)) return e;
  if (!x || !u || n < 1) return set_error("potential_eval: bad arguments");
  Layout L;
  if (!layout_for_dim(pot->d, L)) return set_error("potential_eval: unsupported event size");
  const PotParams P = pot_params(pot);
  cudaStream_t s = (cudaStream_t)stream;
  NFMC_DISPATCH_POT(pot->kind, NFMC_DISPATCH_E(L.E, {
    potential_kernel<POT, E><<<occupancy_grid(potential_kernel<POT, E>, 0, n, L.gs), kThreads, 0, s>>>(P, x, u, grad, n, pot->d, L.gs);
  }));
  return check_cuda(cudaGetLastError(), "potential_kernel launch");
}

extern "C" int nfmc_rng_fill(const nfmc_rng* rng, int32_t stream_id, int64_t chain0, int32_t d, int64_t n, int32_t n_steps,
                             float* normals, float* uniforms, void* stream) {
  if (!rng || n < 1 || n_steps < 1) return set_error("rng_fill: bad arguments");
  Layout L;
  if (!layout_for_dim(d, L)) return set_error("rng_fill: unsupported event size");
  RngArgs R{rng->seed, rng->step0, nullptr, nullptr};
  cudaStream_t s = (cudaStream_t)stream;
  NFMC_DISPATCH_E(L.E, { rng_fill_kernel<E><<<occupancy_grid(rng_fill_kernel<E>, 0, n, L.gs), kThreads, 0, s>>>(R, (unsigned)stream_id, chain0, d, L.gs, n, n_steps, normals, uniforms); });
  return check_cuda(cudaGetLastError(), "rng_fill_kernel launch");
}
