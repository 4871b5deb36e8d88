// flow_api.cu -- C-ABI entry points of the flow / jump / IMH / NeuTra kernels (dispatch over E).
#include "launchers.cuh"

using namespace nfmc;

#ifndef NFMC_NEUTRA_CTAS
#define NFMC_NEUTRA_CTAS 2   // resident CTAs per SM of neutra_hmc_kernel (its __launch_bounds__)
#endif
#ifndef NFMC_JUMP_CTAS
#define NFMC_JUMP_CTAS 3     // ... of jump_kernel
#endif


static int flow_pass(const nfmc_realnvp* flow, int mode, const float* in, float* out, float* aux, int64_t n, void* stream) {
  if (int e = validate_flow(flow)) return e;
  if (!in || n < 1) return set_error("flow pass: bad arguments");
  Layout L;
  if (!layout_for_dim(flow->d, L)) return set_error("flow pass: unsupported event size");
  FlowArgs A;
  const size_t smem = plan_flow_smem(A, flow, L, false);
  const int grid = grid_for(n, L.gs, 4);
  cudaStream_t s = (cudaStream_t)stream;
  NFMC_DISPATCH_E(L.E, { return launch_flow_pass<E>(A, mode, in, out, aux, n, grid, smem, s); });
  return 0;
}

extern "C" int nfmc_realnvp_forward(const nfmc_realnvp* flow, const float* x, float* z, float* log_det, int64_t n, void* stream) {
  return flow_pass(flow, PASS_FORWARD, x, z, log_det, n, stream);
}
extern "C" int nfmc_realnvp_inverse(const nfmc_realnvp* flow, const float* z, float* x, float* log_det, int64_t n, void* stream) {
  return flow_pass(flow, PASS_INVERSE, z, x, log_det, n, stream);
}
extern "C" int nfmc_flow_log_prob(const nfmc_realnvp* flow, const float* x, float* log_q, int64_t n, void* stream) {
  if (!log_q) return set_error("flow_log_prob: log_q is NULL");
  return flow_pass(flow, PASS_LOGPROB, x, nullptr, log_q, n, stream);
}

extern "C" int nfmc_flow_sample(const nfmc_realnvp* flow, const nfmc_rng* rng, int64_t chain0, float* x, float* log_q,
                                int64_t n, void* stream) {
  if (int e = validate_flow(flow)) return e;
  if (!x || !rng || n < 1) return set_error("flow_sample: bad arguments");
  Layout L;
  if (!layout_for_dim(flow->d, L)) return set_error("flow_sample: unsupported event size");
  FlowArgs A;
  const size_t smem = plan_flow_smem(A, flow, L, false);
  RngArgs R{rng->seed, rng->step0, rng->normals, rng->uniforms};
  const int grid = grid_for(n, L.gs, 4);
  cudaStream_t s = (cudaStream_t)stream;
  NFMC_DISPATCH_E(L.E, { return launch_flow_sample<E>(A, R, chain0, x, log_q, n, grid, smem, s); });
  return 0;
}

namespace nfmc { int validate_injected(const nfmc_rng* rng, int adjusted, const char* who); }

static int launch_jump(const nfmc_potential* pot, const nfmc_realnvp* flow, float* x, float* logq_x, int64_t n,
                       int32_t n_steps, int32_t recompute_logq, int32_t adjusted, const nfmc_rng* rng, int64_t chain0,
                       const nfmc_stats* stats, const nfmc_sink* sink, void* stream) {
  if (int e = validate_pot(pot)) return e;
  if (int e = validate_flow(flow)) return e;
  if (pot->d != flow->d) return set_error("jump: potential and flow event sizes differ");
  if (!x || n < 1 || n_steps < 0) return set_error("jump: bad x/n/n_steps");
  if (int e = validate_injected(rng, adjusted, "jump")) return e;
  if (n_steps == 0) return 0;
  Layout L;
  if (!layout_for_dim(pot->d, L)) return set_error("jump: unsupported event size");
  JumpArgs A;
  A.c.pot = pot_params(pot);
  A.c.x = x; A.c.n = n; A.c.chain0 = chain0; A.c.d = pot->d; A.c.gs = L.gs; A.c.n_steps = n_steps;
  A.c.rng = RngArgs{rng ? rng->seed : 0, rng ? rng->step0 : 0, rng ? rng->normals : nullptr, rng ? rng->uniforms : nullptr};
  A.c.stats = StatsArgs{stats ? stats->sum_x : nullptr, stats ? stats->sum_x2 : nullptr, stats ? stats->counts : nullptr};
  A.c.sink = SinkArgs{sink ? sink->samples : nullptr, sink ? sink->seen0 : 0, (sink && sink->thinning > 0) ? sink->thinning : 1};
  A.logq_x = logq_x; A.recompute_logq = recompute_logq; A.adjusted = adjusted;
  const size_t smem = plan_flow_smem(A.f, flow, L, true, true);
  const int grid = grid_for(n, L.gs, NFMC_JUMP_CTAS);
  cudaStream_t s = (cudaStream_t)stream;
  A.pot_kind = pot->kind;
  NFMC_DISPATCH_E(L.E, { return launch_jump<E>(A, grid, smem, s); });
  return 0;
}

extern "C" int nfmc_jump_step(const nfmc_potential* pot, const nfmc_realnvp* flow, float* x, int64_t n, int32_t adjusted,
                              const nfmc_rng* rng, int64_t chain0, const nfmc_stats* stats, const nfmc_sink* sink, void* stream) {
  return launch_jump(pot, flow, x, nullptr, n, 1, 1, adjusted, rng, chain0, stats, sink, stream);
}

// The NF jump as two kernels: log q(x) by a forward pass into `logq_scratch` [n], then proposal + accept with x loaded
// after the inverse pass (flow_kernels.cu: jump_propose_accept_kernel).  Same results as nfmc_jump_step.
static int jump_second_half(const nfmc_potential* pot, const nfmc_realnvp* flow, float* x, float* logq, int64_t n, int32_t adjusted,
                            const nfmc_rng* rng, int64_t chain0, const nfmc_stats* stats, const nfmc_sink* sink, void* stream) {
  Layout L;
  if (!layout_for_dim(pot->d, L)) return set_error("jump: unsupported event size");
  JumpArgs A;
  A.c.pot = pot_params(pot);
  A.c.x = x; A.c.n = n; A.c.chain0 = chain0; A.c.d = pot->d; A.c.gs = L.gs; A.c.n_steps = 1;
  A.c.rng = RngArgs{rng ? rng->seed : 0, rng ? rng->step0 : 0, rng ? rng->normals : nullptr, rng ? rng->uniforms : nullptr};
  A.c.stats = StatsArgs{stats ? stats->sum_x : nullptr, stats ? stats->sum_x2 : nullptr, stats ? stats->counts : nullptr};
  A.c.sink = SinkArgs{sink ? sink->samples : nullptr, sink ? sink->seen0 : 0, (sink && sink->thinning > 0) ? sink->thinning : 1};
  A.logq_x = logq; A.recompute_logq = 0; A.adjusted = adjusted;
  A.pot_kind = pot->kind;
  const size_t smem = plan_flow_smem(A.f, flow, L, true, true);
  const int grid = grid_for(n, L.gs, 3);
  cudaStream_t s = (cudaStream_t)stream;
  NFMC_DISPATCH_E(L.E, { return launch_jump_propose_accept<E>(A, grid, smem, s); });
  return 0;
}

extern "C" int nfmc_jump_step2(const nfmc_potential* pot, const nfmc_realnvp* flow, float* x, float* logq_scratch, int64_t n,
                               int32_t adjusted, const nfmc_rng* rng, int64_t chain0, const nfmc_stats* stats,
                               const nfmc_sink* sink, void* stream) {
  if (int e = validate_pot(pot)) return e;
  if (int e = validate_flow(flow)) return e;
  if (pot->d != flow->d) return set_error("jump_step2: potential and flow event sizes differ");
  if (!x || n < 1 || (adjusted && !logq_scratch)) return set_error("jump_step2: bad x / n / logq_scratch");
  if (int e = validate_injected(rng, adjusted, "jump_step2")) return e;
  if (adjusted)
    if (int e = flow_pass(flow, PASS_LOGPROB, x, nullptr, logq_scratch, n, stream)) return e;       // jump.py:218
  return jump_second_half(pot, flow, x, logq_scratch, n, adjusted, rng, chain0, stats, sink, stream);
}

extern "C" int nfmc_imh_steps(const nfmc_potential* pot, const nfmc_realnvp* flow, float* x, float* log_q_x, int64_t n,
                              int32_t n_steps, int32_t recompute_logq, const nfmc_rng* rng, int64_t chain0,
                              const nfmc_stats* stats, const nfmc_sink* sink, void* stream) {
  if (!recompute_logq && !log_q_x) return set_error("imh_steps: log_q_x is required unless recompute_logq");
  // Large batches: one launch (pair) per iteration of the spill-free two-kernel form -- x makes an HBM round trip per
  // iteration (8*d bytes per chain, ~0.1 ms per 2^20 chains) instead of staying in registers, but the kernel runs without
  // spills: 0.9 ms (cached log q) / 1.6 ms (recomputed) per iteration against 1.2 / 3.0 ms for the state-resident kernel.
  // Small batches are launch-bound and keep the single state-resident launch.
  if (int e = validate_pot(pot)) return e;
  if (int e = validate_flow(flow)) return e;
  if (pot->d != flow->d) return set_error("imh_steps: potential and flow event sizes differ");
  if (!x || n < 1 || n_steps < 0) return set_error("imh_steps: bad x/n/n_steps");
  if (int e = validate_injected(rng, 1, "imh_steps")) return e;
  Layout L;
  if (log_q_x && layout_for_dim(pot->d, L) && n * L.gs / kThreads >= 6ll * sm_count()) {
    for (int32_t i = 0; i < n_steps; ++i) {
      nfmc_rng r{rng ? rng->seed : 0, (rng ? rng->step0 : 0) + (uint64_t)i,
                 (rng && rng->normals) ? rng->normals + (size_t)i * n * pot->d : nullptr,
                 (rng && rng->uniforms) ? rng->uniforms + (size_t)i * n : nullptr};
      nfmc_sink sk{nullptr, 0, 1};
      const nfmc_sink* skp = nullptr;
      if (sink && sink->samples) {                         // rows written before step i of this call
        const int64_t th = sink->thinning > 0 ? sink->thinning : 1;
        const int64_t rows_before = (sink->seen0 + i + th - 1) / th - (sink->seen0 + th - 1) / th;
        sk.samples = sink->samples + (size_t)rows_before * n * pot->d;
        sk.seen0 = sink->seen0 + i;
        sk.thinning = (int32_t)th;
        skp = &sk;
      }
      if (recompute_logq) {
        if (int e = flow_pass(flow, PASS_LOGPROB, x, nullptr, log_q_x, n, stream)) return e;      // imh.py:133-134
      }
      if (int e = jump_second_half(pot, flow, x, log_q_x, n, 1, &r, chain0, stats, skp, stream)) return e;
    }
    return 0;
  }
  return launch_jump(pot, flow, x, log_q_x, n, n_steps, recompute_logq, 1, rng, chain0, stats, sink, stream);
}

extern "C" int64_t nfmc_realnvp_blob_floats(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden) {
  return flow_blob_floats(d, n_coupling, n_linear, hidden);
}
extern "C" int nfmc_layout_for_dim(int32_t d, int32_t* lanes_per_chain, int32_t* slots_per_half) {
  Layout L;
  if (!layout_for_dim(d, L)) return set_error("layout_for_dim: d out of range [1, 1024]");
  if (lanes_per_chain) *lanes_per_chain = L.gs;
  if (slots_per_half) *slots_per_half = L.E;
  return 0;
}


extern "C" int nfmc_neutra_hmc_steps(const nfmc_potential* pot, const nfmc_realnvp* flow, float* z, int64_t n, int32_t n_steps,
                                     float step_size, int32_t n_leapfrog, const float* inv_mass_diag, const nfmc_rng* rng,
                                     int64_t chain0, const nfmc_stats* stats, const nfmc_sink* sink, void* stream) {
  if (int e = validate_pot(pot)) return e;
  if (int e = validate_flow(flow)) return e;
  if (pot->d != flow->d) return set_error("neutra_hmc: potential and flow event sizes differ");
  if (!z || n < 1 || n_steps < 0 || n_leapfrog < 0) return set_error("neutra_hmc: bad arguments");
  if (n_steps == 0) return 0;
  Layout L;
  if (!layout_for_dim(pot->d, L)) return set_error("neutra_hmc: unsupported event size");
  NeutraArgs A;
  A.c.pot = pot_params(pot);
  A.c.x = z; A.c.n = n; A.c.chain0 = chain0; A.c.d = pot->d; A.c.gs = L.gs; A.c.n_steps = n_steps;
  A.c.rng = RngArgs{rng ? rng->seed : 0, rng ? rng->step0 : 0, rng ? rng->normals : nullptr, rng ? rng->uniforms : nullptr};
  A.c.stats = StatsArgs{stats ? stats->sum_x : nullptr, stats ? stats->sum_x2 : nullptr, stats ? stats->counts : nullptr};
  A.c.sink = SinkArgs{sink ? sink->samples : nullptr, sink ? sink->seen0 : 0, (sink && sink->thinning > 0) ? sink->thinning : 1};
  A.tau = step_size; A.imd = inv_mass_diag; A.n_leapfrog = n_leapfrog; A.adjusted = 1;
  size_t smem = plan_flow_smem(A.f, flow, L, true, true) + (size_t)((pot->d + 3) & ~3) * sizeof(float);   // + inverse-mass table
  // conditioner stash: the inverse pass keeps every coupling's conditioner outputs for the backward sweep (small path
  // only), if two CTAs of that size still fit an SM
  const size_t stash_b = (size_t)flow->n_coupling * (2 * L.E + kSmallH + 2) * kThreads * sizeof(float);
  A.stash = (flow_is_small(flow->n_linear, flow->hidden) && NFMC_NEUTRA_CTAS * (smem + stash_b + 1024) <= 227 * 1024) ? 1 : 0;
  if (A.stash) smem += stash_b;
  const int grid = grid_for(n, L.gs, NFMC_NEUTRA_CTAS);
  cudaStream_t s = (cudaStream_t)stream;
  A.pot_kind = pot->kind;
  NFMC_DISPATCH_E(L.E, { return launch_neutra_hmc<E>(A, grid, smem, s); });
  return 0;
}

extern "C" int nfmc_neutra_mh_steps(const nfmc_potential* pot, const nfmc_realnvp* flow, float* z, int64_t n, int32_t n_steps,
                                    const float* inv_mass_diag, int32_t adjusted, const nfmc_rng* rng, int64_t chain0,
                                    const nfmc_stats* stats, const nfmc_sink* sink, void* stream) {
  if (int e = validate_pot(pot)) return e;
  if (int e = validate_flow(flow)) return e;
  if (pot->d != flow->d) return set_error("neutra_mh: potential and flow event sizes differ");
  if (!z || n < 1 || n_steps < 0) return set_error("neutra_mh: bad z/n/n_steps");
  if (n_steps == 0) return 0;
  Layout L;
  if (!layout_for_dim(pot->d, L)) return set_error("neutra_mh: unsupported event size");
  NeutraArgs A;
  A.c.pot = pot_params(pot);
  A.c.x = z; A.c.n = n; A.c.chain0 = chain0; A.c.d = pot->d; A.c.gs = L.gs; A.c.n_steps = n_steps;
  A.c.rng = RngArgs{rng ? rng->seed : 0, rng ? rng->step0 : 0, rng ? rng->normals : nullptr, rng ? rng->uniforms : nullptr};
  A.c.stats = StatsArgs{stats ? stats->sum_x : nullptr, stats ? stats->sum_x2 : nullptr, stats ? stats->counts : nullptr};
  A.c.sink = SinkArgs{sink ? sink->samples : nullptr, sink ? sink->seen0 : 0, (sink && sink->thinning > 0) ? sink->thinning : 1};
  A.tau = 0.f; A.imd = inv_mass_diag; A.n_leapfrog = 0; A.stash = 0; A.adjusted = adjusted;
  A.pot_kind = pot->kind;
  const size_t smem = plan_flow_smem(A.f, flow, L, true, true) + (size_t)((pot->d + 3) & ~3) * sizeof(float);
  const int grid = grid_for(n, L.gs, 3);
  cudaStream_t s = (cudaStream_t)stream;
  NFMC_DISPATCH_E(L.E, { return launch_neutra_mh<E>(A, grid, smem, s); });
  return 0;
}

extern "C" int nfmc_tess_steps(const nfmc_potential* pot, const nfmc_realnvp* flow, float* u, int64_t n, int32_t n_steps,
                               int32_t max_iterations, const nfmc_rng* rng, int64_t chain0, const nfmc_stats* stats,
                               const nfmc_sink* sink, void* stream) {
  if (int e = validate_pot(pot)) return e;
  if (int e = validate_flow(flow)) return e;
  if (pot->d != flow->d) return set_error("tess: potential and flow event sizes differ");
  if (!u || n < 1 || n_steps < 0 || max_iterations < 0) return set_error("tess: bad u/n/n_steps/max_iterations");
  if (rng && ((rng->normals == nullptr) != (rng->uniforms == nullptr)))
    return set_error("tess: inject both normals [steps,n,d] and uniforms [steps,n,2+max_iterations], or neither");
  if (n_steps == 0) return 0;
  Layout L;
  if (!layout_for_dim(pot->d, L)) return set_error("tess: unsupported event size");
  TessArgs A;
  A.c.pot = pot_params(pot);
  A.c.x = u; A.c.n = n; A.c.chain0 = chain0; A.c.d = pot->d; A.c.gs = L.gs; A.c.n_steps = n_steps;
  A.c.rng = RngArgs{rng ? rng->seed : 0, rng ? rng->step0 : 0, rng ? rng->normals : nullptr, rng ? rng->uniforms : nullptr};
  A.c.stats = StatsArgs{stats ? stats->sum_x : nullptr, stats ? stats->sum_x2 : nullptr, stats ? stats->counts : nullptr};
  A.c.sink = SinkArgs{sink ? sink->samples : nullptr, sink ? sink->seen0 : 0, (sink && sink->thinning > 0) ? sink->thinning : 1};
  A.pot_kind = pot->kind;
  A.max_iterations = max_iterations;
  const size_t smem = plan_flow_smem(A.f, flow, L, true, true) + (size_t)4 * L.E * kThreads * sizeof(float);   // + step outcome (x, u)
  const int grid = grid_for(n, L.gs, 2);
  cudaStream_t s = (cudaStream_t)stream;
  NFMC_DISPATCH_E(L.E, { return launch_tess<E>(A, grid, smem, s); });
  return 0;
}

extern "C" int nfmc_neutra_potential(const nfmc_potential* pot, const nfmc_realnvp* flow, const float* z, float* u, float* grad,
                                     int64_t n, void* stream) {
  if (int e = validate_pot(pot)) return e;
  if (int e = validate_flow(flow)) return e;
  if (pot->d != flow->d) return set_error("neutra_potential: potential and flow event sizes differ");
  if (!z || !u || n < 1) return set_error("neutra_potential: bad arguments");
  Layout L;
  if (!layout_for_dim(pot->d, L)) return set_error("neutra_potential: unsupported event size");
  FlowArgs FA;
  const size_t smem = plan_flow_smem(FA, flow, L, false);
  const PotParams P = pot_params(pot);
  const int grid = grid_for(n, L.gs, 2);
  cudaStream_t s = (cudaStream_t)stream;
  NFMC_DISPATCH_E(L.E, { return launch_neutra_potential<E>(FA, pot->kind, P, z, u, grad, n, grid, smem, s); });
  return 0;
}

extern "C" int nfmc_neutra_pullback(const nfmc_realnvp* flow, const float* z, const float* grad_x, float* grad_z, float* log_det,
                                    int64_t n, void* stream) {
  if (int e = validate_flow(flow)) return e;
  if (!z || !grad_x || !grad_z || n < 1) return set_error("neutra_pullback: bad arguments");
  Layout L;
  if (!layout_for_dim(flow->d, L)) return set_error("neutra_pullback: unsupported event size");
  FlowArgs FA;
  const size_t smem = plan_flow_smem(FA, flow, L, false);
  const int grid = grid_for(n, L.gs, 2);
  cudaStream_t s = (cudaStream_t)stream;
  NFMC_DISPATCH_E(L.E, { return launch_neutra_pullback<E>(FA, z, grad_x, grad_z, log_det, n, grid, smem, s); });
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// NF jump / one IMH iteration with the flow passes on the tensor cores (wide conditioners).
// workspace (device): z [n,d] | x' [n,d] | ld_inv [n] | log q(x) [n] | uniforms [n]
// ---------------------------------------------------------------------------------------------------------
extern "C" int64_t nfmc_jump_tc_workspace_bytes(int32_t d, int64_t n) {
  const size_t row = ((size_t)n * d * sizeof(float) + 255) & ~size_t(255), vec = ((size_t)n * sizeof(float) + 255) & ~size_t(255);
  return (int64_t)(2 * row + 3 * vec);
}

// tc_jump.cu: the whole jump as one kernel; -1 = not eligible (shape / alignment / shared memory)
int nfmc_jump_step_tc_fused(const nfmc_potential* pot, const nfmc_realnvp_tc* flow, float* x, float* logq_cache, int recompute_logq,
                            int64_t n, int adjusted, const nfmc_rng* rng, int64_t chain0, const nfmc_stats* stats,
                            const nfmc_sink* sink, void* stream);

// The jump composed from separate launches: log q(x) by `logprob`, the Philox base draw (stream 1, the numbers the fused
// kernels draw), x' = T^-1(z) by `inverse`, then the accept kernel.  Shared by the tensor-core fallback and the row-tile fp32
// path for deep / odd-sized conditioners.
template <class LogProb, class Inverse>
static int composed_jump(const char* who, const nfmc_potential* pot, float* x, float* logq_cache, int32_t recompute_logq, int64_t n,
                         int32_t adjusted, const nfmc_rng* rng, int64_t chain0, const nfmc_stats* stats, const nfmc_sink* sink,
                         void* workspace, int64_t workspace_bytes, void* stream, LogProb logprob, Inverse inverse) {
  if (workspace_bytes < nfmc_jump_tc_workspace_bytes(pot->d, n)) return set_error(std::string(who) + ": workspace too small");
  const int d = pot->d;
  const size_t row = ((size_t)n * d * sizeof(float) + 255) & ~size_t(255), vec = ((size_t)n * sizeof(float) + 255) & ~size_t(255);
  unsigned char* w = static_cast<unsigned char*>(workspace);
  float* z = reinterpret_cast<float*>(w);
  float* xp = reinterpret_cast<float*>(w + row);
  float* ld_inv = reinterpret_cast<float*>(w + 2 * row);
  float* logq_x = reinterpret_cast<float*>(w + 2 * row + vec);
  float* unif = reinterpret_cast<float*>(w + 2 * row + 2 * vec);
  cudaStream_t s = (cudaStream_t)stream;
  // log q(x): forward pass (jump.py:218) unless a valid cache is supplied (imh.py:214)
  const float* fx = logq_x;
  if (adjusted) {
    if (logq_cache && !recompute_logq) fx = logq_cache;
    else if (int e = logprob(x, logq_x)) return e;
  }
  // base draw z and accept uniforms (Philox stream 1, or injected)
  const float* zsrc = z;
  const float* usrc = unif;
  if (rng->normals && rng->uniforms) { zsrc = rng->normals; usrc = rng->uniforms; }
  else {
    nfmc_rng r2{rng->seed, rng->step0, nullptr, nullptr};
    if (int e = nfmc_rng_fill(&r2, 1, chain0, d, n, 1, z, unif, stream)) return e;
    if (rng->normals) zsrc = rng->normals;
    if (rng->uniforms) usrc = rng->uniforms;
  }
  // x' = T^-1(z), log|det| (jump.py:205)
  if (int e = inverse(zsrc, xp, ld_inv)) return e;
  Layout L;
  if (!layout_for_dim(d, L)) return set_error(std::string(who) + ": unsupported event size");
  AcceptArgs A;
  A.c.pot = pot_params(pot);
  A.c.x = x; A.c.n = n; A.c.chain0 = chain0; A.c.d = d; A.c.gs = L.gs; A.c.n_steps = 1;
  A.c.rng = RngArgs{rng->seed, rng->step0, nullptr, nullptr};
  A.c.stats = StatsArgs{stats ? stats->sum_x : nullptr, stats ? stats->sum_x2 : nullptr, stats ? stats->counts : nullptr};
  A.c.sink = SinkArgs{sink ? sink->samples : nullptr, sink ? sink->seen0 : 0, (sink && sink->thinning > 0) ? sink->thinning : 1};
  A.pot_kind = pot->kind; A.adjusted = adjusted;
  A.x_prime = xp; A.z = zsrc; A.ld_inv = ld_inv; A.logq_x = fx; A.uniforms = usrc; A.logq_cache = logq_cache;
  const size_t smem = (cta_stats_bytes_host(d) + 15) & ~size_t(15);
  const int grid = grid_for(n, L.gs, 4);
  NFMC_DISPATCH_E(L.E, { return launch_jump_accept<E>(A, grid, smem, s); });
  return 0;
}

extern "C" int nfmc_jump_step_tc(const nfmc_potential* pot, const nfmc_realnvp_tc* flow, float* x, float* logq_cache,
                                 int32_t recompute_logq, int64_t n, int32_t adjusted, const nfmc_rng* rng, int64_t chain0,
                                 const nfmc_stats* stats, const nfmc_sink* sink, void* workspace, int64_t workspace_bytes,
                                 void* stream) {
  if (int e = validate_pot(pot)) return e;
  if (!flow || !x || !rng || !workspace || n < 1) return set_error("jump_step_tc: bad arguments");
  if (pot->d != flow->d) return set_error("jump_step_tc: potential and flow event sizes differ");
  if (int e = validate_injected(rng, adjusted, "jump_step_tc")) return e;
  const bool no_fused = getenv("NFMC_TC_NO_FUSED_JUMP") != nullptr;             // A-B tests: compose the jump from separate launches
  if (!no_fused) {
    const int rc = nfmc_jump_step_tc_fused(pot, flow, x, logq_cache, recompute_logq, n, adjusted, rng, chain0, stats, sink, stream);
    if (rc >= 0) return rc;
  }
  return composed_jump("jump_step_tc", pot, x, logq_cache, recompute_logq, n, adjusted, rng, chain0, stats, sink, workspace,
                       workspace_bytes, stream,
                       [&](const float* in, float* lq) { return nfmc_flow_tc_pass(flow, 2, in, nullptr, lq, n, stream); },
                       [&](const float* zz, float* out, float* ld) { return nfmc_flow_tc_pass(flow, 1, zz, out, ld, n, stream); });
}

// The same step for conditioner shapes outside the register-resident and the tensor-core paths (n_linear != 2, odd d, d > 128
// with hidden > 8): both flow passes by the row-tile fp32 kernel of train_wide.cu, straight from the module-order parameter
// vector (`flow->blob` = theta, `flow->blob_floats` = its length) -- 3-6x the generic per-chain conditioner of flow.cuh.
extern "C" int nfmc_jump_step_wide(const nfmc_potential* pot, const nfmc_realnvp* flow, int32_t theta_transposed, float* x, float* logq_cache,
                                   int32_t recompute_logq, int64_t n, int32_t adjusted, const nfmc_rng* rng, int64_t chain0,
                                   const nfmc_stats* stats, const nfmc_sink* sink, void* workspace, int64_t workspace_bytes,
                                   void* stream) {
  if (int e = validate_pot(pot)) return e;
  if (!flow || !flow->blob || !x || !rng || !workspace || n < 1) return set_error("jump_step_wide: bad arguments");
  if (pot->d != flow->d) return set_error("jump_step_wide: potential and flow event sizes differ");
  if (flow->blob_floats != nfmc_flow_wide_param_count(flow->d, flow->n_coupling, flow->n_linear, flow->hidden))
    return set_error("jump_step_wide: theta length does not match the flow shape");
  if (int e = validate_injected(rng, adjusted, "jump_step_wide")) return e;
  const int d = flow->d, Lc = flow->n_coupling, M = flow->n_linear, H = flow->hidden;
  const float* theta = flow->blob;
  return composed_jump("jump_step_wide", pot, x, logq_cache, recompute_logq, n, adjusted, rng, chain0, stats, sink, workspace,
                       workspace_bytes, stream,
                       [&](const float* in, float* lq) { return nfmc_flow_wide_log_prob(d, Lc, M, H, theta, theta_transposed, in, lq, n, stream); },
                       [&](const float* zz, float* out, float* ld) {
                         return nfmc_flow_wide_pass(d, Lc, M, H, theta, 1 | (theta_transposed ? 2 : 0), zz, out, ld, n, stream);
                       });
}
