/*
 * nfmc_b200.h -- C ABI of libnfmc_b200.so: the B200 (sm_100a) implementation of nfmc's batched
 * per-iteration hot path (jump_mala / jump_hmc / neutra_hmc / imh + RealNVP).
 *
 * The reference (davidnabergoj/nfmc) is pure Python and has no FFI; the seams this library sits behind are
 * the duck-typed operator contracts listed next to each entry point (paths under /root/reference/nfmc/).
 * A maintainer binds these with ctypes; nfmc_b200/_native.py is that binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; tensors are fp32, row-major
 *     [n, d] with d = prod(event_shape); no ownership is transferred; nothing is allocated by the library
 *     except inside the *_host entry points.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); work is enqueued, never
 *     synchronised, except in *_host entry points which synchronise before returning.
 *   - return value: 0 on success, nonzero on error; nfmc_last_error() gives the message (thread-local).
 *   - chains are keyed by GLOBAL index `chain0 + row`, so a chain's random stream does not depend on how
 *     the batch is sharded over GPUs.
 */
#ifndef NFMC_B200_H
#define NFMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NFMC_ABI_VERSION 1
#if defined(__GNUC__)
#define NFMC_API __attribute__((visibility("default")))
#else
#define NFMC_API
#endif
#define NFMC_MAX_DIM 1024

/* ---- analytic target potentials (negative log densities) -------------------------------------------- */
enum nfmc_potential_kind {
  NFMC_POT_ISO_GAUSSIAN = 0, /* U = 1/2 w sum x_i^2;  scalar[0] = w  (w = 2: README.md:45-46 `sum(x**2)`)  */
  NFMC_POT_DIAG_GAUSSIAN = 1,/* U = 1/2 sum w_i (x_i - mu_i)^2;  params = float2[d] {w_i, mu_i}            */
  NFMC_POT_FUNNEL = 2,       /* U = x0^2/(2 s^2) + (d-1)/2 x0 + 1/2 exp(-x0) sum_{i>=1} x_i^2; scalar[0]=s (3) */
  NFMC_POT_ROSENBROCK = 3,   /* U = sum_{k<d/2} (x_k - 1)^2 + c (x_{k+d/2} - x_k^2)^2;  scalar[0] = c (10)  */
  NFMC_POT_MIXTURE4 = 4      /* U = -logsumexp_k(-1/2 |x - mu_k|^2), mu_k = (+-a, +-a, 0, ...); scalar[0]=a */
};

typedef struct nfmc_potential {
  int32_t kind;        /* nfmc_potential_kind */
  int32_t d;
  const float* params; /* device, per-dimension parameters or NULL */
  float scalar[4];
} nfmc_potential;

/* ---- packed RealNVP (replaces torchflows.RealNVP; call sites sampling/base.py:26, util.py:280-281) ---- */
/* blob layout: see DESIGN.md "flow blob"; produced by nfmc_b200.flow.pack_realnvp() or nfmc_realnvp_blob_floats */
typedef struct nfmc_realnvp {
  int32_t d;          /* event size */
  int32_t n_coupling; /* Lc */
  int32_t n_linear;   /* M: linear layers per conditioner (>= 1) */
  int32_t hidden;     /* H */
  const float* blob;  /* device */
  int64_t blob_floats;
} nfmc_realnvp;

/* RealNVP packed for the tensor-core path (cond_tc.cu): bf16 conditioner weights in the UMMA shared-memory image,
 * fp32 affine tables and biases; M = 2 linear layers, hidden a multiple of 16 in [16, 256], even d <= 128.
 * Produced by nfmc_b200.flow.pack_realnvp_tc(). */
typedef struct nfmc_realnvp_tc {
  int32_t d;
  int32_t n_coupling;
  int32_t hidden;
  int32_t reserved;
  const void* blob;   /* device */
  int64_t blob_bytes;
} nfmc_realnvp_tc;

/* ---- random numbers: counter-based Philox4x32-10, or injected tensors ----------------------------------- */
typedef struct nfmc_rng {
  uint64_t seed;
  uint64_t step0;        /* global index of the first step of this launch (advance by the steps taken) */
  const float* normals;  /* optional injected N(0,1): [steps, n, d]  (NULL -> Philox) */
  const float* uniforms; /* optional injected U[0,1): [steps, n] (ESS: [steps, n, 2+M]) (NULL -> Philox) */
} nfmc_rng;

/* ---- statistics accumulated on the device (replaces MCMCExpectation.update, sampling/base.py:75-95, and
 *      MCMCStatistics.update_counters, sampling/base.py:139-149) ------------------------------------------ */
typedef struct nfmc_stats {
  double* sum_x;              /* [d]  += sum over (step, chain) of x after the accept          (or NULL) */
  double* sum_x2;             /* [d]  += sum of x^2                                              (or NULL) */
  unsigned long long* counts; /* [4]: [0] += accepted, [1] += attempted, [2] += non-finite log-ratios, [3] ESS only */
} nfmc_stats;

/* optional sample sink (MCMCSamples.add, sampling/base.py:234-263): rows written for steps whose global
 * index (seen0 + k) is divisible by `thinning`; row r of the launch goes to samples[r, :, :] */
typedef struct nfmc_sink {
  float* samples;   /* [rows, n, d] or NULL */
  int64_t seen0;
  int32_t thinning; /* >= 1 */
} nfmc_sink;

NFMC_API const char* nfmc_last_error(void);
NFMC_API int nfmc_abi_version(void);

/* lanes-per-chain / slots-per-lane layout chosen for event size d (defines the Philox noise layout) */
NFMC_API int nfmc_layout_for_dim(int32_t d, int32_t* lanes_per_chain, int32_t* slots_per_half);
/* number of floats of the packed blob for (d, Lc, M, H) */
NFMC_API int64_t nfmc_realnvp_blob_floats(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden);

/* U(x) [n] and optionally grad U(x) [n,d]  -- replaces `target(x)` + autograd (langevin.py:66-68, hmc.py:40-48) */
NFMC_API int nfmc_potential_eval(const nfmc_potential* pot, const float* x, float* u, float* grad, int64_t n, void* stream);

/* RealNVP bijection.forward / bijection.inverse (neutra.py:60,122): y [n,d], log_det [n] */
NFMC_API int nfmc_realnvp_forward(const nfmc_realnvp* flow, const float* x, float* z, float* log_det, int64_t n, void* stream);
NFMC_API int nfmc_realnvp_inverse(const nfmc_realnvp* flow, const float* z, float* x, float* log_det, int64_t n, void* stream);
/* Flow.log_prob (jump.py:218, imh.py:133-134,214) */
NFMC_API int nfmc_flow_log_prob(const nfmc_realnvp* flow, const float* x, float* log_q, int64_t n, void* stream);
/* Flow.sample(n, return_log_prob=True) (jump.py:205, imh.py:221): base draw from rng (stream id 1) */
NFMC_API int nfmc_flow_sample(const nfmc_realnvp* flow, const nfmc_rng* rng, int64_t chain0, float* x, float* log_q,
                     int64_t n, void* stream);

/* the same three operators with the conditioner MLP on the tcgen05 tensor cores (bf16 operands, fp32 accumulate);
 * mode 0 = forward (out = z, aux = log_det), 1 = inverse (out = x, aux = log_det), 2 = log_prob (aux = log q, out unused) */
NFMC_API int64_t nfmc_realnvp_tc_blob_bytes(int32_t d, int32_t n_coupling, int32_t hidden);
NFMC_API int nfmc_flow_tc_pass(const nfmc_realnvp_tc* flow, int32_t mode, const float* in, float* out, float* aux,
                               int64_t n, void* stream);

/* one NF jump (or one IMH iteration) for a wide flow: log q(x) and x' = T^-1(z) on the tensor cores, then the accept
 * step of jump.py:212-231 / imh.py:223-233.  logq_cache [n] (optional) is the IMH cache of log q(x): read unless
 * recompute_logq, and updated where accepted.  workspace: nfmc_jump_tc_workspace_bytes(d, n) bytes of device memory. */
NFMC_API int64_t nfmc_jump_tc_workspace_bytes(int32_t d, int64_t n);
NFMC_API int nfmc_jump_step_tc(const nfmc_potential* pot, const nfmc_realnvp_tc* flow, float* x, float* logq_cache,
                               int32_t recompute_logq, int64_t n, int32_t adjusted, const nfmc_rng* rng, int64_t chain0,
                               const nfmc_stats* stats, const nfmc_sink* sink, void* workspace, int64_t workspace_bytes,
                               void* stream);

/* the same step (jump.py:203-243, imh.py:214-249) for conditioner shapes outside the register-resident and tensor-core
 * paths: flow->blob = the module-order parameter vector theta of nfmc_flow_wide_param_count() floats (see the wide
 * training block below; theta_transposed: its linear weights are stored [in][out]), both passes by the row-tile fp32 kernel;
 * workspace as for nfmc_jump_step_tc */
NFMC_API int nfmc_jump_step_wide(const nfmc_potential* pot, const nfmc_realnvp* flow, int32_t theta_transposed, float* x, float* logq_cache,
                                 int32_t recompute_logq, int64_t n, int32_t adjusted, const nfmc_rng* rng, int64_t chain0,
                                 const nfmc_stats* stats, const nfmc_sink* sink, void* workspace, int64_t workspace_bytes,
                                 void* stream);

/* ---- NeuTra on the tensor cores (csrc/tc_neutra.cu): U~(z) = U(T^-1 z) - log|det dT^-1/dz| (nfmc/neutra.py:58-68) and
 * its gradient with the conditioner forward AND its input-VJP (dgrad) on tcgen05.  blob_t = the transposed weight images
 * of nfmc_b200.flow.pack_realnvp_tc_transposed() (nfmc_neutra_tc_transposed_bytes bytes; -1 = shape not eligible: needs
 * hidden % 32 == 0, d % 4 == 0 and a shared-memory plan within 227 KB). */
NFMC_API int64_t nfmc_neutra_tc_transposed_bytes(int32_t d, int32_t n_coupling, int32_t hidden);
NFMC_API int64_t nfmc_neutra_tc_workspace_bytes(int32_t d, int64_t n);
/* value u [n] and gradient grad [n, d] of U~ at z [n, d]; x [n, d] and ld [n] receive T^-1(z) and log|det dx/dz| */
NFMC_API int nfmc_neutra_potential_tc(const nfmc_potential* pot, const nfmc_realnvp_tc* flow, const void* blob_t, int64_t blob_t_bytes,
                             const float* z, float* x, float* ld, float* u, float* grad, int64_t n, void* stream);
/* T NeuTra-HMC iterations (HMC.propose, mcmc/hmc.py:96-126, on the latent potential; same contract as
 * nfmc_neutra_hmc_steps): per step the momentum draw, L + 1 tensor-core gradient evaluations with the leapfrog updates
 * of hmc.py:51-58 between them, the accept test of hmc.py:103-113, moments / counters / sink of the latent state */
NFMC_API int nfmc_neutra_hmc_steps_tc(const nfmc_potential* pot, const nfmc_realnvp_tc* flow, const void* blob_t, int64_t blob_t_bytes,
                             float* z, int64_t n, int32_t n_steps, float step_size, int32_t n_leapfrog,
                             const float* inv_mass_diag, int32_t adjusted, const nfmc_rng* rng, int64_t chain0,
                             const nfmc_stats* stats, const nfmc_sink* sink, void* workspace, int64_t workspace_bytes,
                             void* stream);

/* K Langevin steps for all chains -- Langevin.propose (mcmc/langevin.py:61-122) inside the local loop
 * MCMCSampler.sample (mcmc/base.py:69-99).  inv_mass_diag may be NULL (= ones).  adjusted=0 -> ULA. */
NFMC_API int nfmc_mala_steps(const nfmc_potential* pot, float* x, int64_t n, int32_t n_steps, float step_size,
                    const float* inv_mass_diag, int32_t adjusted, const nfmc_rng* rng, int64_t chain0,
                    const nfmc_stats* stats, const nfmc_sink* sink, void* stream);

/* K random-walk Metropolis steps -- MH.propose (mcmc/mh.py:44-73): x' = x + inv_mass_diag * xi (NULL = ones),
 * accept iff log u < U(x) - U(x'); adjusted=0 -> RandomWalk (always accept) */
NFMC_API int nfmc_mh_steps(const nfmc_potential* pot, float* x, int64_t n, int32_t n_steps, const float* inv_mass_diag,
                  int32_t adjusted, const nfmc_rng* rng, int64_t chain0, const nfmc_stats* stats, const nfmc_sink* sink,
                  void* stream);

/* K elliptical-slice steps with prior N(0, I) -- elliptical_slice_sampling_step (mcmc/ess.py:12-64) wrapped as
 * ESS.propose (mcmc/ess.py:97-116): `nll` is the negative log-likelihood, at most `max_iterations` bracket rounds per
 * step, every chain counts as accepted (ess.py:107); counts[3] += chain-steps whose bracket produced a point.
 * Injected noise (both or neither): rng->normals [steps, n, d] = nu, rng->uniforms [steps, n, 2 + max_iterations] =
 * {u (:35), theta0 (:39), bracket draws (:58)}. */
NFMC_API int nfmc_ess_steps(const nfmc_potential* nll, float* x, int64_t n, int32_t n_steps, int32_t max_iterations,
                   const nfmc_rng* rng, int64_t chain0, const nfmc_stats* stats, const nfmc_sink* sink, void* stream);

/* K HMC steps -- HMC.propose (mcmc/hmc.py:96-126; trajectory :61-77) inside the same local loop */
NFMC_API int nfmc_hmc_steps(const nfmc_potential* pot, float* x, int64_t n, int32_t n_steps, float step_size,
                   int32_t n_leapfrog, const float* inv_mass_diag, int32_t adjusted, const nfmc_rng* rng,
                   int64_t chain0, const nfmc_stats* stats, const nfmc_sink* sink, void* stream);

/* one NF jump for all chains -- JumpNFMC.sample jump block (nfmc/jump.py:203-243); counts[0]/[1] receive
 * accepted / attempted JUMPS */
NFMC_API int nfmc_jump_step(const nfmc_potential* pot, const nfmc_realnvp* flow, float* x, int64_t n, int32_t adjusted,
                   const nfmc_rng* rng, int64_t chain0, const nfmc_stats* stats, const nfmc_sink* sink, void* stream);

/* the same NF jump as two kernels -- log q(x) by a forward pass into logq_scratch [n] (device), then proposal + accept with
 * x loaded after the inverse pass: no register spills, 2.0 ms instead of 3.1 ms per 2^20 jumps at d = 100; same results */
NFMC_API int nfmc_jump_step2(const nfmc_potential* pot, const nfmc_realnvp* flow, float* x, float* logq_scratch, int64_t n,
                    int32_t adjusted, const nfmc_rng* rng, int64_t chain0, const nfmc_stats* stats, const nfmc_sink* sink,
                    void* stream);

/* T independence-MH iterations -- FixedIMH.sample (nfmc/imh.py:214-249).  log_q_x [n] is the cached
 * flow.log_prob(x) (imh.py:214), updated in place; pass recompute_logq=1 for AdaptiveIMH (imh.py:133-134) */
NFMC_API int nfmc_imh_steps(const nfmc_potential* pot, const nfmc_realnvp* flow, float* x, float* log_q_x, int64_t n,
                   int32_t n_steps, int32_t recompute_logq, const nfmc_rng* rng, int64_t chain0,
                   const nfmc_stats* stats, const nfmc_sink* sink, void* stream);

/* T NeuTra-HMC iterations in latent space -- NeuTra.adjusted_target (nfmc/neutra.py:58-68) under HMC.propose */
NFMC_API int nfmc_neutra_hmc_steps(const nfmc_potential* pot, const nfmc_realnvp* flow, float* z, int64_t n, int32_t n_steps,
                          float step_size, int32_t n_leapfrog, const float* inv_mass_diag, const nfmc_rng* rng,
                          int64_t chain0, const nfmc_stats* stats, const nfmc_sink* sink, void* stream);
/* T transport-elliptical-slice steps -- transport_elliptical_slice_sampling_step (nfmc/tess.py:15-75, identity
 * covariance) inside TESS.sample (nfmc/tess.py:151-188).  u [n, d] is the latent chain state (updated in place); the
 * statistics and the sample sink receive the data-space points x = T^-1(u) of every step; counts[0] += chains whose bracket
 * produced a point, counts[1] += n per step.  Injected noise (both or neither): rng->normals [steps, n, d] = v,
 * rng->uniforms [steps, n, 2 + max_iterations] = {w (:41), theta_n -- a NORMAL draw (:45) --, bracket uniforms (:69)}. */
NFMC_API int nfmc_tess_steps(const nfmc_potential* pot, const nfmc_realnvp* flow, float* u, int64_t n, int32_t n_steps,
                    int32_t max_iterations, const nfmc_rng* rng, int64_t chain0, const nfmc_stats* stats,
                    const nfmc_sink* sink, void* stream);

/* T random-walk Metropolis steps in the latent space of the flow -- NeuTraMH (nfmc/neutra.py:147-159) = MH.propose
 * (mcmc/mh.py:44-73) on NeuTra.adjusted_target: z' = z + inv_mass_diag * xi (NULL = ones), accept iff
 * log u < U~(z) - U~(z'); adjusted = 0 -> plain random walk */
NFMC_API int nfmc_neutra_mh_steps(const nfmc_potential* pot, const nfmc_realnvp* flow, float* z, int64_t n, int32_t n_steps,
                         const float* inv_mass_diag, int32_t adjusted, const nfmc_rng* rng, int64_t chain0,
                         const nfmc_stats* stats, const nfmc_sink* sink, void* stream);

/* latent potential and its gradient (for tests): u [n], grad [n,d] (grad may be NULL) */
NFMC_API int nfmc_neutra_potential(const nfmc_potential* pot, const nfmc_realnvp* flow, const float* z, float* u, float* grad,
                          int64_t n, void* stream);

/* Warm-up adaptation on the device -- replaces MetropolisSampler.update_kernel (/root/reference/nfmc/algorithms/sampling/
 * mcmc/base.py:142-161).  nfmc_chain_sums: sums[0:d] = sum over chains of x, sums[d:2d] = sum of x^2 (fp64), sums[2d] =
 * counts[0] (accepted so far; 0 when counts is NULL), sums[2d+1] = n.  A multi-GPU caller all-reduces the 2d+2 doubles,
 * then nfmc_tune_inv_mass applies imd <- c * var + (1 - c) * imd with the unbiased variance of the pooled sums
 * (no change when fewer than two chains, base.py:150). */
NFMC_API int nfmc_chain_sums(const float* x, int64_t n, int32_t d, double* sums, const uint64_t* counts, void* stream);
NFMC_API int nfmc_tune_inv_mass(const double* sums, int32_t d, float imd_adjustment, float* inv_mass_diag, void* stream);

/* the numbers the Philox path would draw (for parity tests): normals [steps, n, d], uniforms [steps, n];
 * stream_id 0 = local kernels, 1 = flow base draw */
NFMC_API int nfmc_rng_fill(const nfmc_rng* rng, int32_t stream_id, int64_t chain0, int32_t d, int64_t n, int32_t n_steps,
                  float* normals, float* uniforms, void* stream);

/* ---- external targets: the reference accepts ANY callable `target([n,*event]) -> [n]` and differentiates it with autograd
 * (sample.py:34-36,65-66; mcmc/langevin.py:66-68; mcmc/hmc.py:40-48).  A callable cannot be fused into the step kernels, so for
 * such targets the host evaluates U and grad U on the device (torch autograd) and these entry points do everything else:
 * proposal, proposal potentials, leapfrog updates, Hamiltonians, log-ratio, accept, masked overwrite, moments, counters,
 * sample sink.  All pointers are device memory, row-major fp32; noise / uniforms are explicit (nfmc_rng_fill draws the
 * Philox numbers the fused kernels would use). ---------------------------------------------------------------------- */
/* x' = x - tau/m^2 grad + sqrt(2 tau)/m noise (langevin.py:74-76); random_walk=1: x' = x + m noise (mh.py:52-56, grad unused) */
NFMC_API int nfmc_ext_langevin_propose(const float* x, const float* grad, const float* noise, const float* inv_mass_diag,
                              float step_size, int32_t random_walk, int64_t n, int32_t d, float* x_prime, void* stream);
/* log_ratio [n] = (-U') - (-U) + (-q(x|x')) - (-q(x'|x)) with both proposal potentials as langevin.py:31-42 (util.py:392);
 * random_walk=1: (-U') - (-U) (mh.py:59) */
NFMC_API int nfmc_ext_langevin_log_ratio(const float* x, const float* x_prime, const float* grad, const float* grad_prime,
                                const float* u, const float* u_prime, const float* inv_mass_diag, float step_size,
                                int32_t random_walk, int64_t n, int32_t d, float* log_ratio, void* stream);
/* p = noise / sqrt(m), kinetic [n] = sum p^2 m (hmc.py:100,104) */
NFMC_API int nfmc_ext_hmc_momentum(const float* noise, const float* inv_mass_diag, int64_t n, int32_t d, float* p, float* kinetic,
                          void* stream);
/* `kicks` (0..2) half-kicks p -= tau/2 grad (hmc.py:51-53), then, if drift, x += tau m p (hmc.py:56-58) */
NFMC_API int nfmc_ext_hmc_leapfrog(float* x, float* p, const float* grad, const float* inv_mass_diag, float step_size,
                          int32_t kicks, int32_t drift, int64_t n, int32_t d, void* stream);
/* log_ratio [n] = -(U1 + 1/2 sum p^2 m) + (U0 + 1/2 kinetic0) (hmc.py:103-111) */
NFMC_API int nfmc_ext_hmc_log_ratio(const float* p, const float* inv_mass_diag, const float* u0, const float* kinetic0,
                           const float* u1, int64_t n, int32_t d, float* log_ratio, void* stream);
/* log_ratio [n] = (-U') - (-U) + log q(x) - log q(x') (jump.py:224-229, imh.py:226-231) */
NFMC_API int nfmc_ext_jump_log_ratio(const float* u, const float* u_prime, const float* log_q, const float* log_q_prime, int64_t n,
                            float* log_ratio, void* stream);
/* accept iff log(uniforms) < log_ratio (adjusted=0: always); x[mask] = x'[mask] (mcmc/base.py:77) together with up to three
 * caches that travel with the state (aux_a, aux_b: [n]; aux_grad: [n,d]; NULL pairs are skipped); moments of the post-accept
 * state and counters into `stats` (counts[1] += n); the post-accept state goes to the sink row of step `sink_step` if the
 * thinning rule keeps it */
NFMC_API int nfmc_ext_accept(float* x, const float* x_prime, const float* log_ratio, const float* uniforms, int32_t adjusted,
                    int64_t n, int32_t d, float* aux_a, const float* aux_a_prime, float* aux_b, const float* aux_b_prime,
                    float* aux_grad, const float* aux_grad_prime, const nfmc_stats* stats, const nfmc_sink* sink,
                    int32_t sink_step, void* stream);

/* elliptical slice sampling around an external negative log-likelihood (mcmc/ess.py:12-64): the 2 + M scalar uniforms of a
 * step as the fused kernel draws them (Philox stream 2); threshold / initial angle / bracket -> state [n][4] = {log y, theta,
 * theta_min, theta_max}, found [n] = 0; f' = f cos(theta) + nu sin(theta); one bracket round given nll(f') */
NFMC_API int nfmc_ext_ess_uniforms(uint64_t seed, uint64_t step, int64_t chain0, int64_t n, int32_t n_uniforms, float* out, void* stream);
NFMC_API int nfmc_ext_ess_begin(const float* nll_cur, const float* uniforms, int32_t n_uniforms, int64_t n, float* state, int32_t* found,
                       void* stream);
NFMC_API int nfmc_ext_ess_rotate(const float* f, const float* nu, const float* state, int64_t n, int32_t d, float* f_prime, void* stream);
NFMC_API int nfmc_ext_ess_update(float* f, const float* f_prime, float* nll_cur, const float* nll_prime, float* state, int32_t* found,
                        const float* uniforms, int32_t n_uniforms, int32_t round, int64_t n, int32_t d, void* stream);

/* NeuTra with an external target: grad_z [ U(T^-1 z) - log|det dT^-1/dz| ] (neutra.py:58-68) from z and grad_x = grad U(x) at
 * x = T^-1 z (which nfmc_realnvp_inverse gives): inverse pass + reversible backward sweep seeded with grad_x.  log_det
 * (optional, [n]) receives log|det dT^-1/dz|. */
NFMC_API int nfmc_neutra_pullback(const nfmc_realnvp* flow, const float* z, const float* grad_x, float* grad_z, float* log_det,
                         int64_t n, void* stream);

/* ---- flow training on the device (register-resident conditioner path: n_linear = 2, hidden <= 8) -----------------
 * Replaces the autograd + AdamW loop behind flow.fit (nfmc/jump.py:139-151,201; nfmc/imh.py:171-175) and
 * flow.variational_fit (nfmc/imh.py:67-72; nfmc/neutra.py:84-91); torchflows itself is absent, its optimiser settings
 * are the kwargs the reference passes (lr = 0.05, AdamW).
 * theta = the module parameters in state_dict order:
 *   affine_0.value[d][2] | Lc x { W1[H][d/2] b1[H] Wl[2(d-d/2)][H] bl[2(d-d/2)] actnorm.value[d][2] } |
 *   affine_T.value[d][2] | actnorm_T.value[d][2]                      (nfmc_flow_param_count floats, -1 = unsupported) */
NFMC_API int64_t nfmc_flow_param_count(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden);
/* theta -> kernel blob (same layout nfmc_b200.flow.pack_realnvp produces on the host) */
NFMC_API int nfmc_flow_pack(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden, const float* theta, float* blob,
                   void* stream);
/* maximum likelihood: *loss (+)= -sum_i log q(x[rows[i]]), grad_blob[blob_floats] (+)= its gradient in blob layout;
 * rows may be NULL (= 0..n-1); accumulate = 0 zeroes grad_blob and *loss first */
NFMC_API int nfmc_flow_nll_grad(const nfmc_realnvp* flow, const float* x, const int64_t* rows, int64_t n, float* grad_blob,
                       double* loss, int32_t accumulate, void* stream);
/* reverse KL: z_i ~ N(0, I) (Philox stream 1 at rng->step0, or rng->normals [n, d]), x_i = T^-1(z_i),
 * *loss (+)= sum_i [log q(x_i) + U(x_i)], grad_blob (+)= its gradient (through x_i, grad U included) */
NFMC_API int nfmc_flow_kl_grad(const nfmc_potential* pot, const nfmc_realnvp* flow, const nfmc_rng* rng, int64_t chain0,
                      int64_t n, float* grad_blob, double* loss, int32_t accumulate, void* stream);
/* blob-layout gradient -> gradient with respect to theta (chain rule through the merged affines), times `scale` */
NFMC_API int nfmc_flow_grad_unpack(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden, const float* theta,
                          const float* grad_blob, float scale, float* grad_theta, void* stream);
/* one AdamW update (torch.optim.AdamW semantics); step counts from 1 */
NFMC_API int nfmc_adamw_step(float* theta, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int32_t step, void* stream);
/* one epoch of minibatch maximum likelihood on one GPU: for each batch of perm[n]: pack, loss + gradient, unpack with
 * 1/batch, AdamW (steps step0+1, ...); losses[batch] = summed loss of the batch; blob / grad_blob / grad_theta are
 * scratch of blob_floats / blob_floats / param_count floats */
NFMC_API int nfmc_flow_fit_epoch(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden, float* theta, float* exp_avg,
                        float* exp_avg_sq, float* blob, float* grad_blob, float* grad_theta, double* losses,
                        const float* x, const int64_t* perm, int64_t n, int64_t batch_size, float lr, float beta1,
                        float beta2, float eps, float weight_decay, int32_t step0, void* stream);

/* ---- flow training for wide / deep conditioners (any n_linear in [1, 16], any hidden): csrc/train_wide.cu ------------
 * Same role as the block above (flow.fit: nfmc/jump.py:139-151,201, nfmc/imh.py:171-175; flow.variational_fit:
 * nfmc/imh.py:67-72, nfmc/neutra.py:84-91) for every conditioner shape the register-resident kernels do not cover.
 * Works directly in module order: theta = affine_0.value[d][2] | Lc x { linear_0 {W[out][in], b[out]} ... linear_{M-1}
 * | actnorm.value[d][2] } | affine_T.value[d][2] | actnorm_T.value[d][2]; gradients come back in the same order. */
NFMC_API int64_t nfmc_flow_wide_param_count(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden);
/* maximum likelihood: *loss (+)= -sum_i log q(x[rows[i]]), grad_theta (+)= its gradient; rows may be NULL; grad_x
 * (optional, [n, d]) receives d(-log q)/dx; accumulate = 0 zeroes grad_theta and *loss first */
NFMC_API int nfmc_flow_wide_nll_grad(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden, const float* theta,
                            const float* x, const int64_t* rows, int64_t n, float* grad_theta, double* loss,
                            float* grad_x, int32_t accumulate, void* stream);
/* one fp32 pass straight from theta: inverse bit 0 = 0: x -> z, log|det dz/dx|; 1: z -> x, log|det dx/dz| (log_det may be NULL);
 * inverse bit 1 (value 2): every linear's weight is stored TRANSPOSED in theta ([in][out], same offsets) -- the layout the
 * sampling path packs (nfmc_b200.flow.RealNVP.theta_descriptor), read coalesced by the forward GEMMs */
NFMC_API int nfmc_flow_wide_pass(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden, const float* theta,
                        int32_t inverse, const float* in, float* out, float* log_det, int64_t n, void* stream);
/* log q(x) = log N(T(x); 0, I) + log|det dT/dx| by the same fp32 pass (Flow.log_prob: nfmc/jump.py:218, nfmc/imh.py:214) for
 * the conditioner shapes outside the register-resident and tensor-core paths (n_linear != 2, odd d, d > 128 with H > 8) */
NFMC_API int nfmc_flow_wide_log_prob(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden, const float* theta,
                            int32_t transposed, const float* x, float* log_q, int64_t n, void* stream);
/* Flow.sample(n, return_log_prob=True) (jump.py:205, imh.py:221) for those shapes: base draw z from rng (Philox stream 1, or
 * rng->normals), x = T^-1(z) [n, d], log_q (optional) = log N(z) - log|det dx/dz| [n] */
NFMC_API int nfmc_flow_wide_sample(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden, const float* theta,
                          int32_t transposed, const nfmc_rng* rng, int64_t chain0, float* x, float* log_q, int64_t n, void* stream);
/* backward sweep of a pass whose OUTPUT y [n, d] and output cotangent grad_y [n, d] are given (reverse KL: inverse = 1,
 * y = x = T^-1(z), grad_y = grad U(x)): grad_theta (+)= d/dtheta [ sum_i (grad_y . y)(theta) -/+ log|det| ], i.e. the
 * gradient of sum_i [U(x_i) - log|det dx/dz|] (inverse = 1) or of sum_i [f(z_i) - log|det dz/dx|] (inverse = 0);
 * grad_in (optional) = the cotangent that reaches the pass input */
NFMC_API int nfmc_flow_wide_sweep(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden, const float* theta,
                         int32_t inverse, const float* y, const float* grad_y, int64_t n, float* grad_theta,
                         float* grad_in, int32_t accumulate, void* stream);
/* the sweep's cotangent path alone (NeuTra's latent gradient, nfmc/neutra.py:58-68 under autograd: grad_in = d/dz [U(x(z)) -
 * log|det dx/dz|] for inverse = 1, y = x, grad_y = grad U(x)); theta_t (optional): theta with every linear's weight transposed,
 * read by the conditioner's forward GEMMs */
NFMC_API int nfmc_flow_wide_pullback(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden, const float* theta,
                            const float* theta_t, int32_t inverse, const float* y, const float* grad_y, int64_t n,
                            float* grad_in, void* stream);
/* AdamW with the gradient multiplied by grad_scale first (1 / batch) */
NFMC_API int nfmc_adamw_step_scaled(float* theta, const float* grad, float grad_scale, float* exp_avg, float* exp_avg_sq,
                           int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                           void* stream);
/* one epoch of minibatch maximum likelihood on one GPU (per batch of perm[n]: loss + gradient, AdamW);
 * losses[batch] = summed loss of the batch; grad_theta is scratch of param_count floats */
NFMC_API int nfmc_flow_wide_fit_epoch(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden, float* theta,
                             float* exp_avg, float* exp_avg_sq, float* grad_theta, double* losses, const float* x,
                             const int64_t* perm, int64_t n, int64_t batch_size, float lr, float beta1, float beta2,
                             float eps, float weight_decay, int32_t step0, void* stream);

/* ---- deterministic Langevin Monte Carlo (nfmc/dlmc.py:44-119): its three particle updates; the MH correction that follows
 * each of them is nfmc_imh_steps(..., n_steps = 1, recompute_logq = 1) and the refit is nfmc_flow_fit_epoch ------------- */
/* x <- x - step * grad U(x)                                        (dlmc.py:60-62, the initial update) */
NFMC_API int nfmc_potential_step(const nfmc_potential* pot, float* x, int64_t n, float step, void* stream);
/* x <- x - step * grad_x [ U(x) + log q(x) ]                       (dlmc.py:86-88; grad log q by the reversible sweep) */
NFMC_API int nfmc_dlmc_update(const nfmc_potential* pot, const nfmc_realnvp* flow, float* x, int64_t n, float step, void* stream);
/* z <- z - step * (grad - z), elementwise over `count` floats       (dlmc.py:84, the latent update between T and T^-1) */
NFMC_API int nfmc_dlmc_latent_update(float* z, const float* grad, float step, int64_t count, void* stream);

/* ---- whole-run entry points --------------------------------------------------------------------------------------
 * JumpNFMC.sample (nfmc/jump.py:156-246) for device-resident chains: n_outer x [n_inner local steps (inner_kind 0 =
 * Langevin, 1 = HMC, 2 = random-walk Metropolis) + one NF jump], Philox noise, no sample sink.  The batch is cut into
 * slabs spread over internal streams so that kernels of different slabs overlap; chains keep their global index, the
 * result equals per-iteration nfmc_*_steps + nfmc_jump_step calls with rng step0 = local_step0 + it*n_inner / jump_step0
 * + it.  logq_scratch: n floats of device memory (may be NULL when jump_adjusted = 0).  Asynchronous: work is forked from and
 * joined back into `stream`. */
/* number of slabs (= kernel launches per local stage / per jump) the two whole-run entry points cut n chains into */
NFMC_API int64_t nfmc_jump_sample_slabs(int32_t d, int64_t n, int32_t host_buffers);
NFMC_API int nfmc_jump_sample_device(const nfmc_potential* pot, const nfmc_realnvp* flow, float* x, int64_t n,
                            int32_t inner_kind, int32_t n_outer, int32_t n_inner, float step_size, int32_t n_leapfrog,
                            const float* inv_mass_diag, int32_t local_adjusted, int32_t jump_adjusted, uint64_t seed,
                            uint64_t local_step0, uint64_t jump_step0, int64_t chain0, const nfmc_stats* local_stats,
                            const nfmc_stats* jump_stats, float* logq_scratch, void* stream);

/* ---- host-buffer entry point (end-to-end measurement; the call a reference-side plugin would make) ------
 * Runs `n_outer` iterations of [n_inner local steps (kind 0 = MALA, 1 = HMC) + one NF jump] on host data:
 * copies x_host [n,d] to the device, runs, copies the final state back into x_host, and returns the pooled
 * moments (sum_x_host/sum_x2_host [d], doubles) and counters (counts_host[4] local, counts_host[4..8) jump).
 * Device buffers are taken from `workspace` (>= nfmc_jump_workspace_bytes bytes of device memory). */
NFMC_API int64_t nfmc_jump_workspace_bytes(int32_t d, int64_t n, int64_t blob_floats);
NFMC_API int nfmc_jump_sample_host(const nfmc_potential* pot_host_desc, const float* pot_params_host, int64_t pot_params_floats,
                          const nfmc_realnvp* flow_host_desc, const float* blob_host, float* x_host, int64_t n,
                          int32_t inner_kind, int32_t n_outer, int32_t n_inner, float step_size, int32_t n_leapfrog,
                          uint64_t seed, int64_t chain0, double* sum_x_host, double* sum_x2_host,
                          unsigned long long* counts_host, void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NFMC_B200_H */
