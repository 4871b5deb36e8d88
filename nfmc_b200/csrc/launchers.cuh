// launchers.cuh -- launch-argument structs and the per-E launcher templates that tie the C-ABI dispatch
// (local_api.cu / flow_api.cu) to the kernels, which are compiled one translation unit per slots-per-half E
// (-DNFMC_ONLY_E=<E>) so that the build parallelises.
#pragma once
#include "chain_kernel.cuh"
#include "flow_args.cuh"
#include "host_common.cuh"

namespace nfmc {

struct LocalArgs {
  ChainArgs c;
  float tau;          // step size
  float sqrt_2tau;    // (float) sqrt(2 * (double) tau)
  const float* imd;   // inverse mass diagonal [d] or nullptr (= ones)
  int adjusted;
  int n_leapfrog;
  int random_walk;    // 1: random-walk MH proposal x' = x + imd * xi, ratio = U(x) - U(x') (mcmc/mh.py:44-73)
};


struct EssArgs {
  ChainArgs c;          // c.pot is the negative log-likelihood (the prior is N(0, I))
  int max_iterations;   // bracket-shrinking rounds per step (ESSParameters.max_ess_step_iterations, mcmc/ess.py:73)
};

struct JumpArgs {
  ChainArgs c;
  FlowArgs f;
  int pot_kind;
  float* logq_x;       // [n] cached log q(x) (IMH) or nullptr
  int recompute_logq;  // compute log q(x) by a forward pass (jump.py:218; AdaptiveIMH imh.py:133)
  int adjusted;
};

struct AcceptArgs {
  ChainArgs c;
  int pot_kind;
  int adjusted;
  const float* x_prime;   // [n, d] proposal
  const float* z;         // [n, d] base draw that produced it
  const float* ld_inv;    // [n] log|det dx'/dz|
  const float* logq_x;    // [n] log q(x)
  const float* uniforms;  // [n]
  float* logq_cache;      // [n] optional: receives log q(x') where accepted (IMH cache)
};

struct NeutraArgs {
  ChainArgs c;
  FlowArgs f;
  int pot_kind;
  float tau;
  const float* imd;
  int n_leapfrog;
  int stash;   // shared memory holds a conditioner stash (flow.cuh) behind the mass table
  int adjusted;  // NeuTra MH only: 0 = plain random walk
};

struct TrainArgs {
  FlowArgs f;             // blob = packed parameters (small conditioner path only)
  float* grad;            // [blob_floats] gradient accumulator in blob layout
  double* loss;           // [1] += sum over rows of the per-row loss (or nullptr)
  const float* x;         // maximum likelihood: data [*, d]
  const long long* rows;  // maximum likelihood: optional row indices [n] into x (a shuffled minibatch)
  long long n;            // rows in this launch
  int stash;              // keep the conditioner outputs of the loss pass in shared memory for the backward sweep
  int kl;                 // 1: reverse KL (base draw from rng, potential below); 0: maximum likelihood
  int pot_kind;
  PotParams pot;
  RngArgs rng;
  long long chain0;
  float* x_rw;            // DLMC update: particles [n, d], updated in place
  float step;             // DLMC update: step size
};

struct TessArgs {
  ChainArgs c;          // c.x = latent state u; c.pot = the potential of tess.py's `potential` argument
  FlowArgs f;
  int pot_kind;
  int max_iterations;   // bracket rounds per step (TESSParameters.max_ess_step_iterations)
};

enum { PASS_FORWARD = 0, PASS_INVERSE = 1, PASS_LOGPROB = 2 };

template <int E> int launch_mala(int pot_kind, bool exact, const LocalArgs& A, int grid, size_t smem, cudaStream_t s);
template <int E> int launch_hmc(int pot_kind, bool exact, const LocalArgs& A, int grid, size_t smem, cudaStream_t s);
template <int E> int launch_ess(int pot_kind, bool exact, const EssArgs& A, int grid, size_t smem, cudaStream_t s);
template <int E> int launch_flow_pass(const FlowArgs& A, int mode, const float* in, float* out, float* aux, long long n,
                                      int grid, size_t smem, cudaStream_t s);
template <int E> int launch_flow_sample(const FlowArgs& A, const RngArgs& R, long long chain0, float* x, float* logq,
                                        long long n, int grid, size_t smem, cudaStream_t s);
template <int E> int launch_flow_train(const TrainArgs& A, int grid, bool shared_grad, cudaStream_t s);
template <int E> int launch_jump(const JumpArgs& A, int grid, size_t smem, cudaStream_t s);
template <int E> int launch_jump_propose_accept(const JumpArgs& A, int grid, size_t smem, cudaStream_t s);
template <int E> int launch_jump_accept(const AcceptArgs& A, int grid, size_t smem, cudaStream_t s);
template <int E> int launch_neutra_hmc(const NeutraArgs& A, int grid, size_t smem, cudaStream_t s);
template <int E> int launch_flow_dlmc(const TrainArgs& A, int grid, cudaStream_t s);
template <int E> int launch_tess(const TessArgs& A, int grid, size_t smem, cudaStream_t s);
template <int E> int launch_neutra_mh(const NeutraArgs& A, int grid, size_t smem, cudaStream_t s);
template <int E> int launch_neutra_potential(const FlowArgs& FA, int pot_kind, const PotParams& P, const float* z, float* u,
                                             float* grad, long long n, int grid, size_t smem, cudaStream_t s);

template <int E> int launch_neutra_pullback(const FlowArgs& FA, const float* z, const float* gx, float* gz, float* ld, long long n,
                                            int grid, size_t smem, cudaStream_t s);

#define NFMC_SET_SMEM_RET(kern, bytes)                                                                     \
  do {                                                                                                     \
    if ((bytes) > 227 * 1024) return set_error("shared-memory plan exceeds 227 KB (conditioner too large for the generic kernel)"); \
    if ((bytes) > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); \
  } while (0)

}  // namespace nfmc
