// ext_kernels.cu -- the sampler arithmetic around an EXTERNAL target: U(x) and grad U(x) are supplied by the caller
// (a Python callable differentiated by autograd on the device, which is how the reference treats every target:
// mcmc/langevin.py:66-68, mcmc/hmc.py:40-48, sample.py:34-36), everything else -- proposal, proposal potentials,
// leapfrog updates, Hamiltonians, log-ratio, accept test, masked overwrite, running moments, counters, sample sink --
// runs in these kernels with the same roundings as the fused kernels (mala_kernel.cu, hmc_kernel.cu, flow_kernels.cu).
//
// Layout: one warp per chain row, lanes stride over the d coordinates (a row is contiguous, so every access is
// coalesced); persistent grid, warps stride over rows.  All of them are plain HBM streaming kernels.
#include "host_common.cuh"
#include "chain_kernel.cuh"

namespace nfmc {

constexpr int kExtThreads = 256;
constexpr int kExtWarps = kExtThreads / 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// x' = x - tau/m^2 grad U(x) + sqrt(2 tau)/m xi   (langevin.py:74-76), or x' = x + m xi (mh.py:52-56)
__global__ void __launch_bounds__(kExtThreads) ext_langevin_propose_kernel(
    const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ noise, const float* __restrict__ imd,
    float tau, float sqrt_2tau, int random_walk, long long total, int d, float* __restrict__ xp) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float xi = __ldg(noise + i), xv = __ldg(x + i);
    float out;
    if (random_walk) {
      out = imd ? fmaf(__ldg(imd + (int)(i % d)), xi, xv) : xv + xi;
    } else if (!imd) {
      out = fmaf(sqrt_2tau, xi, fmaf(-tau, __ldg(g + i), xv));
    } else {
      const float m = __ldg(imd + (int)(i % d));
      out = fmaf(__fdiv_rn(sqrt_2tau, m), xi, fmaf(-__fdiv_rn(tau, m * m), __ldg(g + i), xv));
    }
    xp[i] = out;
  }
}

// log-ratio of one Langevin step: both proposal potentials as langevin.py:31-42, combined as util.py:392
//   qf = || x' - x + tau A grad U(x)  ||^2_{A^-1} / (4 tau),  qr = || x - x' + tau A grad U(x') ||^2_{A^-1} / (4 tau),  A = 1/m^2
//   log_ratio = (-U') - (-U) + (-qr) - (-qf);   random walk (mh.py:59): (-U') - (-U)
__global__ void __launch_bounds__(kExtThreads) ext_langevin_ratio_kernel(
    const float* __restrict__ x, const float* __restrict__ xp, const float* __restrict__ g, const float* __restrict__ gp,
    const float* __restrict__ u, const float* __restrict__ up, const float* __restrict__ imd, float tau, int random_walk,
    long long n, int d, float* __restrict__ log_ratio) {
  const int lane = threadIdx.x & 31;
  const float inv4tau = __fdiv_rn(1.f, 4.f * tau);
  for (long long row = (long long)blockIdx.x * kExtWarps + (threadIdx.x >> 5); row < n; row += (long long)gridDim.x * kExtWarps) {
    float qf = 0.f, qr = 0.f;
    if (!random_walk) {
      const long long b = row * d;
      for (int c = lane; c < d; c += 32) {
        const float xv = __ldg(x + b + c), pv = __ldg(xp + b + c), gv = __ldg(g + b + c), gpv = __ldg(gp + b + c);
        if (!imd) {
          const float tf = pv - xv + tau * gv, tr = xv - pv + tau * gpv;
          qf = fmaf(tf, tf, qf);
          qr = fmaf(tr, tr, qr);
        } else {
          const float m = __ldg(imd + c);
          const float a = __fdiv_rn(1.f, m * m), ta = tau * a, ia = __fdiv_rn(1.f, a);
          const float tf = pv - xv + ta * gv, tr = xv - pv + ta * gpv;
          qf = fmaf(tf * ia, tf, qf);
          qr = fmaf(tr * ia, tr, qr);
        }
      }
      qf = warp_sum(qf) * inv4tau;
      qr = warp_sum(qr) * inv4tau;
    }
    if (lane == 0) {
      const float uu = __ldg(u + row), uup = __ldg(up + row);
      log_ratio[row] = random_walk ? ((-uup) - (-uu) + 0.f - 0.f) : ((-uup) - (-uu) + (-qr) - (-qf));
    }
  }
}

// p = xi / sqrt(m), kinetic = sum p^2 m   (hmc.py:100,104; the 1/2 is applied where the Hamiltonian is formed)
__global__ void __launch_bounds__(kExtThreads) ext_hmc_momentum_kernel(const float* __restrict__ noise, const float* __restrict__ imd,
                                                                       long long n, int d, float* __restrict__ p,
                                                                       float* __restrict__ kinetic) {
  const int lane = threadIdx.x & 31;
  for (long long row = (long long)blockIdx.x * kExtWarps + (threadIdx.x >> 5); row < n; row += (long long)gridDim.x * kExtWarps) {
    const long long b = row * d;
    float k = 0.f;
    for (int c = lane; c < d; c += 32) {
      float v = __ldg(noise + b + c), m = 1.f;
      if (imd) { m = __ldg(imd + c); v *= __fdiv_rn(1.f, sqrtf(m)); }
      k = fmaf(v * v, m, k);
      p[b + c] = v;
    }
    k = warp_sum(k);
    if (lane == 0) kinetic[row] = k;
  }
}

// `kicks` half-kicks p -= tau/2 grad U (separately rounded, hmc.py:51-53), then optionally the drift x += tau p m (hmc.py:56-58)
__global__ void __launch_bounds__(kExtThreads) ext_hmc_leapfrog_kernel(float* __restrict__ x, float* __restrict__ p,
                                                                       const float* __restrict__ g, const float* __restrict__ imd,
                                                                       float tau, int kicks, int drift, long long total, int d) {
  const float half_tau = tau / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float gv = __ldg(g + i);
    float pv = p[i];
    for (int k = 0; k < kicks; ++k) pv = fmaf(-half_tau, gv, pv);
    p[i] = pv;
    if (drift) x[i] = fmaf(tau, imd ? pv * __ldg(imd + (int)(i % d)) : pv, x[i]);
  }
}

// log_ratio = -H1 - (-H0),  H = U + 1/2 sum p^2 m   (hmc.py:103-111)
__global__ void __launch_bounds__(kExtThreads) ext_hmc_ratio_kernel(const float* __restrict__ p, const float* __restrict__ imd,
                                                                    const float* __restrict__ u0, const float* __restrict__ kin0,
                                                                    const float* __restrict__ u1, long long n, int d,
                                                                    float* __restrict__ log_ratio) {
  const int lane = threadIdx.x & 31;
  for (long long row = (long long)blockIdx.x * kExtWarps + (threadIdx.x >> 5); row < n; row += (long long)gridDim.x * kExtWarps) {
    const long long b = row * d;
    float k = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float v = __ldg(p + b + c);
      k = imd ? fmaf(v * v, __ldg(imd + c), k) : fmaf(v, v, k);
    }
    k = warp_sum(k);
    if (lane == 0) {
      const float h0 = __ldg(u0 + row) + 0.5f * __ldg(kin0 + row), h1 = __ldg(u1 + row) + 0.5f * k;
      log_ratio[row] = -h1 - (-h0);
    }
  }
}

// flow-proposal jump / IMH (util.py:392 as used at jump.py:224-229, imh.py:226-231):
//   log_ratio = (-U') - (-U) + log q(x) - log q(x')
__global__ void ext_jump_ratio_kernel(const float* __restrict__ u, const float* __restrict__ up, const float* __restrict__ lq,
                                      const float* __restrict__ lqp, long long n, float* __restrict__ log_ratio) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    log_ratio[i] = (-__ldg(up + i)) - (-__ldg(u + i)) + __ldg(lq + i) - __ldg(lqp + i);
}

struct ExtAcceptArgs {
  float* x;
  const float* xp;
  const float* log_ratio;
  const float* uniforms;
  int adjusted;
  long long n;
  int d;
  float* aux_a; const float* aux_a_p;      // [n]   e.g. U     <- U'
  float* aux_b; const float* aux_b_p;      // [n]   e.g. log q <- log q'
  float* aux_g; const float* aux_g_p;      // [n,d] e.g. grad U <- grad U'
  StatsArgs stats;
  float* sink_row;                         // [n, d] destination of this step's sample row, or nullptr
};

// accept iff log u < log_ratio (langevin.py:106, hmc.py:112-113, jump.py:232); x[mask] = x'[mask] (mcmc/base.py:77) with
// the caches that travel with the state; running moments of the post-accept state (mcmc/base.py:86); counters
// counts[0] += accepted, counts[1] += n, counts[2] += chains whose log-ratio is not finite (they reject)
__global__ void __launch_bounds__(kExtThreads) ext_accept_kernel(const ExtAcceptArgs A) {
  extern __shared__ float sm_mom[];        // [2 d]: per-CTA fp32 partial sums
  __shared__ unsigned int sm_cnt[2];
  const int d = A.d;
  for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) sm_mom[i] = 0.f;
  if (threadIdx.x < 2) sm_cnt[threadIdx.x] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  unsigned int n_acc = 0, n_bad = 0;
  for (long long row = (long long)blockIdx.x * kExtWarps + (threadIdx.x >> 5); row < A.n; row += (long long)gridDim.x * kExtWarps) {
    bool accept = true;
    if (A.adjusted) {
      const float lr = __ldg(A.log_ratio + row);
      accept = logf(__ldg(A.uniforms + row)) < lr;
      if (!(fabsf(lr) <= 3.0e38f) && lane == 0) ++n_bad;
    }
    const long long b = row * d;
    for (int c = lane; c < d; c += 32) {
      float v;
      if (accept) {
        v = __ldg(A.xp + b + c);
        A.x[b + c] = v;
        if (A.aux_g) A.aux_g[b + c] = __ldg(A.aux_g_p + b + c);
      } else {
        v = A.x[b + c];
      }
      atomicAdd(sm_mom + c, v);
      atomicAdd(sm_mom + d + c, v * v);
      if (A.sink_row) A.sink_row[b + c] = v;
    }
    if (lane == 0 && accept) {
      ++n_acc;
      if (A.aux_a) A.aux_a[row] = __ldg(A.aux_a_p + row);
      if (A.aux_b) A.aux_b[row] = __ldg(A.aux_b_p + row);
    }
  }
  n_acc = __reduce_add_sync(0xffffffffu, n_acc);
  n_bad = __reduce_add_sync(0xffffffffu, n_bad);
  if (lane == 0) {
    if (n_acc) atomicAdd(sm_cnt + 0, n_acc);
    if (n_bad) atomicAdd(sm_cnt + 1, n_bad);
  }
  __syncthreads();
  if (A.stats.sum_x && A.stats.sum_x2) {
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
      atomicAdd(A.stats.sum_x + i, (double)sm_mom[i]);
      atomicAdd(A.stats.sum_x2 + i, (double)sm_mom[d + i]);
    }
  }
  if (A.stats.counts && threadIdx.x == 0) {
    if (sm_cnt[0]) atomicAdd(A.stats.counts + 0, (unsigned long long)sm_cnt[0]);
    if (sm_cnt[1]) atomicAdd(A.stats.counts + 2, (unsigned long long)sm_cnt[1]);
    if (blockIdx.x == 0) atomicAdd(A.stats.counts + 1, (unsigned long long)A.n);
  }
}

// ---- elliptical slice sampling around an external negative log-likelihood (mcmc/ess.py:12-64, identity prior) -----------
constexpr float kExtPiF = 3.14159274101257324f;       // (float) torch.pi
constexpr float kExtTwoPiF = 6.28318548202514648f;    // (float) (2 * torch.pi)

// the 2 + M scalar uniforms of one step as the fused kernel draws them (ess_kernel.cu): Philox stream 2, counter quad
// i / 4, word i % 4, lane 0 of the chain's group
__global__ void ext_ess_uniforms_kernel(uint64_t seed, uint64_t step, long long chain0, long long n, int n_uni, float* __restrict__ out) {
  const PhiloxKeys PK = philox_keys(seed);
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
    const RngKey ukey = make_rng_key(seed, 2u, step, (uint64_t)(chain0 + c));
    uint4 uq = make_uint4(0u, 0u, 0u, 0u);
    for (int i = 0; i < n_uni; ++i) {
      if ((i & 3) == 0) uq = rng_quad(PK, ukey, i >> 2, 0);
      const uint32_t w = (i & 3) == 0 ? uq.x : (i & 3) == 1 ? uq.y : (i & 3) == 2 ? uq.z : uq.w;
      out[c * n_uni + i] = uniform_from_bits(w);
    }
  }
}

// threshold log y = -nll(f) + log u (ess.py:35-36), initial angle theta = 2 pi u' and bracket [theta - 2 pi, theta] (:39-41)
// state[n][4] = {log_y, theta, theta_min, theta_max}; found[n] = 0
__global__ void ext_ess_begin_kernel(const float* __restrict__ nll_cur, const float* __restrict__ uniforms, int n_uni, long long n,
                                     float4* __restrict__ state, int* __restrict__ found) {
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
    const float log_y = __fadd_rn(-__ldg(nll_cur + c), logf(__ldg(uniforms + c * n_uni)));
    const float theta = __fmul_rn(__fmul_rn(__ldg(uniforms + c * n_uni + 1), 2.f), kExtPiF);
    state[c] = make_float4(log_y, theta, __fsub_rn(theta, kExtTwoPiF), theta);
    found[c] = 0;
  }
}

// f' = f cos(theta) + nu sin(theta)   (ess.py:46)
__global__ void __launch_bounds__(kExtThreads) ext_ess_rotate_kernel(const float* __restrict__ f, const float* __restrict__ nu,
                                                                     const float4* __restrict__ state, long long n, int d,
                                                                     float* __restrict__ fp) {
  const int lane = threadIdx.x & 31;
  for (long long row = (long long)blockIdx.x * kExtWarps + (threadIdx.x >> 5); row < n; row += (long long)gridDim.x * kExtWarps) {
    float sn, cs;
    sincosf(state[row].y, &sn, &cs);
    const long long b = row * d;
    for (int c = lane; c < d; c += 32)
      fp[b + c] = __fadd_rn(__fmul_rn(__ldg(f + b + c), cs), __fmul_rn(__ldg(nu + b + c), sn));
  }
}

// one bracket round (ess.py:47-62): a chain that has not found its point yet and whose proposal lies above the threshold
// takes it; every chain shrinks its bracket towards 0 and draws the next angle from it
__global__ void __launch_bounds__(kExtThreads) ext_ess_update_kernel(float* __restrict__ f, const float* __restrict__ fp,
                                                                     float* __restrict__ nll_cur, const float* __restrict__ nll_p,
                                                                     float4* __restrict__ state, int* __restrict__ found,
                                                                     const float* __restrict__ uniforms, int n_uni, int round,
                                                                     long long n, int d) {
  const int lane = threadIdx.x & 31;
  for (long long row = (long long)blockIdx.x * kExtWarps + (threadIdx.x >> 5); row < n; row += (long long)gridDim.x * kExtWarps) {
    float4 st = state[row];
    const int was_found = found[row];
    const float up = __ldg(nll_p + row);
    const bool upd = (-up > st.x) && !was_found;                                   // :47,50
    if (upd) {
      const long long b = row * d;
      for (int c = lane; c < d; c += 32) f[b + c] = __ldg(fp + b + c);
    }
    __syncwarp();
    if (lane == 0) {
      if (upd) { nll_cur[row] = up; found[row] = 1; }
      if (st.y < 0.f) st.z = st.y; else st.w = st.y;                               // :53-55
      st.y = __fadd_rn(__fmul_rn(__ldg(uniforms + row * n_uni + 2 + round), __fsub_rn(st.w, st.z)), st.z);   // :58-59
      state[row] = st;
    }
  }
}

static int ext_grid(long long work_items, int per_cta) {
  long long grid = (work_items + per_cta - 1) / per_cta;
  const long long cap = 8ll * sm_count();          // 8 resident CTAs of 256 threads per SM
  if (grid > cap) grid = cap;
  return (int)(grid < 1 ? 1 : grid);
}

}  // namespace nfmc

using namespace nfmc;

static int ext_check(bool ok, const char* who) { return ok ? 0 : set_error(std::string(who) + ": bad arguments"); }

extern "C" int nfmc_ext_langevin_propose(const float* x, const float* grad, const float* noise, const float* inv_mass_diag,
                                         float step_size, int32_t random_walk, int64_t n, int32_t d, float* x_prime, void* stream) {
  if (int e = ext_check(x && noise && x_prime && (random_walk || grad) && n >= 1 && d >= 1 && d <= NFMC_MAX_DIM &&
                        (random_walk || step_size > 0.f), "ext_langevin_propose")) return e;
  const long long total = (long long)n * d;
  const float s2t = (float)sqrt(2.0 * (double)step_size);
  ext_langevin_propose_kernel<<<ext_grid(total, kExtThreads * 4), kExtThreads, 0, (cudaStream_t)stream>>>(
      x, grad, noise, inv_mass_diag, step_size, s2t, random_walk, total, d, x_prime);
  return check_cuda(cudaGetLastError(), "ext_langevin_propose_kernel launch");
}

extern "C" int nfmc_ext_langevin_log_ratio(const float* x, const float* x_prime, const float* grad, const float* grad_prime,
                                           const float* u, const float* u_prime, const float* inv_mass_diag, float step_size,
                                           int32_t random_walk, int64_t n, int32_t d, float* log_ratio, void* stream) {
  if (int e = ext_check(u && u_prime && log_ratio && n >= 1 && d >= 1 && d <= NFMC_MAX_DIM &&
                        (random_walk || (x && x_prime && grad && grad_prime && step_size > 0.f)), "ext_langevin_log_ratio")) return e;
  ext_langevin_ratio_kernel<<<ext_grid(n, kExtWarps), kExtThreads, 0, (cudaStream_t)stream>>>(
      x, x_prime, grad, grad_prime, u, u_prime, inv_mass_diag, step_size, random_walk, n, d, log_ratio);
  return check_cuda(cudaGetLastError(), "ext_langevin_ratio_kernel launch");
}

extern "C" int nfmc_ext_hmc_momentum(const float* noise, const float* inv_mass_diag, int64_t n, int32_t d, float* p, float* kinetic,
                                     void* stream) {
  if (int e = ext_check(noise && p && kinetic && n >= 1 && d >= 1 && d <= NFMC_MAX_DIM, "ext_hmc_momentum")) return e;
  ext_hmc_momentum_kernel<<<ext_grid(n, kExtWarps), kExtThreads, 0, (cudaStream_t)stream>>>(noise, inv_mass_diag, n, d, p, kinetic);
  return check_cuda(cudaGetLastError(), "ext_hmc_momentum_kernel launch");
}

extern "C" int nfmc_ext_hmc_leapfrog(float* x, float* p, const float* grad, const float* inv_mass_diag, float step_size,
                                     int32_t kicks, int32_t drift, int64_t n, int32_t d, void* stream) {
  if (int e = ext_check(x && p && grad && n >= 1 && d >= 1 && d <= NFMC_MAX_DIM && kicks >= 0 && kicks <= 2, "ext_hmc_leapfrog")) return e;
  const long long total = (long long)n * d;
  ext_hmc_leapfrog_kernel<<<ext_grid(total, kExtThreads * 4), kExtThreads, 0, (cudaStream_t)stream>>>(
      x, p, grad, inv_mass_diag, step_size, kicks, drift, total, d);
  return check_cuda(cudaGetLastError(), "ext_hmc_leapfrog_kernel launch");
}

extern "C" int nfmc_ext_hmc_log_ratio(const float* p, const float* inv_mass_diag, const float* u0, const float* kinetic0,
                                      const float* u1, int64_t n, int32_t d, float* log_ratio, void* stream) {
  if (int e = ext_check(p && u0 && kinetic0 && u1 && log_ratio && n >= 1 && d >= 1 && d <= NFMC_MAX_DIM, "ext_hmc_log_ratio")) return e;
  ext_hmc_ratio_kernel<<<ext_grid(n, kExtWarps), kExtThreads, 0, (cudaStream_t)stream>>>(p, inv_mass_diag, u0, kinetic0, u1, n, d, log_ratio);
  return check_cuda(cudaGetLastError(), "ext_hmc_ratio_kernel launch");
}

extern "C" int nfmc_ext_jump_log_ratio(const float* u, const float* u_prime, const float* log_q, const float* log_q_prime, int64_t n,
                                       float* log_ratio, void* stream) {
  if (int e = ext_check(u && u_prime && log_q && log_q_prime && log_ratio && n >= 1, "ext_jump_log_ratio")) return e;
  ext_jump_ratio_kernel<<<ext_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(u, u_prime, log_q, log_q_prime, n, log_ratio);
  return check_cuda(cudaGetLastError(), "ext_jump_ratio_kernel launch");
}

extern "C" int nfmc_ext_accept(float* x, const float* x_prime, const float* log_ratio, const float* uniforms, int32_t adjusted,
                               int64_t n, int32_t d, float* aux_a, const float* aux_a_prime, float* aux_b, const float* aux_b_prime,
                               float* aux_grad, const float* aux_grad_prime, const nfmc_stats* stats, const nfmc_sink* sink,
                               int32_t sink_step, void* stream) {
  if (int e = ext_check(x && x_prime && n >= 1 && d >= 1 && d <= NFMC_MAX_DIM && (!adjusted || (log_ratio && uniforms)) &&
                        (!aux_a == !aux_a_prime) && (!aux_b == !aux_b_prime) && (!aux_grad == !aux_grad_prime), "ext_accept")) return e;
  ExtAcceptArgs A;
  A.x = x; A.xp = x_prime; A.log_ratio = log_ratio; A.uniforms = uniforms; A.adjusted = adjusted; A.n = n; A.d = d;
  A.aux_a = aux_a; A.aux_a_p = aux_a_prime; A.aux_b = aux_b; A.aux_b_p = aux_b_prime; A.aux_g = aux_grad; A.aux_g_p = aux_grad_prime;
  A.stats.sum_x = stats ? stats->sum_x : nullptr;
  A.stats.sum_x2 = stats ? stats->sum_x2 : nullptr;
  A.stats.counts = stats ? reinterpret_cast<unsigned long long*>(stats->counts) : nullptr;
  A.sink_row = nullptr;
  if (sink && sink->samples) {                                                    // the thinning rule of sink_store (chain_kernel.cuh)
    if (sink->thinning < 1) return set_error("ext_accept: sink thinning must be >= 1");
    const long long idx = sink->seen0 + sink_step;
    if (idx % sink->thinning == 0) {
      const long long first = (sink->seen0 + sink->thinning - 1) / sink->thinning;
      A.sink_row = sink->samples + (idx / sink->thinning - first) * n * d;
    }
  }
  ext_accept_kernel<<<ext_grid(n, kExtWarps * 4), kExtThreads, (size_t)2 * d * sizeof(float), (cudaStream_t)stream>>>(A);
  return check_cuda(cudaGetLastError(), "ext_accept_kernel launch");
}

extern "C" int nfmc_ext_ess_uniforms(uint64_t seed, uint64_t step, int64_t chain0, int64_t n, int32_t n_uniforms, float* out, void* stream) {
  if (int e = ext_check(out && n >= 1 && n_uniforms >= 2, "ext_ess_uniforms")) return e;
  ext_ess_uniforms_kernel<<<ext_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(seed, step, chain0, n, n_uniforms, out);
  return check_cuda(cudaGetLastError(), "ext_ess_uniforms_kernel launch");
}

extern "C" int nfmc_ext_ess_begin(const float* nll_cur, const float* uniforms, int32_t n_uniforms, int64_t n, float* state, int32_t* found,
                                  void* stream) {
  if (int e = ext_check(nll_cur && uniforms && state && found && n >= 1 && n_uniforms >= 2 &&
                        (reinterpret_cast<uintptr_t>(state) & 15) == 0, "ext_ess_begin")) return e;
  ext_ess_begin_kernel<<<ext_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(nll_cur, uniforms, n_uniforms, n, reinterpret_cast<float4*>(state), found);
  return check_cuda(cudaGetLastError(), "ext_ess_begin_kernel launch");
}

extern "C" int nfmc_ext_ess_rotate(const float* f, const float* nu, const float* state, int64_t n, int32_t d, float* f_prime, void* stream) {
  if (int e = ext_check(f && nu && state && f_prime && n >= 1 && d >= 1 && d <= NFMC_MAX_DIM, "ext_ess_rotate")) return e;
  ext_ess_rotate_kernel<<<ext_grid(n, kExtWarps), kExtThreads, 0, (cudaStream_t)stream>>>(f, nu, reinterpret_cast<const float4*>(state), n, d, f_prime);
  return check_cuda(cudaGetLastError(), "ext_ess_rotate_kernel launch");
}

extern "C" int nfmc_ext_ess_update(float* f, const float* f_prime, float* nll_cur, const float* nll_prime, float* state, int32_t* found,
                                   const float* uniforms, int32_t n_uniforms, int32_t round, int64_t n, int32_t d, void* stream) {
  if (int e = ext_check(f && f_prime && nll_cur && nll_prime && state && found && uniforms && n >= 1 && d >= 1 && d <= NFMC_MAX_DIM &&
                        round >= 0 && round + 2 < n_uniforms, "ext_ess_update")) return e;
  ext_ess_update_kernel<<<ext_grid(n, kExtWarps), kExtThreads, 0, (cudaStream_t)stream>>>(f, f_prime, nll_cur, nll_prime,
      reinterpret_cast<float4*>(state), found, uniforms, n_uniforms, round, n, d);
  return check_cuda(cudaGetLastError(), "ext_ess_update_kernel launch");
}
