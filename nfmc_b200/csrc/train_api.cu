// train_api.cu -- C-ABI entry points of flow training: parameter pack (module order -> kernel blob), loss + gradient
// launches (train_kernels.cu), gradient unpack (blob layout -> module order, chain rule through the merged elementwise
// affines), AdamW, and a one-call epoch driver.  Register-resident conditioner path only (M = 2, H <= 8).
//
// Parameter vector theta = the module's parameters in state_dict order (oracle/realnvp_ref.py, nfmc_b200/flow.py):
//   affine_0.value[d][2] | Lc x { W1[H][da] b1[H] Wl[2 db][H] bl[2 db] actnorm_l.value[d][2] } | affine_T.value[d][2]
//   | actnorm_T.value[d][2],   value[p] = (u_a, u_b),  alpha = exp(log(1-m) + u_a/2) + m,  beta = u_b/2.
// Merged affine group g (blob table g, preceded by g reversals): g = 0 -> {affine_0}, g = l+1 -> {actnorm_l}; the last
// group also holds {affine_T, actnorm_T}.  Physical coordinate k of group g is logical p = k (g even) or d-1-k (g odd).
#include <algorithm>
#include <cmath>
#include "launchers.cuh"

using namespace nfmc;

namespace {

struct Dims {
  int d, da, db, Lc, H;
  int cpl_theta;   // floats per coupling block in theta (weights + its act-norm)
  int cpl_w;       // ... weights only
  long long n_theta;
  __host__ __device__ int theta_aff0() const { return 0; }
  __host__ __device__ int theta_cpl(int l) const { return 2 * d + l * cpl_theta; }
  __host__ __device__ int theta_act(int l) const { return theta_cpl(l) + cpl_w; }
  __host__ __device__ int theta_aff_tail() const { return 2 * d + Lc * cpl_theta; }
  __host__ __device__ int theta_act_tail() const { return theta_aff_tail() + 2 * d; }
};

Dims make_dims(int d, int Lc, int H) {
  Dims D;
  D.d = d; D.da = d / 2; D.db = d - D.da; D.Lc = Lc; D.H = H;
  D.cpl_w = H * D.da + H + 2 * D.db * H + 2 * D.db;
  D.cpl_theta = D.cpl_w + 2 * d;
  D.n_theta = 2ll * d + (long long)Lc * D.cpl_theta + 4ll * d;
  return D;
}

// members of affine group g: offsets into theta of their value[d][2] tables; returns the count (1 or 3)
__device__ __forceinline__ int group_members(const Dims& D, int g, int (&off)[3]) {
  off[0] = g == 0 ? D.theta_aff0() : D.theta_act(g - 1);
  if (g != D.Lc) return 1;
  off[1] = D.theta_aff_tail();
  off[2] = D.theta_act_tail();
  return 3;
}

__device__ __forceinline__ float alpha_of(float ua) { return expf(kLogOneMinusM + 0.5f * ua) + kMinScale; }

// ---- theta -> blob ---------------------------------------------------------------------------------------------------
__global__ void flow_pack_kernel(Dims D, const float* __restrict__ theta, float* __restrict__ blob) {
  const int d = D.d, da = D.da, db = D.db, H = D.H;
  const int n_aff = (D.Lc + 1) * d;
  const int cpl_blob = flow_coupling_floats(d, 2, H);
  const int off_const = (D.Lc + 1) * 4 * d;
  const long long total = (long long)n_aff + (long long)D.Lc * cpl_blob;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (i < n_aff) {
      const int g = (int)(i / d), k = (int)(i % d);
      const int p = (g & 1) ? d - 1 - k : k;
      int off[3];
      const int m = group_members(D, g, off);
      float alpha = 1.f, beta = 0.f;
      for (int q = 0; q < m; ++q) {                      // y = a (alpha x + beta) + b
        const float a = alpha_of(theta[off[q] + 2 * p]), b = 0.5f * theta[off[q] + 2 * p + 1];
        alpha = a * alpha;
        beta = fmaf(a, beta, b);
      }
      const float ra = 1.f / alpha;
      float* t = blob + (long long)g * 4 * d;
      t[2 * k] = alpha; t[2 * k + 1] = beta;
      t[2 * d + 2 * k] = ra; t[2 * d + 2 * k + 1] = -beta * ra;
    } else {
      const long long c = i - n_aff;
      const int l = (int)(c / cpl_blob), r = (int)(c % cpl_blob);
      const bool odd = ((l + 1) & 1) != 0;
      const float* th = theta + D.theta_cpl(l);
      const float* W1 = th;                 // [H][da]
      const float* b1 = W1 + H * da;        // [H]
      const float* Wl = b1 + H;             // [2 db][H], row 2 t + c
      const float* bl = Wl + 2 * db * H;    // [2 db]
      float v = 0.f;
      const int n_w1 = da * kSmallH, n_b1 = kSmallH, n_wl = db * 2 * kSmallH;
      if (r < n_w1) {
        const int ks = r / kSmallH, h = r % kSmallH;
        if (h < H) v = W1[h * da + (odd ? da - 1 - ks : ks)];
      } else if (r < n_w1 + n_b1) {
        const int h = r - n_w1;
        if (h < H) v = b1[h];
      } else if (r < n_w1 + n_b1 + n_wl) {
        const int q = r - n_w1 - n_b1;
        const int t = q / (2 * kSmallH), cc = (q / kSmallH) & 1, h = q % kSmallH;
        if (h < H) v = Wl[(2 * (odd ? db - 1 - t : t) + cc) * H + h];
      } else {
        const int q = r - n_w1 - n_b1 - n_wl;
        if (q < 2 * db) { const int t = q >> 1, cc = q & 1; v = bl[2 * (odd ? db - 1 - t : t) + cc]; }
      }
      blob[off_const + 4 + c] = v;
    }
  }
  // constant: sum over every elementwise affine of sum_p log alpha_p (block 0)
  if (blockIdx.x == 0) {
    __shared__ float red[32];
    float s = 0.f;
    const int n_tables = D.Lc + 3;
    for (int i = threadIdx.x; i < n_tables * d; i += blockDim.x) {
      const int tb = i / d, p = i % d;
      const int off = tb == 0 ? D.theta_aff0() : tb <= D.Lc ? D.theta_act(tb - 1) : tb == D.Lc + 1 ? D.theta_aff_tail() : D.theta_act_tail();
      s += logf(alpha_of(theta[off + 2 * p]));
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
      blob[off_const] = t; blob[off_const + 1] = 0.f; blob[off_const + 2] = 0.f; blob[off_const + 3] = 0.f;
    }
  }
}

// ---- blob gradient -> theta gradient ---------------------------------------------------------------------------------
__global__ void flow_grad_unpack_kernel(Dims D, const float* __restrict__ theta, const float* __restrict__ gblob, float scale,
                                        float* __restrict__ gtheta) {
  const int d = D.d, da = D.da, db = D.db, H = D.H;
  const int n_aff = (D.Lc + 1) * d;
  const int cpl_blob = flow_coupling_floats(d, 2, H);
  const int off_cpl = (D.Lc + 1) * 4 * d + 4;
  const long long total = (long long)n_aff + (long long)D.Lc * D.cpl_w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (i < n_aff) {
      const int g = (int)(i / d), k = (int)(i % d);
      const int p = (g & 1) ? d - 1 - k : k;
      int off[3];
      const int m = group_members(D, g, off);
      float a[3], b[3], al[4], be[4];
      al[0] = 1.f; be[0] = 0.f;
      for (int q = 0; q < m; ++q) {
        a[q] = alpha_of(theta[off[q] + 2 * p]);
        b[q] = 0.5f * theta[off[q] + 2 * p + 1];
        al[q + 1] = a[q] * al[q];
        be[q + 1] = fmaf(a[q], be[q], b[q]);
      }
      float Ga = gblob[(long long)g * 4 * d + 2 * k], Gb = gblob[(long long)g * 4 * d + 2 * k + 1];
      for (int q = m - 1; q >= 0; --q) {
        const float g_a = Ga * al[q] + Gb * be[q], g_b = Gb;
        gtheta[off[q] + 2 * p] = scale * g_a * (a[q] - kMinScale) * 0.5f;     // d alpha / d u_a
        gtheta[off[q] + 2 * p + 1] = scale * 0.5f * g_b;
        Ga *= a[q];
        Gb *= a[q];
      }
    } else {
      const long long c = i - n_aff;
      const int l = (int)(c / D.cpl_w), r = (int)(c % D.cpl_w);
      const bool odd = ((l + 1) & 1) != 0;
      const float* G = gblob + off_cpl + (long long)l * cpl_blob;
      const int gW1 = 0, gb1 = da * kSmallH, gWl = gb1 + kSmallH, gbl = gWl + db * 2 * kSmallH;
      float v;
      if (r < H * da) {
        const int h = r / da, s = r % da;
        v = G[gW1 + (odd ? da - 1 - s : s) * kSmallH + h];
      } else if (r < H * da + H) {
        v = G[gb1 + (r - H * da)];
      } else if (r < H * da + H + 2 * db * H) {
        const int q = r - H * da - H;
        const int row = q / H, h = q % H, t = row >> 1, cc = row & 1;
        v = G[gWl + ((odd ? db - 1 - t : t) * 2 + cc) * kSmallH + h];
      } else {
        const int q = r - H * da - H - 2 * db * H;
        const int t = q >> 1, cc = q & 1;
        v = G[gbl + 2 * (odd ? db - 1 - t : t) + cc];
      }
      gtheta[D.theta_cpl(l) + r] = scale * v;
    }
  }
}

// ---- AdamW (decoupled weight decay; the update of torch.optim.AdamW) -------------------------------------------------
__global__ void adamw_kernel(float* __restrict__ theta, const float* __restrict__ grad, float* __restrict__ m,
                             float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float wd,
                             float bc1, float bc2_sqrt) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = grad[i];
    float p = theta[i] * (1.f - lr * wd);
    const float mi = fmaf(b1, m[i], (1.f - b1) * g);
    const float vi = fmaf(b2, v[i], (1.f - b2) * g * g);
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p -= (lr / bc1) * (mi / denom);
    theta[i] = p;
  }
}

// x <- x - step * grad U(x)   (nfmc/dlmc.py:60-62)
template <int POT, int E>
__global__ void __launch_bounds__(kThreads) potential_step_kernel(PotParams P, float* __restrict__ x, long long n, int d, int gs, float step) {
  const Geom g = make_geom(d, gs);
  const int cpc = kThreads / gs;
  const long long tiles = (n + cpc - 1) / cpc;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long chain_raw = tile * cpc + threadIdx.x / gs;
    const bool active = chain_raw < n;
    const long long chain = active ? chain_raw : n - 1;
    float* row = x + chain * (long long)d;
    float lo[E], hi[E];
    load_chain(row, g, lo, hi);
    const PotCtx c = pot_prepare<POT, E>(P, g, lo, hi);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      float ul, uh;
      pot_grad<POT, false>(P, c, g, g.j + g.gs * e, lo[e], hi[e], ul, uh);
      lo[e] = __fsub_rn(lo[e], __fmul_rn(step, ul));
      hi[e] = __fsub_rn(hi[e], __fmul_rn(step, uh));
    }
    if (active) store_chain(row, g, lo, hi);
  }
}

// z <- z - step * (grad - z)   (nfmc/dlmc.py:84, elementwise over n*d)
__global__ void latent_update_kernel(float* __restrict__ z, const float* __restrict__ grad, float step, long long count) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    z[i] = __fsub_rn(z[i], __fmul_rn(step, __fsub_rn(grad[i], z[i])));
}

int check_train_shape(int d, int Lc, int M, int H) {
  if (d < 2 || d > NFMC_MAX_DIM || Lc < 0) return set_error("flow training: bad d / n_coupling");
  if (!flow_is_small(M, H) || H < 1)
    return set_error("flow training: the native path covers conditioners with 2 linear layers and <= 8 hidden units");
  return 0;
}

int launch_train(const nfmc_realnvp* flow, TrainArgs& A, int64_t n, float* grad, double* loss, int32_t accumulate, cudaStream_t s) {
  if (int e = validate_flow(flow)) return e;
  if (int e = check_train_shape(flow->d, flow->n_coupling, flow->n_linear, flow->hidden)) return e;
  if (!grad || n < 1) return set_error("flow training: bad grad / n");
  Layout L, W;
  if (!layout_for_dim(flow->d, L)) return set_error("flow training: unsupported event size");
  // a minibatch that fits one wave even at the widest layout is latency-bound: spread each row over more lanes
  if (layout_wide(flow->d, W) && (n * W.gs + kThreads - 1) / kThreads <= 2 * (int64_t)sm_count()) L = W;
  plan_flow_smem(A.f, flow, L, false);
  A.f.stage_blob = 0;
  A.grad = grad; A.loss = loss; A.n = n;
  if (!accumulate) {
    if (int e = check_cuda(cudaMemsetAsync(grad, 0, (size_t)flow->blob_floats * sizeof(float), s), "zero grad")) return e;
    if (loss) if (int e = check_cuda(cudaMemsetAsync(loss, 0, sizeof(double), s), "zero loss")) return e;
  }
  const int grid = grid_for(n, L.gs, 2);
  // CTA-local accumulation pays once a CTA sees several warps' worth of rows and the accumulator fits beside the L1
  const int64_t tiles = (n + kThreads / L.gs - 1) / (kThreads / L.gs);
  const bool shared_grad = tiles >= 4 * (int64_t)grid && (size_t)flow->blob_floats * sizeof(float) <= 96 * 1024;
  const size_t stash_b = (size_t)flow->n_coupling * (2 * L.E + kSmallH + 2) * kThreads * sizeof(float);
  A.stash = (stash_b + (shared_grad ? (size_t)flow->blob_floats * sizeof(float) : 0) <= 100 * 1024) ? 1 : 0;
  NFMC_DISPATCH_E(L.E, { return launch_flow_train<E>(A, grid, shared_grad, s); });
  return 0;
}

}  // namespace

extern "C" int64_t nfmc_flow_param_count(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden) {
  if (check_train_shape(d, n_coupling, n_linear, hidden)) return -1;
  return make_dims(d, n_coupling, hidden).n_theta;
}

extern "C" int nfmc_flow_pack(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden, const float* theta, float* blob,
                              void* stream) {
  if (int e = check_train_shape(d, n_coupling, n_linear, hidden)) return e;
  if (!theta || !blob) return set_error("flow_pack: NULL pointer");
  const Dims D = make_dims(d, n_coupling, hidden);
  const long long total = flow_blob_floats(d, n_coupling, 2, hidden);
  const int grid = (int)std::min<long long>((total + 255) / 256, 4 * sm_count());
  flow_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(D, theta, blob);
  return check_cuda(cudaGetLastError(), "flow_pack_kernel launch");
}

extern "C" int nfmc_flow_nll_grad(const nfmc_realnvp* flow, const float* x, const int64_t* rows, int64_t n, float* grad_blob,
                                  double* loss, int32_t accumulate, void* stream) {
  if (!x) return set_error("flow_nll_grad: x is NULL");
  TrainArgs A{};
  A.x = x; A.rows = reinterpret_cast<const long long*>(rows); A.kl = 0;
  return launch_train(flow, A, n, grad_blob, loss, accumulate, (cudaStream_t)stream);
}

extern "C" int nfmc_flow_kl_grad(const nfmc_potential* pot, const nfmc_realnvp* flow, const nfmc_rng* rng, int64_t chain0,
                                 int64_t n, float* grad_blob, double* loss, int32_t accumulate, void* stream) {
  if (int e = validate_pot(pot)) return e;
  if (!flow || pot->d != flow->d) return set_error("flow_kl_grad: potential and flow event sizes differ");
  if (!rng) return set_error("flow_kl_grad: rng is NULL");
  TrainArgs A{};
  A.kl = 1; A.pot_kind = pot->kind; A.pot = pot_params(pot);
  A.rng = RngArgs{rng->seed, rng->step0, rng->normals, nullptr};
  A.chain0 = chain0;
  return launch_train(flow, A, n, grad_blob, loss, accumulate, (cudaStream_t)stream);
}

extern "C" int nfmc_flow_grad_unpack(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden, const float* theta,
                                     const float* grad_blob, float scale, float* grad_theta, void* stream) {
  if (int e = check_train_shape(d, n_coupling, n_linear, hidden)) return e;
  if (!theta || !grad_blob || !grad_theta) return set_error("flow_grad_unpack: NULL pointer");
  const Dims D = make_dims(d, n_coupling, hidden);
  const int grid = (int)std::min<long long>((D.n_theta + 255) / 256, 4 * sm_count());
  flow_grad_unpack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(D, theta, grad_blob, scale, grad_theta);
  return check_cuda(cudaGetLastError(), "flow_grad_unpack_kernel launch");
}

extern "C" int nfmc_adamw_step(float* theta, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                               float beta1, float beta2, float eps, float weight_decay, int32_t step, void* stream) {
  if (!theta || !grad || !exp_avg || !exp_avg_sq || n < 1 || step < 1) return set_error("adamw_step: bad arguments");
  const float bc1 = 1.f - (float)pow((double)beta1, (double)step);
  const float bc2 = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  const int grid = (int)std::min<long long>((n + 255) / 256, 4 * sm_count());
  adamw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(theta, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                        weight_decay, bc1, bc2);
  return check_cuda(cudaGetLastError(), "adamw_kernel launch");
}

// One epoch of minibatch maximum-likelihood training on one GPU: for every batch of `perm`
//   pack(theta) -> loss/grad -> unpack(1/batch) -> AdamW.   losses[b] receives the summed loss of batch b.
extern "C" int nfmc_flow_fit_epoch(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden, float* theta,
                                   float* exp_avg, float* exp_avg_sq, float* blob, float* grad_blob, float* grad_theta,
                                   double* losses, const float* x, const int64_t* perm, int64_t n, int64_t batch_size,
                                   float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step0,
                                   void* stream) {
  if (int e = check_train_shape(d, n_coupling, n_linear, hidden)) return e;
  if (!perm || n < 1 || batch_size < 1) return set_error("flow_fit_epoch: bad perm / n / batch_size");
  nfmc_realnvp f;
  f.d = d; f.n_coupling = n_coupling; f.n_linear = n_linear; f.hidden = hidden;
  f.blob = blob; f.blob_floats = flow_blob_floats(d, n_coupling, n_linear, hidden);
  const int64_t P = make_dims(d, n_coupling, hidden).n_theta;
  int32_t step = step0;
  int64_t b = 0;
  for (int64_t i = 0; i < n; i += batch_size, ++b) {
    const int64_t m = std::min<int64_t>(batch_size, n - i);
    if (int e = nfmc_flow_pack(d, n_coupling, n_linear, hidden, theta, blob, stream)) return e;
    if (int e = nfmc_flow_nll_grad(&f, x, perm + i, m, grad_blob, losses ? losses + b : nullptr, 0, stream)) return e;
    if (int e = nfmc_flow_grad_unpack(d, n_coupling, n_linear, hidden, theta, grad_blob, 1.f / (float)m, grad_theta, stream)) return e;
    if (int e = nfmc_adamw_step(theta, grad_theta, exp_avg, exp_avg_sq, P, lr, beta1, beta2, eps, weight_decay, ++step, stream)) return e;
  }
  return 0;
}

// ---- deterministic Langevin Monte Carlo pieces (nfmc/dlmc.py:44-119) ------------------------------------------------
extern "C" int nfmc_potential_step(const nfmc_potential* pot, float* x, int64_t n, float step, void* stream) {
  if (int e = validate_pot(pot)) return e;
  if (!x || n < 1) return set_error("potential_step: bad arguments");
  Layout L;
  if (!layout_for_dim(pot->d, L)) return set_error("potential_step: unsupported event size");
  const PotParams P = pot_params(pot);
  cudaStream_t s = (cudaStream_t)stream;
  NFMC_DISPATCH_POT(pot->kind, NFMC_DISPATCH_E(L.E, {
    potential_step_kernel<POT, E><<<occupancy_grid(potential_step_kernel<POT, E>, 0, n, L.gs), kThreads, 0, s>>>(P, x, n, pot->d, L.gs, step);
  }));
  return check_cuda(cudaGetLastError(), "potential_step_kernel launch");
}

extern "C" int nfmc_dlmc_update(const nfmc_potential* pot, const nfmc_realnvp* flow, float* x, int64_t n, float step, void* stream) {
  if (int e = validate_pot(pot)) return e;
  if (int e = validate_flow(flow)) return e;
  if (int e = check_train_shape(flow->d, flow->n_coupling, flow->n_linear, flow->hidden)) return e;
  if (pot->d != flow->d) return set_error("dlmc_update: potential and flow event sizes differ");
  if (!x || n < 1) return set_error("dlmc_update: bad arguments");
  Layout L;
  if (!layout_for_dim(flow->d, L)) return set_error("dlmc_update: unsupported event size");
  TrainArgs A{};
  plan_flow_smem(A.f, flow, L, false);
  A.f.stage_blob = 0;
  A.n = n; A.x_rw = x; A.step = step; A.pot_kind = pot->kind; A.pot = pot_params(pot);
  const size_t stash_b = (size_t)flow->n_coupling * (2 * L.E + kSmallH + 2) * kThreads * sizeof(float);
  A.stash = stash_b <= 100 * 1024 ? 1 : 0;
  const int grid = grid_for(n, L.gs, 2);
  cudaStream_t s = (cudaStream_t)stream;
  NFMC_DISPATCH_E(L.E, { return launch_flow_dlmc<E>(A, grid, s); });
  return 0;
}

extern "C" int nfmc_dlmc_latent_update(float* z, const float* grad, float step, int64_t count, void* stream) {
  if (!z || !grad || count < 1) return set_error("dlmc_latent_update: bad arguments");
  const int grid = (int)std::min<long long>((count + 255) / 256, 8ll * sm_count());
  latent_update_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(z, grad, step, count);
  return check_cuda(cudaGetLastError(), "latent_update_kernel launch");
}
