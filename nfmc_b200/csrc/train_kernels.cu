// train_kernels.cu -- RealNVP training on the device (one translation unit per E): per-row loss and the gradient with
// respect to every flow parameter for the two objectives the reference trains flows with,
//   * maximum likelihood,  L = -sum_i log q(x_i)             -- Flow.fit            (jump.py:139-151,201; imh.py:171-175)
//   * reverse KL,          L =  sum_i [log q(x_i) + U(x_i)],  x_i = T^-1(z_i), z_i ~ N(0, I)
//                                                            -- Flow.variational_fit (imh.py:67-72; neutra.py:84-91)
// for the register-resident conditioner path (M = 2, H <= 8: every default conditioner).
//
// No activation is stored.  After the pass that produces the loss (x -> z for maximum likelihood, z -> x for reverse
// KL) the chain walks the layers back in the opposite order; a layer's input is re-derived from its output (the flow is
// invertible) while the cotangent is pulled through it and the layer's parameter gradients are emitted.  Both
// directions share one formulation: in pass direction every layer acts as  y = r*x + s  and contributes  -log r  to the
// loss, with (r, s) = (alpha, beta) going x -> z and (1/alpha, -beta/alpha) going z -> x; so
//   dL/dr = g_y*x - 1/r,  dL/ds = g_y,  g_x = g_y*r,
// converted to (dalpha, dbeta) and, for couplings, pushed through the conditioner MLP.
//
// Gradients are accumulated IN BLOB LAYOUT (the forward tables of the merged elementwise affines and the packed
// conditioner weights): lanes of a warp that own the same parameter are summed with shuffles, then one atomicAdd per
// warp and parameter.  train_param_kernels (api side) maps blob gradients back to module parameters.
#include "launchers.cuh"

#ifndef NFMC_ONLY_E
#error "compile with -DNFMC_ONLY_E=<slots per half>"
#endif

namespace nfmc {

// Gradient accumulator: the global blob-layout array, or (SG) a per-CTA copy in shared memory that is flushed once at
// the end of the kernel -- for large batches this turns one global atomic per warp and parameter into one per CTA.
template <bool SG>
struct GradSink {
  float* g;        // global accumulator
  unsigned sbase;  // shared-window address of the CTA-local accumulator (SG only)
  bool off;        // no parameter gradients wanted (input-gradient-only sweep): the emitters return at once
  __device__ __forceinline__ void add(int off, float v) const {
    if (SG) asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(sbase + 4u * (unsigned)off), "f"(v) : "memory");
    else atomicAdd(g + off, v);
  }
  // contiguous parameters (off even / a multiple of 4): one vector reduction to global memory (REDG.ADD.F32x2 / x4)
  __device__ __forceinline__ void add2(int off, float a, float b) const {
    if (SG) { add(off, a); add(off + 1, b); }
    else asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(g + off), "f"(a), "f"(b) : "memory");
  }
  __device__ __forceinline__ void add4(int off, float a, float b, float c, float d) const {
    if (SG) { add(off, a); add(off + 1, b); add(off + 2, c); add(off + 3, d); }
    else asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g + off), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
  }
};

// v = this lane's contribution to a parameter that is owned by lane position j (same parameter on lanes j, j+gs, ...)
template <bool SG>
__device__ __forceinline__ void emit(const GradSink<SG>& G, int off, float v, const Geom& g, bool ok) {
  if (G.off) return;
  const float s = across_groups_sum(v, g.gs);
  if (g.lane < g.gs && ok) G.add(off, s);
}
// v = a per-chain value (identical on the lanes of a group): counted once per group
template <bool SG>
__device__ __forceinline__ void emit_chain(const GradSink<SG>& G, int off, float v, const Geom& g) {
  if (G.off) return;
  const float s = across_groups_sum(g.j == 0 ? v : 0.f, g.gs);
  if (g.lane == 0) G.add(off, s);
}

template <bool SG>
__device__ __forceinline__ void emit2(const GradSink<SG>& G, int off, float a, float b, const Geom& g, bool ok) {
  if (G.off) return;
  const float sa = across_groups_sum(a, g.gs), sb = across_groups_sum(b, g.gs);
  if (g.lane < g.gs && ok) G.add2(off, sa, sb);
}
// eight contiguous parameters (one hidden-minor weight row)
template <bool SG>
__device__ __forceinline__ void emit8(const GradSink<SG>& G, int off, const float (&v)[kSmallH], float scale, const Geom& g, bool ok) {
  if (G.off) return;
  float s[kSmallH];
#pragma unroll
  for (int h = 0; h < kSmallH; ++h) s[h] = across_groups_sum(v[h] * scale, g.gs);
  if (g.lane < g.gs && ok) {
    G.add4(off, s[0], s[1], s[2], s[3]);
    G.add4(off + 4, s[4], s[5], s[6], s[7]);
  }
}
template <bool SG>
__device__ __forceinline__ void emit8_chain(const GradSink<SG>& G, int off, const float (&v)[kSmallH], float scale, const Geom& g) {
  if (G.off) return;
  float s[kSmallH];
#pragma unroll
  for (int h = 0; h < kSmallH; ++h) s[h] = across_groups_sum(g.j == 0 ? v[h] * scale : 0.f, g.gs);
  if (g.lane == 0) {
    G.add4(off, s[0], s[1], s[2], s[3]);
    G.add4(off + 4, s[4], s[5], s[6], s[7]);
  }
}

// (y, gy) = output of the layer and dL/dy  ->  (x, gx); returns (dalpha, dbeta)
__device__ __forceinline__ float2 affine_back(bool inv, float act, float alpha, float beta, float ralpha, float& y, float& gy) {
  const float r = inv ? ralpha : alpha;
  const float xin = inv ? fmaf(alpha, y, beta) : (y - beta) * ralpha;
  const float dr = gy * xin - act * (inv ? alpha : ralpha);
  const float ds = gy;
  const float dal = inv ? (ds * beta - dr) * ralpha * ralpha : dr;
  const float dbe = inv ? -ds * ralpha : ds;
  y = xin;
  gy *= r;
  return make_float2(dal, dbe);
}

template <int E, bool SG>
__device__ __forceinline__ void affine_train(const FlowDesc& F, const Geom& g, int a, bool inv, float act, float (&lo)[E],
                                             float (&hi)[E], float (&glo)[E], float (&ghi)[E], const GradSink<SG>& G) {
  const int fw = a * 4 * F.d, iv = fw + 2 * F.d;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int k = g.j + g.gs * e;
    const bool vl = k < g.da, vh = k < g.db;
    const int il = 2 * (vl ? k : 0), ih = 2 * (g.da + (vh ? k : 0));
    const float2 pl = ldp2<false>(F, fw + il), ph = ldp2<false>(F, fw + ih);
    const float rl = ldp<false>(F, iv + il), rh = ldp<false>(F, iv + ih);
    float2 dl = affine_back(inv, act, pl.x, pl.y, rl, lo[e], glo[e]);
    float2 dh = affine_back(inv, act, ph.x, ph.y, rh, hi[e], ghi[e]);
    if (!vl) { lo[e] = 0.f; glo[e] = 0.f; dl = make_float2(0.f, 0.f); }
    if (!vh) { hi[e] = 0.f; ghi[e] = 0.f; dh = make_float2(0.f, 0.f); }
    emit2(G, fw + il, dl.x, dl.y, g, vl);
    emit2(G, fw + ih, dh.x, dh.y, g, vh);
  }
}

// conditioner backward with weight gradients (small path); dsrc[e] += input-VJP
template <int E, bool SG>
__device__ __forceinline__ void cond_backward_small_train(const FlowDesc& F, const Geom& g, int W, int shift, int nt_main,
                                                          bool has_x, const float (&hid)[kSmallH], const float (&src)[E],
                                                          const float (&dua)[E], const float (&dub)[E], float dua_x,
                                                          float dub_x, float (&dsrc)[E], const GradSink<SG>& G) {
  const int da = F.da, db = F.db;
  const int b1 = W + da * kSmallH;
  const int Wl = b1 + kSmallH;
  const int bl = Wl + db * 2 * kSmallH;
  float acc[kSmallH];
#pragma unroll
  for (int h = 0; h < kSmallH; ++h) acc[h] = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int t = g.j + g.gs * e;
    const bool ok = t < nt_main;
    const float va = ok ? dua[e] : 0.f, vb = ok ? dub[e] : 0.f;
    const int w = Wl + (ok ? t : 0) * 2 * kSmallH;
#pragma unroll
    for (int h = 0; h < kSmallH; ++h)
      acc[h] = fmaf(ldp<false>(F, w + h), va, fmaf(ldp<false>(F, w + kSmallH + h), vb, acc[h]));
    emit8(G, w, hid, va, g, ok);                     // d/dWl[t][c][h] = hid[h] * d out[t][c]  (padded h: hid = 0)
    emit8(G, w + kSmallH, hid, vb, g, ok);
    emit2(G, bl + 2 * (ok ? t : 0), va, vb, g, ok);
  }
  if (has_x) {                                       // the extra target t = da lives on lane j = 0 of each group
    const int w = Wl + da * 2 * kSmallH;
    const float va = g.j == 0 ? dua_x : 0.f, vb = g.j == 0 ? dub_x : 0.f;
#pragma unroll
    for (int h = 0; h < kSmallH; ++h)
      acc[h] = fmaf(ldp<false>(F, w + h), va, fmaf(ldp<false>(F, w + kSmallH + h), vb, acc[h]));
    emit8_chain(G, w, hid, dua_x, g);
    emit8_chain(G, w + kSmallH, hid, dub_x, g);
    emit_chain(G, bl + 2 * da, dua_x, g);
    emit_chain(G, bl + 2 * da + 1, dub_x, g);
  }
  float dpre[kSmallH];
#pragma unroll
  for (int h = 0; h < kSmallH; ++h) dpre[h] = group_sum(acc[h], g.gs) * (1.f - hid[h] * hid[h]);
  emit8_chain(G, b1, dpre, 1.f, g);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int ks = g.j + g.gs * e - shift;
    const bool ok = ks >= 0 && ks < da;
    const int w = W + (ok ? ks : 0) * kSmallH;
    const float v = ok ? src[e] : 0.f;
    float s = 0.f;
#pragma unroll
    for (int h = 0; h < kSmallH; ++h) s = fmaf(ldp<false>(F, w + h), dpre[h], s);
    emit8(G, w, dpre, v, g, ok);                      // d/dW1[h][ks] = src[ks] * dpre[h]
    if (ok) dsrc[e] += s;
  }
}

template <int E, bool SG>
__device__ __forceinline__ void coupling_train(const FlowDesc& F, const Geom& g, int l, bool inv, float act, float (&lo)[E],
                                               float (&hi)[E], float (&glo)[E], float (&ghi)[E], const GradSink<SG>& G,
                                               const float* stash) {
  float ua[E], ub[E], ua_x = 0.f, ub_x = 0.f, dua_x = 0.f, dub_x = 0.f;
  float hid[kSmallH];
  const bool src_is_hi = (l & 1) == 0;
  const int shift = src_is_hi ? F.db - F.da : 0;
  const bool has_x = shift > 0;
  const int nt_main = src_is_hi ? F.da : F.db;
  const int Woff = F.off_coupling + l * F.coupling_stride;
  swap_halves(lo, hi, src_is_hi);
  swap_halves(glo, ghi, src_is_hi);
  if (stash) {                                       // conditioner outputs kept by the pass that produced the loss
#pragma unroll
    for (int e = 0; e < E; ++e) { ua[e] = stash[e * kThreads]; ub[e] = stash[(E + e) * kThreads]; }
#pragma unroll
    for (int h = 0; h < kSmallH; ++h) hid[h] = stash[(2 * E + h) * kThreads];
    ua_x = stash[(2 * E + kSmallH) * kThreads];
    ub_x = stash[(2 * E + kSmallH + 1) * kThreads];
  } else {
    cond_forward_small<E, false, false>(F, g, Woff, lo, shift, nt_main, has_x, hid, ua, ub, ua_x, ub_x);
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    float d_ua = 0.f, d_ub = 0.f;
    if (g.j + g.gs * e < nt_main) {
      float al, be;
      affine_coef(ua[e], ub[e], al, be);
      const float2 dd = affine_back(inv, act, al, be, __fdividef(1.f, al), hi[e], ghi[e]);
      d_ua = dd.x * (al - kMinScale) * 0.5f;
      d_ub = 0.5f * dd.y;
    }
    ua[e] = d_ua;
    ub[e] = d_ub;
  }
  if (has_x && g.j == 0) {
    float al, be;
    affine_coef(ua_x, ub_x, al, be);
    const float2 dd = affine_back(inv, act, al, be, __fdividef(1.f, al), lo[0], glo[0]);
    dua_x = dd.x * (al - kMinScale) * 0.5f;
    dub_x = 0.5f * dd.y;
  }
  cond_backward_small_train<E, SG>(F, g, Woff, shift, nt_main, has_x, hid, lo, ua, ub, dua_x, dub_x, glo, G);
  swap_halves(lo, hi, src_is_hi);
  swap_halves(glo, ghi, src_is_hi);
}

// walk the layers back: `inv` = direction of the pass that produced (lo, hi)
template <int E, bool SG>
__device__ __forceinline__ void flow_train_sweep(const FlowDesc& F, const Geom& g, bool inv, float act, float (&lo)[E],
                                                 float (&hi)[E], float (&glo)[E], float (&ghi)[E], const GradSink<SG>& G,
                                                 const float* stash) {
  const int n_ops = 2 * F.Lc + 1;
#pragma unroll 1
  for (int i = 0; i < n_ops; ++i) {
    const int op = inv ? i : n_ops - 1 - i;
    if (op & 1)
      coupling_train<E, SG>(F, g, op >> 1, inv, act, lo, hi, glo, ghi, G,
                            stash ? stash + (size_t)(op >> 1) * cond_stash_floats<E>() * kThreads : nullptr);
    else affine_train<E, SG>(F, g, op >> 1, inv, act, lo, hi, glo, ghi, G);
  }
}

template <int E, bool SG>
__global__ void __launch_bounds__(kThreads, 2) flow_train_kernel(const TrainArgs A) {
  extern __shared__ __align__(16) float sgrad[];
  GradSink<SG> G;
  G.g = A.grad;
  G.sbase = 0u;
  G.off = false;
  if (SG) {
    for (int i = threadIdx.x; i < (int)A.f.blob_floats; i += blockDim.x) sgrad[i] = 0.f;
    G.sbase = (unsigned)__cvta_generic_to_shared(sgrad);
    __syncthreads();
  }
  const Geom g = make_geom(A.f.d, A.f.gs);
  const FlowDesc F = make_flow_desc(A.f.blob, A.f.d, A.f.Lc, A.f.M, A.f.H);
  const int cpc = kThreads / A.f.gs;
  const long long tiles = (A.n + cpc - 1) / cpc;
  const bool flip = (A.f.Lc & 1) != 0;
  // conditioner stash (flow.cuh) behind the optional shared accumulator
  float* stash = A.stash ? sgrad + (SG ? (((int)A.f.blob_floats + 3) & ~3) : 0) + threadIdx.x : nullptr;
  double loss = 0.0;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long chain_raw = tile * cpc + threadIdx.x / A.f.gs;
    const bool active = chain_raw < A.n;
    const long long chain = active ? chain_raw : A.n - 1;
    const float act = active ? 1.f : 0.f;
    float lo[E], hi[E], glo[E], ghi[E];
    float li;
    if (!A.kl) {
      const long long row = A.rows ? A.rows[chain] : chain;
      load_chain(A.x + row * (long long)A.f.d, g, lo, hi);
      const float ld = flow_pass<E, false, false, true>(F, g, false, lo, hi, nullptr, stash);
      li = -(base_log_prob(g, lo, hi) + ld);                                   // -log q(x)
#pragma unroll
      for (int e = 0; e < E; ++e) { glo[e] = act * lo[e]; ghi[e] = act * hi[e]; }   // d/dz of |z|^2 / 2
      flow_train_sweep<E, SG>(F, g, false, act, lo, hi, glo, ghi, G, stash);
    } else {
      draw_base(A.rng, g, flip, A.n, chain, A.chain0, 0, lo, hi);
      const float lbase = base_log_prob(g, lo, hi);
      const float ld_inv = flow_pass<E, false, false, true>(F, g, true, lo, hi, nullptr, stash);
      const PotCtx c = pot_prepare_rt<E>(A.pot_kind, A.pot, g, lo, hi);
      li = lbase - ld_inv + c.u;                                                // log q(x) + U(x)
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int kk = g.j + g.gs * e;
        pot_grad_rt(A.pot_kind, A.pot, c, g, kk, lo[e], hi[e], glo[e], ghi[e]);
        glo[e] = kk < g.da ? act * glo[e] : 0.f;
        ghi[e] = kk < g.db ? act * ghi[e] : 0.f;
      }
      flow_train_sweep<E, SG>(F, g, true, act, lo, hi, glo, ghi, G, stash);
    }
    if (active && g.j == 0) loss += (double)li;
  }
  for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
  if ((threadIdx.x & 31) == 0 && A.loss) atomicAdd(A.loss, loss);
  if (SG) {
    __syncthreads();
    for (int i = threadIdx.x; i < (int)A.f.blob_floats; i += blockDim.x) {
      const float v = sgrad[i];
      if (v != 0.f) atomicAdd(A.grad + i, v);
    }
  }
}

// One deterministic-Langevin update of every particle (nfmc/dlmc.py:86-88): x <- x - step * grad_x [ U(x) + log q(x) ].
// grad_x log q comes from the same reversible sweep as the training gradient with the parameter emitters switched off
// (the cotangent that reaches the input is d(-log q)/dx); grad U is the closed form.
template <int E>
__global__ void __launch_bounds__(kThreads, 2) flow_dlmc_kernel(const TrainArgs A) {
  extern __shared__ __align__(16) float sgrad[];
  GradSink<false> G;
  G.g = nullptr; G.sbase = 0u; G.off = true;
  const Geom g = make_geom(A.f.d, A.f.gs);
  const FlowDesc F = make_flow_desc(A.f.blob, A.f.d, A.f.Lc, A.f.M, A.f.H);
  const int cpc = kThreads / A.f.gs;
  const long long tiles = (A.n + cpc - 1) / cpc;
  float* stash = A.stash ? sgrad + threadIdx.x : nullptr;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long chain_raw = tile * cpc + threadIdx.x / A.f.gs;
    const bool active = chain_raw < A.n;
    const long long chain = active ? chain_raw : A.n - 1;
    float* row = A.x_rw + chain * (long long)A.f.d;
    float lo[E], hi[E], glo[E], ghi[E];
    load_chain(row, g, lo, hi);
    flow_pass<E, false, false, true>(F, g, false, lo, hi, nullptr, stash);          // z
#pragma unroll
    for (int e = 0; e < E; ++e) { glo[e] = lo[e]; ghi[e] = hi[e]; }                    // d(-log q)/dz = z
    flow_train_sweep<E, false>(F, g, false, 1.f, lo, hi, glo, ghi, G, stash);          // -> d(-log q)/dx
    load_chain(row, g, lo, hi);                                                        // the exact x, not the re-derived one
    const PotCtx c = pot_prepare_rt<E>(A.pot_kind, A.pot, g, lo, hi);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int kk = g.j + g.gs * e;
      float ul, uh;
      pot_grad_rt(A.pot_kind, A.pot, c, g, kk, lo[e], hi[e], ul, uh);
      lo[e] = __fsub_rn(lo[e], __fmul_rn(A.step, __fsub_rn(ul, glo[e])));            // grad = grad U + grad log q
      hi[e] = __fsub_rn(hi[e], __fmul_rn(A.step, __fsub_rn(uh, ghi[e])));
    }
    if (active) store_chain(row, g, lo, hi);
  }
}

// shared_grad: accumulate per CTA in shared memory (smem = blob_floats * 4 bytes); chosen by the caller for large batches
template <int E>
int launch_flow_train(const TrainArgs& A, int grid, bool shared_grad, cudaStream_t s) {
  const size_t stash_b = A.stash ? (size_t)A.f.Lc * cond_stash_floats<E>() * kThreads * sizeof(float) : 0;
  if (shared_grad) {
    const size_t smem = (size_t)((A.f.blob_floats + 3) & ~3ll) * sizeof(float) + stash_b;
    NFMC_SET_SMEM_RET((flow_train_kernel<E, true>), smem);
    flow_train_kernel<E, true><<<occupancy_grid(flow_train_kernel<E, true>, smem, A.n, A.f.gs), kThreads, smem, s>>>(A);
  } else {
    NFMC_SET_SMEM_RET((flow_train_kernel<E, false>), stash_b);
    flow_train_kernel<E, false><<<occupancy_grid(flow_train_kernel<E, false>, stash_b, A.n, A.f.gs), kThreads, stash_b, s>>>(A);
  }
  return check_cuda(cudaGetLastError(), "flow_train_kernel launch");
}
template int launch_flow_train<NFMC_ONLY_E>(const TrainArgs&, int, bool, cudaStream_t);

template <int E>
int launch_flow_dlmc(const TrainArgs& A, int grid, cudaStream_t s) {
  const size_t stash_b = A.stash ? (size_t)A.f.Lc * cond_stash_floats<E>() * kThreads * sizeof(float) : 0;
  NFMC_SET_SMEM_RET((flow_dlmc_kernel<E>), stash_b);
  flow_dlmc_kernel<E><<<occupancy_grid(flow_dlmc_kernel<E>, stash_b, A.n, A.f.gs), kThreads, stash_b, s>>>(A);
  return check_cuda(cudaGetLastError(), "flow_dlmc_kernel launch");
}
template int launch_flow_dlmc<NFMC_ONLY_E>(const TrainArgs&, int, cudaStream_t);

}  // namespace nfmc
