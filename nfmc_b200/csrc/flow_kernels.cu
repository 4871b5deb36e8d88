// flow_kernels.cu -- RealNVP passes, flow sampling, the NF jump and independence-MH, one chain per lane group
// (one translation unit per E).
//
// Replaces (reference paths under /root/reference/nfmc/algorithms/sampling/):
//   flow.bijection.forward / inverse          nfmc/neutra.py:60,122
//   flow.log_prob                             nfmc/jump.py:218, nfmc/imh.py:133-134,214
//   flow.sample(n, return_log_prob=True)      nfmc/jump.py:205, nfmc/imh.py:221
//   the jump block of JumpNFMC.sample         nfmc/jump.py:203-243
//   FixedIMH.sample / AdaptiveIMH.sample      nfmc/imh.py:214-249, 122-150
#include "launchers.cuh"

#ifndef NFMC_ONLY_E
#error "compile with -DNFMC_ONLY_E=<slots per half>"
#endif

#ifndef NFMC_JUMP_MINB
#define NFMC_JUMP_MINB 3
#endif
#ifndef NFMC_FLOW_MINB
#define NFMC_FLOW_MINB 3      // resident CTAs per SM the flow-pass / propose-accept kernels are compiled for
#endif

namespace nfmc {

template <int E, bool SB, bool X, bool SM>
__global__ void __launch_bounds__(kThreads, NFMC_FLOW_MINB) flow_pass_kernel(FlowArgs A, int mode, const float* __restrict__ in,
                                                            float* __restrict__ out, float* __restrict__ aux, long long n) {
  extern __shared__ __align__(16) unsigned char smem[];
  const Geom g = make_geom(A.d, A.gs);
  FlowSmem S = flow_smem_init<SB>(smem, A, false);
  const bool flip = (A.Lc & 1) != 0;
  const int cpc = kThreads / A.gs;
  const long long tiles = (n + cpc - 1) / cpc;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long chain_raw = tile * cpc + threadIdx.x / A.gs;
    const bool active = chain_raw < n;
    const long long chain = active ? chain_raw : n - 1;
    float lo[E], hi[E];
    const float* src = in + chain * (long long)A.d;
    if (mode == PASS_INVERSE && flip) load_chain_flipped(src, g, lo, hi);
    else load_chain(src, g, lo, hi);
    float r = flow_pass<E, SB, X, SM>(S.F, g, mode == PASS_INVERSE, lo, hi, S.scr);
    if (mode == PASS_LOGPROB) r += base_log_prob(g, lo, hi);
    if (active) {
      if (out) {
        float* dst = out + chain * (long long)A.d;
        if (mode != PASS_INVERSE && flip) store_chain_flipped(dst, g, lo, hi);
        else store_chain(dst, g, lo, hi);
      }
      if (aux && g.j == 0) aux[chain] = r;
    }
  }
}

template <int E, bool SB, bool X, bool SM>
__global__ void __launch_bounds__(kThreads, 3) flow_sample_kernel(FlowArgs A, RngArgs R, long long chain0, float* __restrict__ x,
                                                              float* __restrict__ logq, long long n) {
  extern __shared__ __align__(16) unsigned char smem[];
  const Geom g = make_geom(A.d, A.gs);
  FlowSmem S = flow_smem_init<SB>(smem, A, false);
  const bool flip = (A.Lc & 1) != 0;
  const int cpc = kThreads / A.gs;
  const long long tiles = (n + cpc - 1) / cpc;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long chain_raw = tile * cpc + threadIdx.x / A.gs;
    const bool active = chain_raw < n;
    const long long chain = active ? chain_raw : n - 1;
    float lo[E], hi[E];
    draw_base<E>(R, g, flip, n, chain, chain0, 0, lo, hi);
    const float blp = base_log_prob(g, lo, hi);
    const float ld = flow_inverse<E, SB, X, SM>(S.F, g, lo, hi, S.scr);
    if (active) {
      store_chain(x + chain * (long long)A.d, g, lo, hi);
      if (logq && g.j == 0) logq[chain] = blp - ld;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// NF jump / IMH.  IMH = n_steps iterations with the state, U(x) and log q(x) kept on chip; the jump is the
// same kernel with n_steps = 1 and log q(x) computed from x (jump.py:218).
// ---------------------------------------------------------------------------------------------------------


template <int E, bool SB, bool X, bool SM>
__global__ void __launch_bounds__(kThreads, NFMC_JUMP_MINB) jump_kernel(const JumpArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ChainArgs& C = A.c;
  const Geom g = make_geom(C.d, C.gs);
  FlowSmem S = flow_smem_init<SB>(smem, A.f, true);
  const bool flip = (A.f.Lc & 1) != 0;
  const int cpc = kThreads / C.gs;
  const long long tiles = (C.n + cpc - 1) / cpc;
  unsigned int n_acc = 0, n_bad = 0;

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long chain_raw = tile * cpc + threadIdx.x / C.gs;
    const bool active = chain_raw < C.n;
    const long long chain = active ? chain_raw : C.n - 1;
    float* row = C.x + chain * (long long)C.d;

    float lo[E], hi[E];
    load_chain(row, g, lo, hi);
    const bool multi = C.n_steps > 1;   // one step (NF jump): the moments are flushed straight from the registers
    if (multi) {
#pragma unroll
      for (int e = 0; e < E; ++e) S.mom[e * kThreads] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float u_x = 0.f, f_x = 0.f;
    if (A.adjusted) {
      u_x = pot_prepare_rt<E>(A.pot_kind, C.pot, g, lo, hi).u;                                   // jump.py:212
      if (!A.recompute_logq && A.logq_x) f_x = __ldg(A.logq_x + chain);                 // imh.py:214
    }
    for (int k = 0; k < C.n_steps; ++k) {
      // pass 0: log q(x) by a forward pass on a copy of x (jump.py:218 / imh.py:133), skipped when it is cached;
      // pass 1: x', log q(x') = flow.sample(n, return_log_prob=True) (jump.py:205, imh.py:221).  One flow_pass call site.
      float plo[E], phi[E];
      uint32_t ubits = 0;
      float f_p = 0.f;
      const bool need_fx = A.adjusted && (A.recompute_logq || !A.logq_x);
#pragma unroll 1
      for (int pass = need_fx ? 0 : 1; pass < 2; ++pass) {
        float blp = 0.f;
        if (pass == 0) {
#pragma unroll
          for (int e = 0; e < E; ++e) { plo[e] = lo[e]; phi[e] = hi[e]; }
        } else {
          ubits = draw_base<E>(C.rng, g, flip, C.n, chain, C.chain0, k, plo, phi);
          blp = base_log_prob(g, plo, phi);
        }
        const float ld = flow_pass<E, SB, X, SM>(S.F, g, pass == 1, plo, phi, S.scr);
        if (pass == 0) f_x = base_log_prob(g, plo, phi) + ld;
        else f_p = blp - ld;
      }
      bool accept = true;
      float u_p = 0.f;
      if (A.adjusted) {
        u_p = pot_prepare_rt<E>(A.pot_kind, C.pot, g, plo, phi).u;                               // jump.py:213
        const float log_alpha = (-u_p) - (-u_x) + f_x - f_p;                           // jump.py:219-224, util.py:392
        float u;
        if (C.rng.uniforms) u = __ldg(C.rng.uniforms + (long long)k * C.n + chain);
        else u = uniform_from_bits(__shfl_sync(0xffffffffu, ubits, g.grp_base));
        accept = logf(u) < log_alpha;                                                  // jump.py:225
        if (!(fabsf(log_alpha) <= 3.0e38f) && g.j == 0 && active) ++n_bad;
      }
#pragma unroll
      for (int e = 0; e < E; ++e) {                                                    // jump.py:231, imh.py:232
        lo[e] = accept ? plo[e] : lo[e];
        hi[e] = accept ? phi[e] : hi[e];
      }
      u_x = accept ? u_p : u_x;
      f_x = accept ? f_p : f_x;                                                        // imh.py:233
      if (accept && g.j == 0 && active) ++n_acc;
      if (multi) {                                                                     // jump.py:240, imh.py:242
#pragma unroll
        for (int e = 0; e < E; ++e) {
          float4 m = S.mom[e * kThreads];
          m.x += lo[e]; m.y += hi[e]; m.z = fmaf(lo[e], lo[e], m.z); m.w = fmaf(hi[e], hi[e], m.w);
          S.mom[e * kThreads] = m;
        }
      }
      if (C.sink.samples && active) sink_store(C.sink, g, C.n, chain, k, lo, hi);      // jump.py:243, imh.py:249
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      float4 m = multi ? S.mom[e * kThreads] : make_float4(lo[e], hi[e], lo[e] * lo[e], hi[e] * hi[e]);
      if (!active) m = make_float4(0.f, 0.f, 0.f, 0.f);
      const int kk = g.j + g.gs * e;
      const float a = across_groups_sum(m.x, g.gs), b = across_groups_sum(m.y, g.gs);
      const float c = across_groups_sum(m.z, g.gs), dd = across_groups_sum(m.w, g.gs);
      if (g.lane < g.gs) {
        if (kk < g.da) { atomicAdd(S.st.sx + kk, (double)a); atomicAdd(S.st.sx2 + kk, (double)c); }
        if (kk < g.db) { atomicAdd(S.st.sx + g.da + kk, (double)b); atomicAdd(S.st.sx2 + g.da + kk, (double)dd); }
      }
    }
    if (active) {
      store_chain(row, g, lo, hi);
      if (A.logq_x && g.j == 0) A.logq_x[chain] = f_x;
    }
  }
  n_acc = __reduce_add_sync(0xffffffffu, n_acc);
  n_bad = __reduce_add_sync(0xffffffffu, n_bad);
  if ((threadIdx.x & 31) == 0) {
    if (n_acc) atomicAdd(S.st.cnt + 0, (unsigned long long)n_acc);
    if (n_bad) atomicAdd(S.st.cnt + 2, (unsigned long long)n_bad);
  }
  if (threadIdx.x == 0) {
    long long mine = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const long long first = tile * cpc;
      mine += (C.n - first) < cpc ? (C.n - first) : cpc;
    }
    atomicAdd(S.st.cnt + 1, (unsigned long long)(mine * C.n_steps));
  }
  cta_stats_finish(S.st, C.stats, C.d);
}


// ---------------------------------------------------------------------------------------------------------
// Second half of a two-kernel NF jump (jump.py:203-243): log q(x) has been written to A.logq_x by a forward pass
// (flow_pass_kernel, PASS_LOGPROB); this kernel draws z, runs the inverse pass to x' with log q(x'), and only THEN loads
// x for U(x), the accept test, the overwrite and the moments.  One state vector is live during the flow pass, so the
// kernel carries no spills -- the fused jump_kernel keeps x, its working copy and the proposal and spills at 168 registers;
// split in two the jump takes 2.0 ms instead of 3.1 ms (2^20 chains, d = 100).  Rejected chains are not written back.
// ---------------------------------------------------------------------------------------------------------
template <int E, bool SB, bool X, bool SM>
__global__ void __launch_bounds__(kThreads, NFMC_FLOW_MINB) jump_propose_accept_kernel(const JumpArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ChainArgs& C = A.c;
  const Geom g = make_geom(C.d, C.gs);
  FlowSmem S = flow_smem_init<SB>(smem, A.f, true);
  const bool flip = (A.f.Lc & 1) != 0;
  const int cpc = kThreads / C.gs;
  const long long tiles = (C.n + cpc - 1) / cpc;
  unsigned int n_acc = 0, n_bad = 0;
  // running moments: per-lane fp32 sums in shared memory over all of this CTA's tiles, reduced across the warp once at the end
#pragma unroll
  for (int e = 0; e < E; ++e) S.mom[e * kThreads] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long chain_raw = tile * cpc + threadIdx.x / C.gs;
    const bool active = chain_raw < C.n;
    const long long chain = active ? chain_raw : C.n - 1;
    float* row = C.x + chain * (long long)C.d;

    float plo[E], phi[E];
    const uint32_t ubits = draw_base<E>(C.rng, g, flip, C.n, chain, C.chain0, 0, plo, phi);    // jump.py:205
    const float blp = base_log_prob(g, plo, phi);
    const float ld = flow_pass<E, SB, X, SM>(S.F, g, true, plo, phi, S.scr);
    const float f_p = blp - ld;

    float lo[E], hi[E];
    load_chain(row, g, lo, hi);
    bool accept = true;
    if (A.adjusted) {
      const float u_x = pot_prepare_rt<E>(A.pot_kind, C.pot, g, lo, hi).u;                      // jump.py:212
      const float u_p = pot_prepare_rt<E>(A.pot_kind, C.pot, g, plo, phi).u;                    // jump.py:213
      const float f_x = __ldg(A.logq_x + chain);                                                // jump.py:218 (first kernel)
      const float log_alpha = (-u_p) - (-u_x) + f_x - f_p;                                      // jump.py:219-224, util.py:392
      float u;
      if (C.rng.uniforms) u = __ldg(C.rng.uniforms + chain);
      else u = uniform_from_bits(__shfl_sync(0xffffffffu, ubits, g.grp_base));
      accept = logf(u) < log_alpha;                                                             // jump.py:225
      if (!(fabsf(log_alpha) <= 3.0e38f) && g.j == 0 && active) ++n_bad;
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {                                                               // jump.py:231
      lo[e] = accept ? plo[e] : lo[e];
      hi[e] = accept ? phi[e] : hi[e];
    }
    if (accept && g.j == 0 && active) {
      ++n_acc;
      if (A.logq_x && A.adjusted) A.logq_x[chain] = f_p;                                        // imh.py:233 (the log q(x) cache)
    }
    if (C.sink.samples && active) sink_store(C.sink, g, C.n, chain, 0, lo, hi);                 // jump.py:243
    if (active) {
#pragma unroll
      for (int e = 0; e < E; ++e) {                                                             // jump.py:240
        float4 m = S.mom[e * kThreads];
        m.x += lo[e]; m.y += hi[e]; m.z = fmaf(lo[e], lo[e], m.z); m.w = fmaf(hi[e], hi[e], m.w);
        S.mom[e * kThreads] = m;
      }
    }
    if (active && accept) store_chain(row, g, lo, hi);
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const float4 m = S.mom[e * kThreads];
    const int kk = g.j + g.gs * e;
    const float a = across_groups_sum(m.x, g.gs), b = across_groups_sum(m.y, g.gs);
    const float c = across_groups_sum(m.z, g.gs), dd = across_groups_sum(m.w, g.gs);
    if (g.lane < g.gs) {
      if (kk < g.da) { atomicAdd(S.st.sx + kk, (double)a); atomicAdd(S.st.sx2 + kk, (double)c); }
      if (kk < g.db) { atomicAdd(S.st.sx + g.da + kk, (double)b); atomicAdd(S.st.sx2 + g.da + kk, (double)dd); }
    }
  }
  n_acc = __reduce_add_sync(0xffffffffu, n_acc);
  n_bad = __reduce_add_sync(0xffffffffu, n_bad);
  if ((threadIdx.x & 31) == 0) {
    if (n_acc) atomicAdd(S.st.cnt + 0, (unsigned long long)n_acc);
    if (n_bad) atomicAdd(S.st.cnt + 2, (unsigned long long)n_bad);
  }
  if (threadIdx.x == 0) {
    long long mine = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const long long first = tile * cpc;
      mine += (C.n - first) < cpc ? (C.n - first) : cpc;
    }
    atomicAdd(S.st.cnt + 1, (unsigned long long)mine);
  }
  cta_stats_finish(S.st, C.stats, C.d);
}


// ---------------------------------------------------------------------------------------------------------
// Accept step of a jump whose flow passes ran elsewhere (tensor-core path, cond_tc.cu): given x, the proposal
// x' = T^-1(z) with log|det dx'/dz|, the base draw z and log q(x), do jump.py:212-231 / imh.py:223-233:
// U(x), U(x'), log alpha, accept, overwrite, moments, counters.  HBM-bound (3 rows in, 1 row out).
// ---------------------------------------------------------------------------------------------------------
template <int E>
__global__ void __launch_bounds__(kThreads) jump_accept_kernel(const AcceptArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ChainArgs& C = A.c;
  const Geom g = make_geom(C.d, C.gs);
  CtaStats st = cta_stats_init(smem, C.d);
  const int cpc = kThreads / C.gs;
  const long long tiles = (C.n + cpc - 1) / cpc;
  unsigned int n_acc = 0, n_bad = 0;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long chain_raw = tile * cpc + threadIdx.x / C.gs;
    const bool active = chain_raw < C.n;
    const long long chain = active ? chain_raw : C.n - 1;
    float* row = C.x + chain * (long long)C.d;
    float lo[E], hi[E], plo[E], phi[E];
    load_chain(row, g, lo, hi);
    load_chain(A.x_prime + chain * (long long)C.d, g, plo, phi);
    bool accept = true;
    float f_p = 0.f;
    if (A.adjusted) {
      float zlo[E], zhi[E];
      load_chain(A.z + chain * (long long)C.d, g, zlo, zhi);       // N(z) is permutation invariant: no flip needed
      f_p = base_log_prob(g, zlo, zhi) - __ldg(A.ld_inv + chain);   // log q(x')
      const float f_x = __ldg(A.logq_x + chain);
      const float u_x = pot_prepare_rt<E>(A.pot_kind, C.pot, g, lo, hi).u;
      const float u_p = pot_prepare_rt<E>(A.pot_kind, C.pot, g, plo, phi).u;
      const float log_alpha = (-u_p) - (-u_x) + f_x - f_p;          // jump.py:219-224, util.py:392
      const float u = __ldg(A.uniforms + chain);
      accept = logf(u) < log_alpha;                                 // jump.py:225
      if (!(fabsf(log_alpha) <= 3.0e38f) && g.j == 0 && active) ++n_bad;
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      lo[e] = accept ? plo[e] : lo[e];
      hi[e] = accept ? phi[e] : hi[e];
    }
    if (accept && g.j == 0 && active) ++n_acc;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      float4 m = make_float4(lo[e], hi[e], lo[e] * lo[e], hi[e] * hi[e]);
      if (!active) m = make_float4(0.f, 0.f, 0.f, 0.f);
      const int kk = g.j + g.gs * e;
      const float a = across_groups_sum(m.x, g.gs), b = across_groups_sum(m.y, g.gs);
      const float c = across_groups_sum(m.z, g.gs), dd = across_groups_sum(m.w, g.gs);
      if (g.lane < g.gs) {
        if (kk < g.da) { atomicAdd(st.sx + kk, (double)a); atomicAdd(st.sx2 + kk, (double)c); }
        if (kk < g.db) { atomicAdd(st.sx + g.da + kk, (double)b); atomicAdd(st.sx2 + g.da + kk, (double)dd); }
      }
    }
    if (active) {
      store_chain(row, g, lo, hi);
      if (accept && A.adjusted && A.logq_cache && g.j == 0) A.logq_cache[chain] = f_p;   // imh.py:233
      if (C.sink.samples) sink_store(C.sink, g, C.n, chain, 0, lo, hi);
    }
  }
  n_acc = __reduce_add_sync(0xffffffffu, n_acc);
  n_bad = __reduce_add_sync(0xffffffffu, n_bad);
  if ((threadIdx.x & 31) == 0) {
    if (n_acc) atomicAdd(st.cnt + 0, (unsigned long long)n_acc);
    if (n_bad) atomicAdd(st.cnt + 2, (unsigned long long)n_bad);
  }
  if (threadIdx.x == 0) {
    long long mine = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const long long first = tile * cpc;
      mine += (C.n - first) < cpc ? (C.n - first) : cpc;
    }
    atomicAdd(st.cnt + 1, (unsigned long long)mine);
  }
  cta_stats_finish(st, C.stats, C.d);
}

template <int E>
int launch_jump_accept(const AcceptArgs& A, int grid, size_t smem, cudaStream_t s) {
  NFMC_SET_SMEM_RET(jump_accept_kernel<E>, smem);
  jump_accept_kernel<E><<<occupancy_grid(jump_accept_kernel<E>, smem, A.c.n, A.c.gs), kThreads, smem, s>>>(A);
  return check_cuda(cudaGetLastError(), "jump_accept_kernel launch");
}
template int launch_jump_accept<NFMC_ONLY_E>(const AcceptArgs&, int, size_t, cudaStream_t);

template <int E>
int launch_flow_pass(const FlowArgs& A, int mode, const float* in, float* out, float* aux, long long n, int grid,
                     size_t smem, cudaStream_t s) {
  // exact-layout specialisations are instantiated for the production shapes (E = 13, 16) only; the generic
  // (non-small) conditioner path is a separate, single variant per blob placement
  constexpr bool XE = (E == 13 || E == 16);
  const bool small = flow_is_small(A.M, A.H);
  const bool xl = A.exact && XE;
#define NFMC_LAUNCH(SBv, Xv, Sv)                                                              \
  do {                                                                                        \
    NFMC_SET_SMEM_RET((flow_pass_kernel<E, SBv, Xv, Sv>), smem);                                        \
    flow_pass_kernel<E, SBv, Xv, Sv><<<occupancy_grid(flow_pass_kernel<E, SBv, Xv, Sv>, smem, n, A.gs), kThreads, smem, s>>>(A, mode, in, out, aux, n);                              \
  } while (0)
  if (!small) { if (A.stage_blob) NFMC_LAUNCH(true, false, false); else NFMC_LAUNCH(false, false, false); }
  else if (A.stage_blob && xl) NFMC_LAUNCH(true, XE, true);
  else if (A.stage_blob) NFMC_LAUNCH(true, false, true);
  else if (xl) NFMC_LAUNCH(false, XE, true);
  else NFMC_LAUNCH(false, false, true);
#undef NFMC_LAUNCH
  return check_cuda(cudaGetLastError(), "flow_pass_kernel launch");
}
template <int E>
int launch_flow_sample(const FlowArgs& A, const RngArgs& R, long long chain0, float* x, float* logq, long long n, int grid,
                       size_t smem, cudaStream_t s) {
  // exact-layout specialisations are instantiated for the production shapes (E = 13, 16) only; the generic
  // (non-small) conditioner path is a separate, single variant per blob placement
  constexpr bool XE = (E == 13 || E == 16);
  const bool small = flow_is_small(A.M, A.H);
  const bool xl = A.exact && XE;
#define NFMC_LAUNCH(SBv, Xv, Sv)                                                              \
  do {                                                                                        \
    NFMC_SET_SMEM_RET((flow_sample_kernel<E, SBv, Xv, Sv>), smem);                                        \
    flow_sample_kernel<E, SBv, Xv, Sv><<<occupancy_grid(flow_sample_kernel<E, SBv, Xv, Sv>, smem, n, A.gs), kThreads, smem, s>>>(A, R, chain0, x, logq, n);                              \
  } while (0)
  if (!small) { if (A.stage_blob) NFMC_LAUNCH(true, false, false); else NFMC_LAUNCH(false, false, false); }
  else if (A.stage_blob && xl) NFMC_LAUNCH(true, XE, true);
  else if (A.stage_blob) NFMC_LAUNCH(true, false, true);
  else if (xl) NFMC_LAUNCH(false, XE, true);
  else NFMC_LAUNCH(false, false, true);
#undef NFMC_LAUNCH
  return check_cuda(cudaGetLastError(), "flow_sample_kernel launch");
}
template <int E>
int launch_jump(const JumpArgs& A, int grid, size_t smem, cudaStream_t s) {
  // exact-layout specialisations are instantiated for the production shapes (E = 13, 16) only; the generic
  // (non-small) conditioner path is a separate, single variant per blob placement
  constexpr bool XE = (E == 13 || E == 16);
  const bool small = flow_is_small(A.f.M, A.f.H);
  const bool xl = A.f.exact && XE;
#define NFMC_LAUNCH(SBv, Xv, Sv)                                                              \
  do {                                                                                        \
    NFMC_SET_SMEM_RET((jump_kernel<E, SBv, Xv, Sv>), smem);                                        \
    jump_kernel<E, SBv, Xv, Sv><<<occupancy_grid(jump_kernel<E, SBv, Xv, Sv>, smem, A.c.n, A.c.gs), kThreads, smem, s>>>(A);                              \
  } while (0)
  if (!small) { if (A.f.stage_blob) NFMC_LAUNCH(true, false, false); else NFMC_LAUNCH(false, false, false); }
  else if (A.f.stage_blob && xl) NFMC_LAUNCH(true, XE, true);
  else if (A.f.stage_blob) NFMC_LAUNCH(true, false, true);
  else if (xl) NFMC_LAUNCH(false, XE, true);
  else NFMC_LAUNCH(false, false, true);
#undef NFMC_LAUNCH
  return check_cuda(cudaGetLastError(), "jump_kernel launch");
}
template <int E>
int launch_jump_propose_accept(const JumpArgs& A, int grid, size_t smem, cudaStream_t s) {
  // exact-layout specialisations are instantiated for the production shapes (E = 13, 16) only; the generic
  // (non-small) conditioner path is a separate, single variant per blob placement
  constexpr bool XE = (E == 13 || E == 16);
  const bool small = flow_is_small(A.f.M, A.f.H);
  const bool xl = A.f.exact && XE;
#define NFMC_LAUNCH(SBv, Xv, Sv)                                                              \
  do {                                                                                        \
    NFMC_SET_SMEM_RET((jump_propose_accept_kernel<E, SBv, Xv, Sv>), smem);                                        \
    jump_propose_accept_kernel<E, SBv, Xv, Sv><<<occupancy_grid(jump_propose_accept_kernel<E, SBv, Xv, Sv>, smem, A.c.n, A.c.gs), kThreads, smem, s>>>(A);                              \
  } while (0)
  if (!small) { if (A.f.stage_blob) NFMC_LAUNCH(true, false, false); else NFMC_LAUNCH(false, false, false); }
  else if (A.f.stage_blob && xl) NFMC_LAUNCH(true, XE, true);
  else if (A.f.stage_blob) NFMC_LAUNCH(true, false, true);
  else if (xl) NFMC_LAUNCH(false, XE, true);
  else NFMC_LAUNCH(false, false, true);
#undef NFMC_LAUNCH
  return check_cuda(cudaGetLastError(), "jump_propose_accept_kernel launch");
}
template int launch_flow_pass<NFMC_ONLY_E>(const FlowArgs&, int, const float*, float*, float*, long long, int, size_t, cudaStream_t);
template int launch_flow_sample<NFMC_ONLY_E>(const FlowArgs&, const RngArgs&, long long, float*, float*, long long, int, size_t, cudaStream_t);
template int launch_jump<NFMC_ONLY_E>(const JumpArgs&, int, size_t, cudaStream_t);
template int launch_jump_propose_accept<NFMC_ONLY_E>(const JumpArgs&, int, size_t, cudaStream_t);

}  // namespace nfmc
