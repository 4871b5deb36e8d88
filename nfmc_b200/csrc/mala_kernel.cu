// mala_kernel.cu -- fused K-step MALA / ULA kernel (one translation unit per E).
//
// One group of `gs` lanes owns one chain for the whole launch: the state stays in registers across all K
// steps, so HBM sees the chain state once on the way in and once on the way out (plus the optional sample
// sink / injected noise).  Replaces Langevin.propose (mcmc/langevin.py:61-122) and the local loop
// MCMCSampler.sample (mcmc/base.py:69-99) of the reference.
#include "launchers.cuh"

#ifndef NFMC_ONLY_E
#error "compile with -DNFMC_ONLY_E=<slots per half>"
#endif

#ifndef NFMC_MALA_MINB
#define NFMC_MALA_MINB 4
#endif
#ifndef NFMC_HMC_MINB
#define NFMC_HMC_MINB 4
#endif

namespace nfmc {

// per-dimension coefficients in shared memory when the mass is not the identity
//   MALA: {c1 = -tau/imd^2, c2 = sqrt(2 tau)/imd, tauA = tau/imd^2, invA = imd^2}   (langevin.py:74-75,95)
//   HMC : {rs = 1/sqrt(imd), imd, 0, 0}                                              (hmc.py:100,58,104)
__device__ __forceinline__ float4 mala_coef(float tau, float s2t, float m) {
  const float a = __fdiv_rn(1.f, m * m);
  return make_float4(-__fdiv_rn(tau, m * m), __fdiv_rn(s2t, m), tau * a, __fdiv_rn(1.f, a));
}

template <int POT, int E>
__global__ void __launch_bounds__(kThreads, NFMC_MALA_MINB) mala_kernel(const LocalArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ChainArgs& C = A.c;
  const Geom g = make_geom(C.d, C.gs);
  CtaStats st = cta_stats_init(smem, C.d);
  float4* coef = reinterpret_cast<float4*>(smem + ((cta_stats_bytes(C.d) + 15) & ~size_t(15)));
  const bool unit_mass = (A.imd == nullptr);
  if (!unit_mass) {
    for (int i = threadIdx.x; i < C.d; i += blockDim.x) coef[i] = mala_coef(A.tau, A.sqrt_2tau, __ldg(A.imd + i));
    __syncthreads();
  }
  const int cpc = kThreads / C.gs;
  const long long tiles = (C.n + cpc - 1) / cpc;
  const float inv4tau = __fdiv_rn(1.f, 4.f * A.tau);
  unsigned int n_acc = 0, n_bad = 0;

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long chain_raw = tile * cpc + threadIdx.x / C.gs;
    const bool active = chain_raw < C.n;
    const long long chain = active ? chain_raw : C.n - 1;
    float* row = C.x + chain * (long long)C.d;

    float lo[E], hi[E], m1lo[E], m1hi[E], m2lo[E], m2hi[E];
    load_chain(row, g, lo, hi);
#pragma unroll
    for (int e = 0; e < E; ++e) m1lo[e] = m1hi[e] = m2lo[e] = m2hi[e] = 0.f;
    PotCtx ctx = pot_prepare<POT, E>(C.pot, g, lo, hi);

    for (int k = 0; k < C.n_steps; ++k) {
      // ---- noise for this step (langevin.py:63) ---------------------------------------------------------
      StepNoise<E> nz;
      if (C.rng.normals) {
        const float* nr = C.rng.normals + ((long long)k * C.n + chain) * (long long)C.d;
        load_chain(nr, g, nz.lo, nz.hi);
        nz.ubits = 0;
      } else {
        const RngKey key = make_rng_key(C.rng.seed, 0u, C.rng.step0 + (uint64_t)k, (uint64_t)(C.chain0 + chain));
        draw_step_noise<E>(key, g.j, nz);
      }
      // ---- proposal x' = x - tau/m^2 grad U(x) + sqrt(2 tau)/m xi  (langevin.py:74-76) ------------------
      float plo[E], phi[E];
      float qf = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int kk = g.j + g.gs * e;
        float glo, ghi;
        pot_grad<POT>(C.pot, ctx, g, kk, lo[e], hi[e], glo, ghi);
        if (unit_mass) {
          plo[e] = fmaf(A.sqrt_2tau, nz.lo[e], fmaf(-A.tau, glo, lo[e]));
          phi[e] = fmaf(A.sqrt_2tau, nz.hi[e], fmaf(-A.tau, ghi, hi[e]));
          const float tl = plo[e] - lo[e] + A.tau * glo, th = phi[e] - hi[e] + A.tau * ghi;  // langevin.py:41
          if (kk < g.da) qf = fmaf(tl, tl, qf);
          if (kk < g.db) qf = fmaf(th, th, qf);
        } else {
          const float4 cl = coef[min(kk, g.da - 1 < 0 ? 0 : g.da - 1)], ch = coef[g.da + min(kk, g.db - 1)];
          plo[e] = fmaf(cl.y, nz.lo[e], fmaf(cl.x, glo, lo[e]));
          phi[e] = fmaf(ch.y, nz.hi[e], fmaf(ch.x, ghi, hi[e]));
          const float tl = plo[e] - lo[e] + cl.z * glo, th = phi[e] - hi[e] + ch.z * ghi;
          if (kk < g.da) qf = fmaf(tl * cl.w, tl, qf);
          if (kk < g.db) qf = fmaf(th * ch.w, th, qf);
        }
        if (kk >= g.da) plo[e] = 0.f;
        if (kk >= g.db) phi[e] = 0.f;
      }
      bool accept = true;
      PotCtx ctxp = pot_prepare<POT, E>(C.pot, g, plo, phi);  // U(x') (langevin.py:80-82); also next step's ctx
      if (A.adjusted) {
        // ---- reverse proposal term with grad U(x')  (langevin.py:91-97) ----------------------------------
        float qr = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int kk = g.j + g.gs * e;
          float glo, ghi;
          pot_grad<POT>(C.pot, ctxp, g, kk, plo[e], phi[e], glo, ghi);
          if (unit_mass) {
            const float tl = lo[e] - plo[e] + A.tau * glo, th = hi[e] - phi[e] + A.tau * ghi;
            if (kk < g.da) qr = fmaf(tl, tl, qr);
            if (kk < g.db) qr = fmaf(th, th, qr);
          } else {
            const float4 cl = coef[min(kk, g.da - 1 < 0 ? 0 : g.da - 1)], ch = coef[g.da + min(kk, g.db - 1)];
            const float tl = lo[e] - plo[e] + cl.z * glo, th = hi[e] - phi[e] + ch.z * ghi;
            if (kk < g.da) qr = fmaf(tl * cl.w, tl, qr);
            if (kk < g.db) qr = fmaf(th * ch.w, th, qr);
          }
        }
        qf = group_sum(qf, g.gs) * inv4tau;
        qr = group_sum(qr, g.gs) * inv4tau;
        // util.py:392 with target = -U, proposal = -Q   (langevin.py:88-105)
        const float log_ratio = (-ctxp.u) - (-ctx.u) + (-qr) - (-qf);
        float u;
        if (C.rng.uniforms) u = __ldg(C.rng.uniforms + (long long)k * C.n + chain);
        else u = uniform_from_bits(__shfl_sync(0xffffffffu, nz.ubits, g.grp_base));
        accept = logf(u) < log_ratio;                                                    // langevin.py:106
        if (!(fabsf(log_ratio) <= 3.0e38f) && g.j == 0 && active) ++n_bad;
      }
      // ---- x[mask] = x'[mask]  (mcmc/base.py:77) -----------------------------------------------------------
#pragma unroll
      for (int e = 0; e < E; ++e) {
        lo[e] = accept ? plo[e] : lo[e];
        hi[e] = accept ? phi[e] : hi[e];
      }
      ctx = select_ctx(accept, ctxp, ctx);
      if (accept && g.j == 0 && active) ++n_acc;
      accumulate_moments(lo, hi, m1lo, m1hi, m2lo, m2hi);                                // mcmc/base.py:86
      if (C.sink.samples && active) sink_store(C.sink, g, C.n, chain, k, lo, hi);        // mcmc/base.py:90
    }
    if (!active) {
#pragma unroll
      for (int e = 0; e < E; ++e) m1lo[e] = m1hi[e] = m2lo[e] = m2hi[e] = 0.f;
    }
    flush_moments(g, m1lo, m1hi, m2lo, m2hi, st.sx, st.sx2);
    if (active) store_chain(row, g, lo, hi);
  }
  // counters (mcmc/base.py:79-85): one atomic per warp into shared, then one per CTA into global
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
      const long long cnt = (C.n - first) < cpc ? (C.n - first) : cpc;
      mine += cnt;
    }
    atomicAdd(st.cnt + 1, (unsigned long long)(mine * C.n_steps));
  }
  cta_stats_finish(st, C.stats, C.d);
}


template <int E>
int launch_mala(int pot_kind, const LocalArgs& A, int grid, size_t smem, cudaStream_t s) {
  NFMC_DISPATCH_POT(pot_kind, {
    NFMC_SET_SMEM_RET((mala_kernel<POT, E>), smem);
    mala_kernel<POT, E><<<grid, kThreads, smem, s>>>(A);
  });
  return check_cuda(cudaGetLastError(), "mala_kernel launch");
}
template int launch_mala<NFMC_ONLY_E>(int, const LocalArgs&, int, size_t, cudaStream_t);

}  // namespace nfmc
