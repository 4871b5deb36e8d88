// hmc_kernel.cu -- fused K-step HMC / UHMC kernel (one translation unit per E).
// Replaces HMC.propose (mcmc/hmc.py:96-126) and the local loop MCMCSampler.sample (mcmc/base.py:69-99).
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

// ---------------------------------------------------------------------------------------------------------
// HMC: p = xi / sqrt(m); L x [p -= tau/2 g; x += tau p m; p -= tau/2 g]; H = U + 1/2 sum p^2 m   (hmc.py:96-126)
// The reference evaluates grad U twice at the same point between consecutive leapfrog steps (hmc.py:69-71);
// the value is identical, so it is computed once and the two half-kicks are still applied separately.
// ---------------------------------------------------------------------------------------------------------
template <int POT, int E>
__global__ void __launch_bounds__(kThreads, NFMC_HMC_MINB) hmc_kernel(const LocalArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ChainArgs& C = A.c;
  const Geom g = make_geom(C.d, C.gs);
  CtaStats st = cta_stats_init(smem, C.d);
  float2* coef = reinterpret_cast<float2*>(smem + ((cta_stats_bytes(C.d) + 15) & ~size_t(15)));
  const bool unit_mass = (A.imd == nullptr);
  if (!unit_mass) {
    for (int i = threadIdx.x; i < C.d; i += blockDim.x) {
      const float m = __ldg(A.imd + i);
      coef[i] = make_float2(__fdiv_rn(1.f, sqrtf(m)), m);
    }
    __syncthreads();
  }
  const int cpc = kThreads / C.gs;
  const long long tiles = (C.n + cpc - 1) / cpc;
  const float half_tau = A.tau / 2;
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
      float plo[E], phi[E];  // momentum
      uint32_t ubits = 0;
      {
        StepNoise<E> nz;
        if (C.rng.normals) {
          const float* nr = C.rng.normals + ((long long)k * C.n + chain) * (long long)C.d;
          load_chain(nr, g, nz.lo, nz.hi);
          nz.ubits = 0;
        } else {
          const RngKey key = make_rng_key(C.rng.seed, 0u, C.rng.step0 + (uint64_t)k, (uint64_t)(C.chain0 + chain));
          draw_step_noise<E>(key, g.j, nz);
        }
        ubits = nz.ubits;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int kk = g.j + g.gs * e;
          plo[e] = (kk < g.da) ? nz.lo[e] : 0.f;
          phi[e] = (kk < g.db) ? nz.hi[e] : 0.f;
          if (!unit_mass) {
            plo[e] *= coef[min(kk, g.da - 1 < 0 ? 0 : g.da - 1)].x;
            phi[e] *= coef[g.da + min(kk, g.db - 1)].x;
          }
        }
      }
      float xlo[E], xhi[E];
      float kin0 = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        xlo[e] = lo[e];
        xhi[e] = hi[e];
        if (unit_mass) { kin0 = fmaf(plo[e], plo[e], fmaf(phi[e], phi[e], kin0)); }
        else {
          const int kk = g.j + g.gs * e;
          kin0 = fmaf(plo[e] * plo[e], coef[min(kk, g.da - 1 < 0 ? 0 : g.da - 1)].y, kin0);
          kin0 = fmaf(phi[e] * phi[e], coef[g.da + min(kk, g.db - 1)].y, kin0);
        }
      }
      PotCtx cur = ctx;
      for (int l = 0; l < A.n_leapfrog; ++l) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int kk = g.j + g.gs * e;
          float glo, ghi;
          pot_grad<POT>(C.pot, cur, g, kk, xlo[e], xhi[e], glo, ghi);
          plo[e] = fmaf(-half_tau, glo, plo[e]);                                     // hmc.py:51-53
          phi[e] = fmaf(-half_tau, ghi, phi[e]);
          if (unit_mass) {
            xlo[e] = fmaf(A.tau, plo[e], xlo[e]);                                    // hmc.py:56-58
            xhi[e] = fmaf(A.tau, phi[e], xhi[e]);
          } else {
            xlo[e] = fmaf(A.tau, plo[e] * coef[min(kk, g.da - 1 < 0 ? 0 : g.da - 1)].y, xlo[e]);
            xhi[e] = fmaf(A.tau, phi[e] * coef[g.da + min(kk, g.db - 1)].y, xhi[e]);
          }
          if (kk >= g.da) xlo[e] = 0.f;
          if (kk >= g.db) xhi[e] = 0.f;
        }
        cur = pot_prepare<POT, E>(C.pot, g, xlo, xhi);
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int kk = g.j + g.gs * e;
          float glo, ghi;
          pot_grad<POT>(C.pot, cur, g, kk, xlo[e], xhi[e], glo, ghi);
          plo[e] = (kk < g.da) ? fmaf(-half_tau, glo, plo[e]) : 0.f;
          phi[e] = (kk < g.db) ? fmaf(-half_tau, ghi, phi[e]) : 0.f;
        }
      }
      bool accept = true;
      if (A.adjusted) {
        float kin1 = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if (unit_mass) { kin1 = fmaf(plo[e], plo[e], fmaf(phi[e], phi[e], kin1)); }
          else {
            const int kk = g.j + g.gs * e;
            kin1 = fmaf(plo[e] * plo[e], coef[min(kk, g.da - 1 < 0 ? 0 : g.da - 1)].y, kin1);
            kin1 = fmaf(phi[e] * phi[e], coef[g.da + min(kk, g.db - 1)].y, kin1);
          }
        }
        const float h0 = ctx.u + 0.5f * group_sum(kin0, g.gs);                        // hmc.py:103-106
        const float h1 = cur.u + 0.5f * group_sum(kin1, g.gs);                        // hmc.py:107-110
        const float log_acc = -h1 - (-h0);                                            // hmc.py:111
        float u;
        if (C.rng.uniforms) u = __ldg(C.rng.uniforms + (long long)k * C.n + chain);
        else u = uniform_from_bits(__shfl_sync(0xffffffffu, ubits, g.grp_base));
        accept = logf(u) < log_acc;                                                   // hmc.py:112-113
        if (!(fabsf(log_acc) <= 3.0e38f) && g.j == 0 && active) ++n_bad;
      }
#pragma unroll
      for (int e = 0; e < E; ++e) {
        lo[e] = accept ? xlo[e] : lo[e];
        hi[e] = accept ? xhi[e] : hi[e];
      }
      ctx = select_ctx(accept, cur, ctx);
      if (accept && g.j == 0 && active) ++n_acc;
      accumulate_moments(lo, hi, m1lo, m1hi, m2lo, m2hi);
      if (C.sink.samples && active) sink_store(C.sink, g, C.n, chain, k, lo, hi);
    }
    if (!active) {
#pragma unroll
      for (int e = 0; e < E; ++e) m1lo[e] = m1hi[e] = m2lo[e] = m2hi[e] = 0.f;
    }
    flush_moments(g, m1lo, m1hi, m2lo, m2hi, st.sx, st.sx2);
    if (active) store_chain(row, g, lo, hi);
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
    atomicAdd(st.cnt + 1, (unsigned long long)(mine * C.n_steps));
  }
  cta_stats_finish(st, C.stats, C.d);
}


template <int E>
int launch_hmc(int pot_kind, const LocalArgs& A, int grid, size_t smem, cudaStream_t s) {
  NFMC_DISPATCH_POT(pot_kind, {
    NFMC_SET_SMEM_RET((hmc_kernel<POT, E>), smem);
    hmc_kernel<POT, E><<<grid, kThreads, smem, s>>>(A);
  });
  return check_cuda(cudaGetLastError(), "hmc_kernel launch");
}
template int launch_hmc<NFMC_ONLY_E>(int, const LocalArgs&, int, size_t, cudaStream_t);

}  // namespace nfmc
