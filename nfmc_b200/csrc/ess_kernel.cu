// ess_kernel.cu -- fused K-step elliptical slice sampling kernel (one translation unit per E).
//
// Replaces elliptical_slice_sampling_step (mcmc/ess.py:12-64, identity prior covariance) and ESS.propose
// (mcmc/ess.py:97-116) inside the local loop MCMCSampler.sample (mcmc/base.py:69-99).  Same ownership model as the
// Langevin kernel: a group of `gs` lanes keeps one chain (2E floats per lane) and its ellipse direction nu (2E floats)
// in registers for all K steps; the bracket-shrinking loop is at most `max_iterations` evaluations of the negative
// log-likelihood per step, each one pass over the registers plus a group reduction.
//
// Random numbers per (chain, step): nu = d normals (Philox stream 0, pairs 1..E as everywhere else) and 2 + M
// uniforms {u (:35), theta0 (:39), M bracket draws (:58)} -- injected as uniforms[step][chain][2 + M], or taken from
// Philox stream 2, counter quad i/4 word i%4 of lane j = 0.
#include "launchers.cuh"

#ifndef NFMC_ONLY_E
#error "compile with -DNFMC_ONLY_E=<slots per half>"
#endif

namespace nfmc {

constexpr float kPiF = 3.14159274101257324f;       // (float) torch.pi
constexpr float kTwoPiF = 6.28318548202514648f;    // (float) (2 * torch.pi)

template <int POT, int E, bool EXACT>
__global__ void __launch_bounds__(kThreads, 4) ess_kernel(const EssArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ChainArgs& C = A.c;
  const Geom g = make_geom(C.d, C.gs);
  CtaStats st = cta_stats_init(smem, C.d);
  const size_t off = (cta_stats_bytes(C.d) + 15) & ~size_t(15);
  float4* mom = reinterpret_cast<float4*>(smem + off) + threadIdx.x;  // slot e at mom[e * kThreads]
  const int cpc = kThreads / C.gs;
  const long long tiles = (C.n + cpc - 1) / cpc;
  const bool inject = C.rng.normals != nullptr;
  const PhiloxKeys PK = philox_keys(C.rng.seed);
  const int M = A.max_iterations;
  const int n_uni = 2 + M;
  unsigned int n_found = 0;
  constexpr int NQ = (E + 2) / 2;

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long chain_raw = tile * cpc + threadIdx.x / C.gs;
    const bool active = chain_raw < C.n;
    const long long chain = active ? chain_raw : C.n - 1;
    float* row = C.x + chain * (long long)C.d;

    float lo[E], hi[E];
    load_chain(row, g, lo, hi);
#pragma unroll
    for (int e = 0; e < E; ++e) mom[e * kThreads] = make_float4(0.f, 0.f, 0.f, 0.f);
    float u_cur = pot_prepare<POT, E>(C.pot, g, lo, hi).u;                       // nll(f)

    for (int k = 0; k < C.n_steps; ++k) {
      const uint64_t step = C.rng.step0 + (uint64_t)k;
      const RngKey key = make_rng_key(C.rng.seed, 0u, step, (uint64_t)(C.chain0 + chain));
      const RngKey ukey = make_rng_key(C.rng.seed, 2u, step, (uint64_t)(C.chain0 + chain));
      const float* nrow = inject ? C.rng.normals + ((long long)k * C.n + chain) * (long long)C.d : nullptr;
      const float* urow = C.rng.uniforms ? C.rng.uniforms + ((long long)k * C.n + chain) * (long long)n_uni : nullptr;
      uint4 uq = make_uint4(0u, 0u, 0u, 0u);
      int uq_idx = -1;
      auto uniform = [&](int i) {                                                  // i-th scalar uniform of this step
        if (urow) return __ldg(urow + i);
        if ((i >> 2) != uq_idx) { uq_idx = i >> 2; uq = rng_quad(PK, ukey, uq_idx, 0); }
        const uint32_t w = (i & 3) == 0 ? uq.x : (i & 3) == 1 ? uq.y : (i & 3) == 2 ? uq.z : uq.w;
        return uniform_from_bits(w);
      };

      // ---- 1. ellipse direction nu ~ N(0, I) (ess.py:32) ---------------------------------------------------------
      float nlo[E], nhi[E];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        uint4 w = make_uint4(0u, 0u, 0u, 0u);
        if (!inject) w = rng_quad(PK, key, q, g.j);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int e = 2 * q + hh - 1;
          if (e < 0 || e >= E) continue;
          const int kk = g.j + g.gs * e;
          const bool vl = slot_ok<EXACT, E>(e, kk, g.da), vh = slot_ok<EXACT, E>(e, kk, g.db);
          float a, b;
          if (inject) {
            a = vl ? __ldg(nrow + kk) : 0.f;
            b = vh ? __ldg(nrow + g.da + kk) : 0.f;
          } else {
            const float2 zz = box_muller2(hh ? w.z : w.x, hh ? w.w : w.y);
            a = vl ? zz.x : 0.f;
            b = vh ? zz.y : 0.f;
          }
          nlo[e] = a; nhi[e] = b;
        }
      }
      // ---- 2. log-likelihood threshold (ess.py:35-36); 3. initial angle and bracket (:39-41) ---------------------
      const float log_y = __fadd_rn(-u_cur, logf(uniform(0)));
      float theta = __fmul_rn(__fmul_rn(uniform(1), 2.f), kPiF);
      float t_min = __fsub_rn(theta, kTwoPiF), t_max = theta;
      bool found = false;
      for (int it = 0; it < M; ++it) {
        if (__all_sync(0xffffffffu, found)) break;                               // later rounds cannot change a found chain (:50)
        float sn, cs;
        sincosf(theta, &sn, &cs);
        float plo[E], phi[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {                                            // f' = f cos(theta) + nu sin(theta) (:46)
          plo[e] = __fadd_rn(__fmul_rn(lo[e], cs), __fmul_rn(nlo[e], sn));
          phi[e] = __fadd_rn(__fmul_rn(hi[e], cs), __fmul_rn(nhi[e], sn));
        }
        const float u_p = pot_prepare<POT, E>(C.pot, g, plo, phi).u;
        const bool upd = (-u_p > log_y) && !found;                               // :47,50
        if (upd) {
#pragma unroll
          for (int e = 0; e < E; ++e) { lo[e] = plo[e]; hi[e] = phi[e]; }
          u_cur = u_p;
        }
        if (theta < 0.f) t_min = theta; else t_max = theta;                      // :53-55
        theta = __fadd_rn(__fmul_rn(uniform(2 + it), __fsub_rn(t_max, t_min)), t_min);   // :58-59
        found = found || upd;                                                    // :62
      }
      if (found && g.j == 0 && active) ++n_found;
      // ---- mask is all ones (ess.py:107): running moments and the sample row of the post-step state --------------
#pragma unroll
      for (int e = 0; e < E; ++e) {
        float4 m = mom[e * kThreads];
        m.x += lo[e]; m.y += hi[e];
        m.z = fmaf(lo[e], lo[e], m.z); m.w = fmaf(hi[e], hi[e], m.w);
        mom[e * kThreads] = m;
      }
      if (C.sink.samples && active) sink_store(C.sink, g, C.n, chain, k, lo, hi);
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      float4 m = mom[e * kThreads];
      if (!active) m = make_float4(0.f, 0.f, 0.f, 0.f);
      const int kk = g.j + g.gs * e;
      const float a = across_groups_sum(m.x, g.gs), b = across_groups_sum(m.y, g.gs);
      const float c = across_groups_sum(m.z, g.gs), dd = across_groups_sum(m.w, g.gs);
      if (g.lane < g.gs) {
        if (kk < g.da) { atomicAdd(st.sx + kk, (double)a); atomicAdd(st.sx2 + kk, (double)c); }
        if (kk < g.db) { atomicAdd(st.sx + g.da + kk, (double)b); atomicAdd(st.sx2 + g.da + kk, (double)dd); }
      }
    }
    if (active) store_chain(row, g, lo, hi);
  }
  n_found = __reduce_add_sync(0xffffffffu, n_found);
  if ((threadIdx.x & 31) == 0 && n_found) atomicAdd(st.cnt + 3, (unsigned long long)n_found);
  if (threadIdx.x == 0) {
    long long mine = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const long long first = tile * cpc;
      mine += (C.n - first) < cpc ? (C.n - first) : cpc;
    }
    // every chain counts as accepted in every step (mcmc/ess.py:107 + mcmc/base.py:79-85)
    atomicAdd(st.cnt + 0, (unsigned long long)(mine * C.n_steps));
    atomicAdd(st.cnt + 1, (unsigned long long)(mine * C.n_steps));
  }
  cta_stats_finish(st, C.stats, C.d);
}

template <int E>
int launch_ess(int pot_kind, bool exact, const EssArgs& A, int grid, size_t smem, cudaStream_t s) {
  NFMC_DISPATCH_POT(pot_kind, {
    if (exact) {
      NFMC_SET_SMEM_RET((ess_kernel<POT, E, true>), smem);
      ess_kernel<POT, E, true><<<occupancy_grid(ess_kernel<POT, E, true>, smem, A.c.n, A.c.gs), kThreads, smem, s>>>(A);
    } else {
      NFMC_SET_SMEM_RET((ess_kernel<POT, E, false>), smem);
      ess_kernel<POT, E, false><<<occupancy_grid(ess_kernel<POT, E, false>, smem, A.c.n, A.c.gs), kThreads, smem, s>>>(A);
    }
  });
  return check_cuda(cudaGetLastError(), "ess_kernel launch");
}
template int launch_ess<NFMC_ONLY_E>(int, bool, const EssArgs&, int, size_t, cudaStream_t);

}  // namespace nfmc
