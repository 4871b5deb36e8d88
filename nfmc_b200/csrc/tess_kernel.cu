// tess_kernel.cu -- transport elliptical slice sampling (one translation unit per E).
//
// Replaces transport_elliptical_slice_sampling_step (nfmc/tess.py:15-75, identity covariance) inside TESS.sample
// (nfmc/tess.py:151-188): the chain state is the latent u; every step draws an ellipse direction v ~ N(0, I), a threshold
//   log s = log pi^(u) + log phi(v) + log w,     log pi^(u) = -U(T^-1 u) - log|det dT^-1/du|  (sign as tess.py:31),
// and walks at most M bracket rounds u' = u cos t + v sin t, v' = v cos t - u sin t, accepting the first round with
// log pi^(u') + log phi(v') > log s.  The recorded sample is the data-space point x = T^-1(u) of the step's outcome.
// Reference quirks kept: the initial angle is a NORMAL draw times 2 pi (:44-45); theta_max aliases it (:48).
//
// Per lane: u, v and one working vector (u' -> x') live in registers; the step's current outcome (x, u) sits in per-thread
// shared memory.  One flow inverse pass per bracket round plus one per step.
// Random numbers per (chain, step): v = d normals (Philox stream 0, pairs 1..E) and the scalars {w, theta_n, M bracket
// uniforms} -- injected as uniforms[step][chain][2 + M] = {w, theta_n (a normal), bracket...}, or from Philox stream 2
// on lane j = 0: word 0 -> w, Box-Muller(words 1, 2) -> theta_n, word 3 + i -> bracket draw i.
#include "launchers.cuh"

#ifndef NFMC_ONLY_E
#error "compile with -DNFMC_ONLY_E=<slots per half>"
#endif

namespace nfmc {

template <int E, bool SB, bool X, bool SM>
__global__ void __launch_bounds__(kThreads, 2) tess_kernel(const TessArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ChainArgs& C = A.c;
  const Geom g = make_geom(C.d, C.gs);
  FlowSmem S = flow_smem_init<SB>(smem, A.f, true);
  const bool flip = (A.f.Lc & 1) != 0;
  // per-thread outcome of the step in shared memory: x[2E] then u[2E], element i at keep[i * kThreads]
  float* keep = reinterpret_cast<float*>(S.mom - threadIdx.x + (size_t)E * kThreads) + threadIdx.x;
  const int cpc = kThreads / C.gs;
  const long long tiles = (C.n + cpc - 1) / cpc;
  const int M = A.max_iterations;
  const int n_uni = 2 + M;
  const float half_d_log2pi = 0.5f * (float)C.d * 1.8378770664093453f;
  const float two_pi = 6.28318548202514648f;   // (float)(2 * torch.pi)
  const PhiloxKeys PK = philox_keys(C.rng.seed);
  unsigned int n_found = 0;

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long chain_raw = tile * cpc + threadIdx.x / C.gs;
    const bool active = chain_raw < C.n;
    const long long chain = active ? chain_raw : C.n - 1;
    float* row = C.x + chain * (long long)C.d;
    float ulo[E], uhi[E];
    if (flip) load_chain_flipped(row, g, ulo, uhi); else load_chain(row, g, ulo, uhi);
#pragma unroll
    for (int e = 0; e < E; ++e) S.mom[e * kThreads] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int k = 0; k < C.n_steps; ++k) {
      const uint64_t step = C.rng.step0 + (uint64_t)k;
      // ---- ellipse direction v (tess.py:38) -------------------------------------------------------------------------
      float vlo[E], vhi[E];
      {
        StepNoise<E> nz;
        if (C.rng.normals) {
          const float* nr = C.rng.normals + ((long long)k * C.n + chain) * (long long)C.d;
          if (flip) load_chain_flipped(nr, g, nz.lo, nz.hi); else load_chain(nr, g, nz.lo, nz.hi);
        } else {
          const RngKey key = make_rng_key(C.rng.seed, 0u, step, (uint64_t)(C.chain0 + chain));
          draw_step_noise<E>(key, g.j, nz);
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int kk = g.j + g.gs * e;
          vlo[e] = slot_ok<X, E>(e, kk, g.da) ? nz.lo[e] : 0.f;
          vhi[e] = slot_ok<X, E>(e, kk, g.db) ? nz.hi[e] : 0.f;
        }
      }
      // ---- scalars {w, theta_n, bracket draws} ------------------------------------------------------------------------
      const float* urow = C.rng.uniforms ? C.rng.uniforms + ((long long)k * C.n + chain) * (long long)n_uni : nullptr;
      const RngKey skey = make_rng_key(C.rng.seed, 2u, step, (uint64_t)(C.chain0 + chain));
      uint4 sq = make_uint4(0u, 0u, 0u, 0u);
      int sq_idx = -1;
      auto sword = [&](int i) {                                                   // i-th scalar word of this step
        if ((i >> 2) != sq_idx) { sq_idx = i >> 2; sq = rng_quad(PK, skey, sq_idx, 0); }
        return (i & 3) == 0 ? sq.x : (i & 3) == 1 ? sq.y : (i & 3) == 2 ? sq.z : sq.w;
      };
      float w, theta_n;
      if (urow) { w = __ldg(urow); theta_n = __ldg(urow + 1); }
      else {
        w = uniform_from_bits(sword(0));
        float t1;
        const uint32_t a = sword(1), b = sword(2);
        box_muller(a, b, theta_n, t1);
      }
      // ---- x = T^-1(u), log pi^(u) (tess.py:29-32,51-52) ----------------------------------------------------------------
      float xlo[E], xhi[E];
#pragma unroll
      for (int e = 0; e < E; ++e) { xlo[e] = ulo[e]; xhi[e] = uhi[e]; }
      float ld = flow_inverse<E, SB, X, SM>(S.F, g, xlo, xhi, S.scr);
      float pu = pot_prepare_rt<E>(A.pot_kind, C.pot, g, xlo, xhi).u;
#pragma unroll
      for (int e = 0; e < E; ++e) { keep[e * kThreads] = xlo[e]; keep[(E + e) * kThreads] = xhi[e]; }
      float sv = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) sv = fmaf(vlo[e], vlo[e], fmaf(vhi[e], vhi[e], sv));
      const float log_phi_v = -0.5f * group_sum(sv, g.gs) - half_d_log2pi;
      const float log_s = (-pu - ld) + log_phi_v + logf(w);                        // :42
      float theta = __fmul_rn(theta_n, two_pi);                                    // :45
      float t_min = __fsub_rn(theta, two_pi), t_max = theta;                       // :48
      bool found = false;
      for (int it = 0; it < M; ++it) {
        if (__all_sync(0xffffffffu, found)) break;                               // only the first acceptable round counts (:60-61)
        float sn, cs;
        sincosf(theta, &sn, &cs);
        float s2 = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const float a = __fsub_rn(__fmul_rn(vlo[e], cs), __fmul_rn(ulo[e], sn));   // v' (:55)
          const float b = __fsub_rn(__fmul_rn(vhi[e], cs), __fmul_rn(uhi[e], sn));
          s2 = fmaf(a, a, fmaf(b, b, s2));
          xlo[e] = __fadd_rn(__fmul_rn(ulo[e], cs), __fmul_rn(vlo[e], sn));          // u' (:54)
          xhi[e] = __fadd_rn(__fmul_rn(uhi[e], cs), __fmul_rn(vhi[e], sn));
        }
        ld = flow_inverse<E, SB, X, SM>(S.F, g, xlo, xhi, S.scr);                    // x' (:56)
        pu = pot_prepare_rt<E>(A.pot_kind, C.pot, g, xlo, xhi).u;
        const float lhs = (-pu - ld) + (-0.5f * group_sum(s2, g.gs) - half_d_log2pi);
        const bool upd = (lhs > log_s) && !found;                                  // :57,60
        if (upd) {
#pragma unroll
          for (int e = 0; e < E; ++e) {
            keep[e * kThreads] = xlo[e];
            keep[(E + e) * kThreads] = xhi[e];
            keep[(2 * E + e) * kThreads] = __fadd_rn(__fmul_rn(ulo[e], cs), __fmul_rn(vlo[e], sn));
            keep[(3 * E + e) * kThreads] = __fadd_rn(__fmul_rn(uhi[e], cs), __fmul_rn(vhi[e], sn));
          }
        }
        if (theta < 0.f) t_min = theta; else t_max = theta;                        // :64-66
        const float un = urow ? __ldg(urow + 2 + it) : uniform_from_bits(sword(3 + it));
        theta = __fadd_rn(__fmul_rn(un, __fsub_rn(t_max, t_min)), t_min);          // :69-70
        found = found || upd;                                                      // :73
      }
      if (found) {
#pragma unroll
        for (int e = 0; e < E; ++e) { ulo[e] = keep[(2 * E + e) * kThreads]; uhi[e] = keep[(3 * E + e) * kThreads]; }
        if (g.j == 0 && active) ++n_found;
      }
      // ---- the recorded sample is x (tess.py:179-180) ---------------------------------------------------------------------
#pragma unroll
      for (int e = 0; e < E; ++e) { xlo[e] = keep[e * kThreads]; xhi[e] = keep[(E + e) * kThreads]; }
#pragma unroll
      for (int e = 0; e < E; ++e) {
        float4 m = S.mom[e * kThreads];
        m.x += xlo[e]; m.y += xhi[e]; m.z = fmaf(xlo[e], xlo[e], m.z); m.w = fmaf(xhi[e], xhi[e], m.w);
        S.mom[e * kThreads] = m;
      }
      if (C.sink.samples && active) sink_store(C.sink, g, C.n, chain, k, xlo, xhi);
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      float4 m = S.mom[e * kThreads];
      if (!active) m = make_float4(0.f, 0.f, 0.f, 0.f);
      const int kk = g.j + g.gs * e;
      const float a = across_groups_sum(m.x, g.gs), b = across_groups_sum(m.y, g.gs);
      const float c = across_groups_sum(m.z, g.gs), dd = across_groups_sum(m.w, g.gs);
      if (g.lane < g.gs) {
        if (kk < g.da) { atomicAdd(S.st.sx + kk, (double)a); atomicAdd(S.st.sx2 + kk, (double)c); }
        if (kk < g.db) { atomicAdd(S.st.sx + g.da + kk, (double)b); atomicAdd(S.st.sx2 + g.da + kk, (double)dd); }
      }
    }
    if (active) { if (flip) store_chain_flipped(row, g, ulo, uhi); else store_chain(row, g, ulo, uhi); }
  }
  n_found = __reduce_add_sync(0xffffffffu, n_found);
  if ((threadIdx.x & 31) == 0 && n_found) atomicAdd(S.st.cnt + 0, (unsigned long long)n_found);
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

template <int E>
int launch_tess(const TessArgs& A, int grid, size_t smem, cudaStream_t s) {
  constexpr bool XE = (E == 13 || E == 16);
  const bool small = flow_is_small(A.f.M, A.f.H);
  const bool xl = A.f.exact && XE;
#define NFMC_LAUNCH(SBv, Xv, Sv)                                          \
  do {                                                                    \
    NFMC_SET_SMEM_RET((tess_kernel<E, SBv, Xv, Sv>), smem);               \
    tess_kernel<E, SBv, Xv, Sv><<<occupancy_grid(tess_kernel<E, SBv, Xv, Sv>, smem, A.c.n, A.c.gs), kThreads, smem, s>>>(A);          \
  } while (0)
  if (!small) { if (A.f.stage_blob) NFMC_LAUNCH(true, false, false); else NFMC_LAUNCH(false, false, false); }
  else if (A.f.stage_blob && xl) NFMC_LAUNCH(true, XE, true);
  else if (A.f.stage_blob) NFMC_LAUNCH(true, false, true);
  else if (xl) NFMC_LAUNCH(false, XE, true);
  else NFMC_LAUNCH(false, false, true);
#undef NFMC_LAUNCH
  return check_cuda(cudaGetLastError(), "tess_kernel launch");
}
template int launch_tess<NFMC_ONLY_E>(const TessArgs&, int, size_t, cudaStream_t);

}  // namespace nfmc
