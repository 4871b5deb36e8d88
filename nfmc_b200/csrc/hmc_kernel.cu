// hmc_kernel.cu -- fused K-step HMC / UHMC kernel (one translation unit per E).
// Replaces HMC.propose (mcmc/hmc.py:96-126) and the local loop MCMCSampler.sample (mcmc/base.py:69-99).
//
//   p = xi / sqrt(m);  L x [ p -= tau/2 grad U(x);  x += tau p m;  p -= tau/2 grad U(x) ];  H = U + 1/2 sum p^2 m
//
// The reference evaluates grad U twice at every interior point of the trajectory (end of leapfrog l, start of
// leapfrog l+1, hmc.py:69-71).  The value is identical, so it is computed once per point (L+1 evaluations) and the
// two half-kicks are still applied as two separate roundings.  For potentials whose gradient is elementwise
// (Gaussians, Rosenbrock pairs) no reduction happens inside the trajectory; U is evaluated at its two ends only.
#include "launchers.cuh"

#ifndef NFMC_ONLY_E
#error "compile with -DNFMC_ONLY_E=<slots per half>"
#endif

#ifndef NFMC_HMC_MINB
#define NFMC_HMC_MINB 4
#endif

namespace nfmc {

// FAST: exact layout, Philox noise, identity mass -- all three known at compile time (single straight-line step body).
// !FAST: inexact layouts, injected noise and non-identity mass handled by run-time (warp-uniform) tests.
template <int POT, int E, bool FAST>
__global__ void __launch_bounds__(kThreads, NFMC_HMC_MINB) hmc_kernel(const LocalArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ChainArgs& C = A.c;
  const Geom g = make_geom(C.d, C.gs);
  CtaStats st = cta_stats_init(smem, C.d);
  size_t off = (cta_stats_bytes(C.d) + 15) & ~size_t(15);
  // per-dimension {1/sqrt(m), m, 0, 0} when the mass is not the identity (hmc.py:100,58,104)
  float4* coef = reinterpret_cast<float4*>(smem + off);
  const bool unit_mass = FAST ? true : (A.imd == nullptr);
  if (!unit_mass) {
    for (int i = threadIdx.x; i < C.d; i += blockDim.x) {
      const float m = __ldg(A.imd + i);
      coef[i] = make_float4(__fdiv_rn(1.f, sqrtf(m)), m, 0.f, 0.f);
    }
    off += (size_t)C.d * sizeof(float4);
    __syncthreads();
  }
  float4* mom = reinterpret_cast<float4*>(smem + off) + threadIdx.x;
  const int cpc = kThreads / C.gs;
  const long long tiles = (C.n + cpc - 1) / cpc;
  const float half_tau = A.tau / 2;
  constexpr bool EXACT = FAST;
  const bool inject = FAST ? false : (C.rng.normals != nullptr);
  const PhiloxKeys PK = philox_keys(C.rng.seed);
  unsigned int n_acc = 0, n_bad = 0;
  constexpr int NQ = (E + 2) / 2;
  constexpr bool CTX = pot_grad_needs_ctx<POT>();

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long chain_raw = tile * cpc + threadIdx.x / C.gs;
    const bool active = chain_raw < C.n;
    const long long chain = active ? chain_raw : C.n - 1;
    float* row = C.x + chain * (long long)C.d;

    float lo[E], hi[E];
    load_chain(row, g, lo, hi);
#pragma unroll
    for (int e = 0; e < E; ++e) mom[e * kThreads] = make_float4(0.f, 0.f, 0.f, 0.f);
    PotCtx ctx = pot_prepare<POT, E>(C.pot, g, lo, hi);

    for (int k = 0; k < C.n_steps; ++k) {
      const RngKey key = make_rng_key(C.rng.seed, 0u, C.rng.step0 + (uint64_t)k, (uint64_t)(C.chain0 + chain));
      const float* nrow = inject ? C.rng.normals + ((long long)k * C.n + chain) * (long long)C.d : nullptr;
      float plo[E], phi[E], xlo[E], xhi[E];
      float kin0 = 0.f;
      uint32_t ubits = 0;
      // ---- momentum p = xi / sqrt(m) (hmc.py:100), first half-kick with grad U(x0) (hmc.py:51-53) ------------------
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        uint4 w = make_uint4(0u, 0u, 0u, 0u);
        if (!inject) w = rng_quad(PK, key, q, g.j);
        if (q == 0) ubits = w.x;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int e = 2 * q + hh - 1;
          if (e < 0 || e >= E) continue;
          const int kk = g.j + g.gs * e;
          const bool vl = slot_ok<EXACT, E>(e, kk, g.da), vh = slot_ok<EXACT, E>(e, kk, g.db);
          float nlo, nhi;
          if (inject) {
            nlo = vl ? __ldg(nrow + kk) : 0.f;
            nhi = vh ? __ldg(nrow + g.da + kk) : 0.f;
          } else {
            box_muller(hh ? w.z : w.x, hh ? w.w : w.y, nlo, nhi);
          }
          nlo = vl ? nlo : 0.f;
          nhi = vh ? nhi : 0.f;
          float ml = 1.f, mh = 1.f;
          if (!unit_mass) {
            const float4 cl = coef[vl ? kk : 0], ch = coef[g.da + (vh ? kk : 0)];
            nlo *= cl.x; nhi *= ch.x;
            ml = cl.y; mh = ch.y;
          }
          kin0 = fmaf(nlo * nlo, ml, fmaf(nhi * nhi, mh, kin0));
          float glo, ghi;
          if (EXACT && e < E - 1) pot_grad<POT, true>(C.pot, ctx, g, kk, lo[e], hi[e], glo, ghi);
          else pot_grad<POT, false>(C.pot, ctx, g, kk, lo[e], hi[e], glo, ghi);
          plo[e] = (A.n_leapfrog > 0) ? fmaf(-half_tau, glo, nlo) : nlo;
          phi[e] = (A.n_leapfrog > 0) ? fmaf(-half_tau, ghi, nhi) : nhi;
          xlo[e] = lo[e];
          xhi[e] = hi[e];
        }
      }
      // ---- trajectory (hmc.py:61-77) -----------------------------------------------------------------------------
      PotCtx cur = ctx;
      for (int l = 0; l < A.n_leapfrog; ++l) {
        const bool more = l + 1 < A.n_leapfrog;
        if (CTX) {
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const int kk = g.j + g.gs * e;
            const bool vl = slot_ok<EXACT, E>(e, kk, g.da), vh = slot_ok<EXACT, E>(e, kk, g.db);
            float ml = 1.f, mh = 1.f;
            if (!unit_mass) { ml = coef[vl ? kk : 0].y; mh = coef[g.da + (vh ? kk : 0)].y; }
            xlo[e] = vl ? fmaf(A.tau, unit_mass ? plo[e] : plo[e] * ml, xlo[e]) : 0.f;      // hmc.py:56-58
            xhi[e] = vh ? fmaf(A.tau, unit_mass ? phi[e] : phi[e] * mh, xhi[e]) : 0.f;
          }
          cur = pot_prepare<POT, E>(C.pot, g, xlo, xhi);
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int kk = g.j + g.gs * e;
          const bool vl = slot_ok<EXACT, E>(e, kk, g.da), vh = slot_ok<EXACT, E>(e, kk, g.db);
          if (!CTX) {
            float ml = 1.f, mh = 1.f;
            if (!unit_mass) { ml = coef[vl ? kk : 0].y; mh = coef[g.da + (vh ? kk : 0)].y; }
            xlo[e] = vl ? fmaf(A.tau, unit_mass ? plo[e] : plo[e] * ml, xlo[e]) : 0.f;
            xhi[e] = vh ? fmaf(A.tau, unit_mass ? phi[e] : phi[e] * mh, xhi[e]) : 0.f;
          }
          float glo, ghi;
          if (EXACT && e < E - 1) pot_grad<POT, true>(C.pot, cur, g, kk, xlo[e], xhi[e], glo, ghi);
          else pot_grad<POT, false>(C.pot, cur, g, kk, xlo[e], xhi[e], glo, ghi);
          float pl = fmaf(-half_tau, glo, plo[e]), ph = fmaf(-half_tau, ghi, phi[e]);       // second half-kick
          if (more) { pl = fmaf(-half_tau, glo, pl); ph = fmaf(-half_tau, ghi, ph); }         // next step's first
          plo[e] = vl ? pl : 0.f;
          phi[e] = vh ? ph : 0.f;
        }
      }
      bool accept = true;
      if (!CTX && A.n_leapfrog > 0) cur = pot_prepare<POT, E>(C.pot, g, xlo, xhi);
      if (A.adjusted) {
        float kin1 = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if (unit_mass) { kin1 = fmaf(plo[e], plo[e], fmaf(phi[e], phi[e], kin1)); }
          else {
            const int kk = g.j + g.gs * e;
            const bool vl = slot_ok<EXACT, E>(e, kk, g.da), vh = slot_ok<EXACT, E>(e, kk, g.db);
            kin1 = fmaf(plo[e] * plo[e], coef[vl ? kk : 0].y, fmaf(phi[e] * phi[e], coef[g.da + (vh ? kk : 0)].y, kin1));
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
        float4 m = mom[e * kThreads];
        m.x += lo[e];
        m.y += hi[e];
        m.z = fmaf(lo[e], lo[e], m.z);
        m.w = fmaf(hi[e], hi[e], m.w);
        mom[e * kThreads] = m;
      }
      ctx = select_ctx(accept, cur, ctx);
      if (accept && g.j == 0 && active) ++n_acc;
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

// ---------------------------------------------------------------------------------------------------------
// FAST kernel: exact layout, Philox noise, identity mass.  Same arithmetic as hmc_kernel<.., true> (bit-identical
// results), with the state, momentum and trajectory point held as float2 {lo[e], hi[e]} pairs so that the leapfrog
// updates issue as packed FFMA2 (two fp32 operations per slot and instruction): the trajectory is pure fp32 issue.
// ---------------------------------------------------------------------------------------------------------
template <int POT, int E>
__global__ void __launch_bounds__(kThreads, NFMC_HMC_MINB) hmc_fast_kernel(const LocalArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ChainArgs& C = A.c;
  const Geom g = make_geom(C.d, C.gs);
  CtaStats st = cta_stats_init(smem, C.d);
  const size_t off = (cta_stats_bytes(C.d) + 15) & ~size_t(15);
  float4* mom = reinterpret_cast<float4*>(smem + off) + threadIdx.x;
  const int cpc = kThreads / C.gs;
  const long long tiles = (C.n + cpc - 1) / cpc;
  const float half_tau = A.tau / 2;
  const PhiloxKeys PK = philox_keys(C.rng.seed);
  unsigned int n_acc = 0, n_bad = 0;
  constexpr int NQ = (E + 2) / 2;
  constexpr bool CTX = pot_grad_needs_ctx<POT>();
  constexpr bool DIAG = POT == NFMC_POT_DIAG_GAUSSIAN;
  const int k_last = g.j + g.gs * (E - 1);
  const bool vl_last = k_last < g.da, vh_last = k_last < g.db;   // only the last slot can be invalid in an exact layout
  // Invalid slots are kept at exactly zero WITHOUT select instructions inside the trajectory: the state and the masked
  // momentum draw start at 0 there, and the last slot's step constants are 0 in its invalid components, so drift and
  // kicks leave them at 0 (the selects and the predicates re-derived for them were 20 % of the loop's issue slots).
  const float2 TAU = splat2(A.tau), NHT = splat2(-half_tau);
  const float2 TAU_L = make_float2(vl_last ? A.tau : 0.f, vh_last ? A.tau : 0.f);
  const float2 NHT_L = make_float2(vl_last ? -half_tau : 0.f, vh_last ? -half_tau : 0.f);
  // Diagonal Gaussian: this lane's precisions {w[k], w[da + k]} and negated means per slot, staged once in shared memory
  // (zero in invalid slots) -- the gradient is then one packed multiply per slot instead of two parameter loads from
  // global memory per slot and leapfrog.  `centered` (every mean is 0, e.g. the ill-conditioned Gaussian of config C2)
  // drops the subtraction: x - 0 == x bit for bit.
  // (one entry per slot and LANE, so that every access is this thread's base pointer + a compile-time offset)
  float2* wt = reinterpret_cast<float2*>(smem + off + (size_t)E * kThreads * sizeof(float4));
  float2* nm = wt + E * 32;
  bool centered = true;
  if constexpr (DIAG) {
    const float2* wm = reinterpret_cast<const float2*>(C.pot.params);
    int any_mean = 0;
    for (int i = threadIdx.x; i < E * 32; i += kThreads) {
      const int e = i >> 5, kk = ((i & 31) & (g.gs - 1)) + g.gs * e;
      const float2 pl = kk < g.da ? __ldg(wm + kk) : make_float2(0.f, 0.f);
      const float2 ph = kk < g.db ? __ldg(wm + g.da + kk) : make_float2(0.f, 0.f);
      wt[i] = make_float2(pl.x, ph.x);
      nm[i] = make_float2(-pl.y, -ph.y);
      any_mean |= (pl.y != 0.f) | (ph.y != 0.f);
    }
    centered = !__syncthreads_or(any_mean);
    wt += g.j;      // lanes with the same j read the same address: one shared-memory wavefront per access
    nm += g.j;
  }

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long chain_raw = tile * cpc + threadIdx.x / C.gs;
    const bool active = chain_raw < C.n;
    const long long chain = active ? chain_raw : C.n - 1;
    float* row = C.x + chain * (long long)C.d;

    float2 x[E];
    {
      float lo[E], hi[E];
      load_chain(row, g, lo, hi);
#pragma unroll
      for (int e = 0; e < E; ++e) x[e] = make_float2(lo[e], hi[e]);
    }
#pragma unroll
    for (int e = 0; e < E; ++e) mom[e * kThreads] = make_float4(0.f, 0.f, 0.f, 0.f);
    auto prepare = [&](const float2 (&v)[E]) {
      float lo[E], hi[E];
#pragma unroll
      for (int e = 0; e < E; ++e) { lo[e] = v[e].x; hi[e] = v[e].y; }
      return pot_prepare<POT, E>(C.pot, g, lo, hi);
    };
    // `cen` (a std::bool_constant): the diagonal Gaussian's means are all zero -- decided once per kernel, so the
    // trajectory exists in two straight-line versions instead of carrying both operands and a select per slot
    auto grad = [&](auto cen, const PotCtx& c, int e, float2 v) {
      if constexpr (POT == NFMC_POT_ISO_GAUSSIAN) {
        return mul2(splat2(C.pot.s0), v);
      } else if constexpr (DIAG) {
        if constexpr (decltype(cen)::value) return mul2(wt[e * 32], v);
        else return mul2(wt[e * 32], add2(v, nm[e * 32]));
      } else {
        float glo, ghi;
        if (e < E - 1) pot_grad<POT, true>(C.pot, c, g, g.j + g.gs * e, v.x, v.y, glo, ghi);
        else pot_grad<POT, false>(C.pot, c, g, g.j + g.gs * e, v.x, v.y, glo, ghi);
        return make_float2(glo, ghi);
      }
    };
    auto mask_last = [&](int e, float2 v) {
      if (e == E - 1) { v.x = vl_last ? v.x : 0.f; v.y = vh_last ? v.y : 0.f; }
      return v;
    };
    PotCtx ctx = prepare(x);

    for (int k = 0; k < C.n_steps; ++k) {
      const RngKey key = make_rng_key(C.rng.seed, 0u, C.rng.step0 + (uint64_t)k, (uint64_t)(C.chain0 + chain));
      float2 p[E], xt[E];
      float kin0 = 0.f;
      uint32_t ubits = 0;
      PotCtx cur = ctx;
      auto propose = [&](auto cen) {
      // ---- momentum p = xi (hmc.py:100), first half-kick with grad U(x0) (hmc.py:51-53) ---------------------------------
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const uint4 w = rng_quad(PK, key, q, g.j);
        if (q == 0) ubits = w.x;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int e = 2 * q + hh - 1;
          if (e < 0 || e >= E) continue;
          float nlo, nhi;
          box_muller(hh ? w.z : w.x, hh ? w.w : w.y, nlo, nhi);
          const float2 nz = mask_last(e, make_float2(nlo, nhi));
          kin0 = fmaf(nz.x * nz.x, 1.f, fmaf(nz.y * nz.y, 1.f, kin0));
          const float2 gv = grad(cen, ctx, e, x[e]);
          p[e] = (A.n_leapfrog > 0) ? fma2(e == E - 1 ? NHT_L : NHT, gv, nz) : nz;
          xt[e] = x[e];
        }
      }
      // ---- trajectory (hmc.py:61-77) -----------------------------------------------------------------------------
      for (int l = 0; l < A.n_leapfrog; ++l) {
        const bool more = l + 1 < A.n_leapfrog;
        if (CTX) {
#pragma unroll
          for (int e = 0; e < E; ++e) xt[e] = fma2(e == E - 1 ? TAU_L : TAU, p[e], xt[e]);    // hmc.py:56-58
          cur = prepare(xt);
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if (!CTX) xt[e] = fma2(e == E - 1 ? TAU_L : TAU, p[e], xt[e]);
          const float2 gv = grad(cen, cur, e, xt[e]);
          const float2 nht = e == E - 1 ? NHT_L : NHT;
          float2 pv = fma2(nht, gv, p[e]);                                                  // second half-kick
          if (more) pv = fma2(nht, gv, pv);                                                 // next step's first
          p[e] = pv;
        }
      }
      };
      if constexpr (DIAG) {
        if (centered) propose(std::true_type{});
        else propose(std::false_type{});
      } else {
        propose(std::true_type{});
      }
      bool accept = true;
      if (!CTX && A.n_leapfrog > 0) cur = prepare(xt);
      if (A.adjusted) {
        float kin1 = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) kin1 = fmaf(p[e].x, p[e].x, fmaf(p[e].y, p[e].y, kin1));
        const float h0 = ctx.u + 0.5f * group_sum(kin0, g.gs);                        // hmc.py:103-106
        const float h1 = cur.u + 0.5f * group_sum(kin1, g.gs);                        // hmc.py:107-110
        const float log_acc = -h1 - (-h0);                                            // hmc.py:111
        const float u = uniform_from_bits(__shfl_sync(0xffffffffu, ubits, g.grp_base));
        accept = logf(u) < log_acc;                                                   // hmc.py:112-113
        if (!(fabsf(log_acc) <= 3.0e38f) && g.j == 0 && active) ++n_bad;
      }
#pragma unroll
      for (int e = 0; e < E; ++e) {
        x[e].x = accept ? xt[e].x : x[e].x;
        x[e].y = accept ? xt[e].y : x[e].y;
        const float4 m = mom[e * kThreads];
        const float2 m1 = add2(make_float2(m.x, m.y), x[e]), m2 = fma2(x[e], x[e], make_float2(m.z, m.w));
        mom[e * kThreads] = make_float4(m1.x, m1.y, m2.x, m2.y);
      }
      ctx = select_ctx(accept, cur, ctx);
      if (accept && g.j == 0 && active) ++n_acc;
      if (C.sink.samples && active) {
        float lo[E], hi[E];
#pragma unroll
        for (int e = 0; e < E; ++e) { lo[e] = x[e].x; hi[e] = x[e].y; }
        sink_store(C.sink, g, C.n, chain, k, lo, hi);
      }
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
    if (active) {
      float lo[E], hi[E];
#pragma unroll
      for (int e = 0; e < E; ++e) { lo[e] = x[e].x; hi[e] = x[e].y; }
      store_chain(row, g, lo, hi);
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
    atomicAdd(st.cnt + 1, (unsigned long long)(mine * C.n_steps));
  }
  cta_stats_finish(st, C.stats, C.d);
}

template <int E>
int launch_hmc(int pot_kind, bool exact, const LocalArgs& A, int grid, size_t smem, cudaStream_t s) {
  NFMC_DISPATCH_POT(pot_kind, {
    if (exact && !A.c.rng.normals && !A.imd) {
      // measured (2^20 chains, d = 100, L = 20, ms per step, packed vs scalar): iso 0.55 / 0.67, funnel 1.00 / 1.19,
      // mixture 1.64 / 1.81, diagonal 0.64 / 0.83 (precisions staged per lane in shared memory; with two parameter loads
      // from global memory per slot and leapfrog the packed form took 1.11), Rosenbrock 0.84 / 0.73 -- its gradient pairs
      // the two halves of a slot, so it keeps the scalar kernel
      constexpr bool PACKED = POT == NFMC_POT_ISO_GAUSSIAN || POT == NFMC_POT_FUNNEL || POT == NFMC_POT_MIXTURE4 ||
                              POT == NFMC_POT_DIAG_GAUSSIAN;
      if constexpr (PACKED) {
        NFMC_SET_SMEM_RET((hmc_fast_kernel<POT, E>), smem);
        hmc_fast_kernel<POT, E><<<occupancy_grid(hmc_fast_kernel<POT, E>, smem, A.c.n, A.c.gs), kThreads, smem, s>>>(A);
      } else {
        NFMC_SET_SMEM_RET((hmc_kernel<POT, E, true>), smem);
        hmc_kernel<POT, E, true><<<occupancy_grid(hmc_kernel<POT, E, true>, smem, A.c.n, A.c.gs), kThreads, smem, s>>>(A);
      }
    } else {
      NFMC_SET_SMEM_RET((hmc_kernel<POT, E, false>), smem);
      hmc_kernel<POT, E, false><<<occupancy_grid(hmc_kernel<POT, E, false>, smem, A.c.n, A.c.gs), kThreads, smem, s>>>(A);
    }
  });
  return check_cuda(cudaGetLastError(), "hmc_kernel launch");
}
template int launch_hmc<NFMC_ONLY_E>(int, bool, const LocalArgs&, int, size_t, cudaStream_t);

}  // namespace nfmc
