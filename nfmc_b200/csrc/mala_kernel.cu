// mala_kernel.cu -- fused K-step MALA / ULA kernel (one translation unit per E).
//
// One group of `gs` lanes owns one chain for the whole launch: the state stays in registers across all K
// steps, so HBM sees the chain state once on the way in and once on the way out (plus the optional sample
// sink / injected noise).  Replaces Langevin.propose (mcmc/langevin.py:61-122) and the local loop
// MCMCSampler.sample (mcmc/base.py:69-99) of the reference.
//
// Register budget: 4 CTAs of 128 threads per SM (<= 128 registers).  Live per-lane state is the chain (2E) and
// the proposal (2E); the noise is consumed Philox quad by quad as it is generated, and the running moments
// (4E floats per lane) live in shared memory as float4 {sum lo, sum hi, sum lo^2, sum hi^2} per slot.
#include "launchers.cuh"

#ifndef NFMC_ONLY_E
#error "compile with -DNFMC_ONLY_E=<slots per half>"
#endif

#ifndef NFMC_MALA_MINB
#define NFMC_MALA_MINB 4
#endif

namespace nfmc {

// per-dimension coefficients in shared memory when the mass is not the identity
//   {c1 = -tau/imd^2, c2 = sqrt(2 tau)/imd, tauA = tau/imd^2, invA = imd^2}   (langevin.py:74-75,95)
__device__ __forceinline__ float4 mala_coef(float tau, float s2t, float m) {
  const float a = __fdiv_rn(1.f, m * m);
  return make_float4(-__fdiv_rn(tau, m * m), __fdiv_rn(s2t, m), tau * a, __fdiv_rn(1.f, a));
}

// FAST: exact layout, Philox noise, identity mass -- all three known at compile time (single straight-line step body).
// !FAST: inexact layouts, injected noise and non-identity mass handled by run-time (warp-uniform) tests.
template <int POT, int E, bool FAST>
__global__ void __launch_bounds__(kThreads, NFMC_MALA_MINB) mala_kernel(const LocalArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ChainArgs& C = A.c;
  const Geom g = make_geom(C.d, C.gs);
  CtaStats st = cta_stats_init(smem, C.d);
  size_t off = (cta_stats_bytes(C.d) + 15) & ~size_t(15);
  float4* coef = reinterpret_cast<float4*>(smem + off);
  const bool unit_mass = FAST ? true : (A.imd == nullptr);
  const bool rw = !FAST && A.random_walk;
  if (!unit_mass) {
    for (int i = threadIdx.x; i < C.d; i += blockDim.x) {
      const float mi = __ldg(A.imd + i);
      coef[i] = rw ? make_float4(0.f, mi, 0.f, 0.f) : mala_coef(A.tau, A.sqrt_2tau, mi);
    }
    off += (size_t)C.d * sizeof(float4);
    __syncthreads();
  }
  float4* mom = reinterpret_cast<float4*>(smem + off) + threadIdx.x;  // slot e at mom[e * kThreads]
  const int cpc = kThreads / C.gs;
  const long long tiles = (C.n + cpc - 1) / cpc;
  const float inv4tau = __fdiv_rn(1.f, 4.f * A.tau);
  constexpr bool EXACT = FAST;
  const bool inject = FAST ? false : (C.rng.normals != nullptr);
  const PhiloxKeys PK = philox_keys(C.rng.seed);
  unsigned int n_acc = 0, n_bad = 0;
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
    PotCtx ctx = pot_prepare<POT, E>(C.pot, g, lo, hi);

    for (int k = 0; k < C.n_steps; ++k) {
      const RngKey key = make_rng_key(C.rng.seed, 0u, C.rng.step0 + (uint64_t)k, (uint64_t)(C.chain0 + chain));
      const float* nrow = inject ? C.rng.normals + ((long long)k * C.n + chain) * (long long)C.d : nullptr;
      float plo[E], phi[E];
      float qf = 0.f, qfh = 0.f;
      uint32_t ubits = 0;
      // ---- noise (langevin.py:63) consumed quad by quad; proposal x' = x - tau/m^2 grad U + sqrt(2 tau)/m xi (:74-76)
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        uint4 w = make_uint4(0u, 0u, 0u, 0u);
        if (!inject) w = rng_quad(PK, key, q, g.j);
        if (q == 0) ubits = w.x;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int e = 2 * q + hh - 1;          // pair p = e + 1 lives in quad p/2, words 2(p%2), 2(p%2)+1
          if (e < 0 || e >= E) continue;
          const int kk = g.j + g.gs * e;
          const bool vl = slot_ok<EXACT, E>(e, kk, g.da), vh = slot_ok<EXACT, E>(e, kk, g.db);
          float nlo, nhi;
          if (inject) {
            nlo = vl ? __ldg(nrow + kk) : 0.f;
            nhi = vh ? __ldg(nrow + g.da + kk) : 0.f;
          } else {
            const float2 zz = box_muller2(hh ? w.z : w.x, hh ? w.w : w.y);
            nlo = zz.x; nhi = zz.y;
          }
          float glo, ghi;
          if (EXACT && e < E - 1) pot_grad<POT, true>(C.pot, ctx, g, kk, lo[e], hi[e], glo, ghi);
          else pot_grad<POT, false>(C.pot, ctx, g, kk, lo[e], hi[e], glo, ghi);
          float pl, ph, tl, th;
          if (FAST) {
            // packed fp32x2: (lo, hi) of a slot go through FFMA2 / FADD2 together (bit-identical to the scalar branch)
            const float2 xv = make_float2(lo[e], hi[e]), gv = make_float2(glo, ghi);
            const float2 pv = fma2(splat2(A.sqrt_2tau), make_float2(nlo, nhi), fma2(splat2(-A.tau), gv, xv));
            float2 tv = fma2(splat2(A.tau), gv, sub2(pv, xv));
            pl = pv.x; ph = pv.y;
            tv.x = vl ? tv.x : 0.f;
            tv.y = vh ? tv.y : 0.f;
            const float2 qv = fma2(tv, tv, make_float2(qf, qfh));
            qf = qv.x; qfh = qv.y;
          } else if (rw && unit_mass) {
            pl = lo[e] + nlo;                                                  // mh.py:52-56 with imd = 1
            ph = hi[e] + nhi;
          } else if (unit_mass) {
            pl = fmaf(A.sqrt_2tau, nlo, fmaf(-A.tau, glo, lo[e]));
            ph = fmaf(A.sqrt_2tau, nhi, fmaf(-A.tau, ghi, hi[e]));
            tl = pl - lo[e] + A.tau * glo;                                     // langevin.py:41
            th = ph - hi[e] + A.tau * ghi;
            tl = vl ? tl : 0.f;
            th = vh ? th : 0.f;
            qf = fmaf(tl, tl, qf); qfh = fmaf(th, th, qfh);
          } else {
            const float4 cl = coef[vl ? kk : 0], ch = coef[g.da + (vh ? kk : 0)];
            pl = fmaf(cl.y, nlo, fmaf(cl.x, glo, lo[e]));
            ph = fmaf(ch.y, nhi, fmaf(ch.x, ghi, hi[e]));
            tl = pl - lo[e] + cl.z * glo;
            th = ph - hi[e] + ch.z * ghi;
            tl = vl ? tl : 0.f;
            th = vh ? th : 0.f;
            qf = fmaf(tl * cl.w, tl, qf); qfh = fmaf(th * ch.w, th, qfh);
          }
          plo[e] = vl ? pl : 0.f;
          phi[e] = vh ? ph : 0.f;
        }
      }
      bool accept = true;
      const PotCtx ctxp = pot_prepare<POT, E>(C.pot, g, plo, phi);  // U(x') (langevin.py:80-82); also next step's ctx
      if (A.adjusted) {
        // ---- reverse proposal term with grad U(x')  (langevin.py:91-97) ----------------------------------
        float qr = 0.f, qrh = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int kk = g.j + g.gs * e;
          const bool vl = slot_ok<EXACT, E>(e, kk, g.da), vh = slot_ok<EXACT, E>(e, kk, g.db);
          float glo, ghi;
          if (EXACT && e < E - 1) pot_grad<POT, true>(C.pot, ctxp, g, kk, plo[e], phi[e], glo, ghi);
          else pot_grad<POT, false>(C.pot, ctxp, g, kk, plo[e], phi[e], glo, ghi);
          float tl, th;
          if (FAST) {
            const float2 xv = make_float2(lo[e], hi[e]), pv = make_float2(plo[e], phi[e]);
            float2 tv = fma2(splat2(A.tau), make_float2(glo, ghi), sub2(xv, pv));
            tv.x = vl ? tv.x : 0.f;
            tv.y = vh ? tv.y : 0.f;
            const float2 qv = fma2(tv, tv, make_float2(qr, qrh));
            qr = qv.x; qrh = qv.y;
          } else if (unit_mass) {
            tl = lo[e] - plo[e] + A.tau * glo;
            th = hi[e] - phi[e] + A.tau * ghi;
            tl = vl ? tl : 0.f;
            th = vh ? th : 0.f;
            qr = fmaf(tl, tl, qr); qrh = fmaf(th, th, qrh);
          } else {
            const float4 cl = coef[vl ? kk : 0], ch = coef[g.da + (vh ? kk : 0)];
            tl = lo[e] - plo[e] + cl.z * glo;
            th = hi[e] - phi[e] + ch.z * ghi;
            tl = vl ? tl : 0.f;
            th = vh ? th : 0.f;
            qr = fmaf(tl * cl.w, tl, qr); qrh = fmaf(th * ch.w, th, qrh);
          }
        }
        qf = group_sum(qf + qfh, g.gs) * inv4tau;
        qr = group_sum(qr + qrh, g.gs) * inv4tau;
        // util.py:392 with target = -U, proposal = -Q   (langevin.py:88-105)
        const float log_ratio = rw ? ((-ctxp.u) - (-ctx.u) + 0.f - 0.f)          // mh.py:59
                                   : ((-ctxp.u) - (-ctx.u) + (-qr) - (-qf));
        float u;
        if (C.rng.uniforms) u = __ldg(C.rng.uniforms + (long long)k * C.n + chain);
        else u = uniform_from_bits(__shfl_sync(0xffffffffu, ubits, g.grp_base));
        accept = logf(u) < log_ratio;                                                    // langevin.py:106
        if (!(fabsf(log_ratio) <= 3.0e38f) && g.j == 0 && active) ++n_bad;
      }
      // ---- x[mask] = x'[mask] (mcmc/base.py:77); running moments of the post-accept state (mcmc/base.py:86) ----
#pragma unroll
      for (int e = 0; e < E; ++e) {
        lo[e] = accept ? plo[e] : lo[e];
        hi[e] = accept ? phi[e] : hi[e];
        float4 m = mom[e * kThreads];
        const float2 xv = make_float2(lo[e], hi[e]);
        const float2 m1 = add2(make_float2(m.x, m.y), xv), m2 = fma2(xv, xv, make_float2(m.z, m.w));
        mom[e * kThreads] = make_float4(m1.x, m1.y, m2.x, m2.y);
      }
      ctx = select_ctx(accept, ctxp, ctx);
      if (accept && g.j == 0 && active) ++n_acc;
      if (C.sink.samples && active) sink_store(C.sink, g, C.n, chain, k, lo, hi);        // mcmc/base.py:90
    }
    // ---- per-tile flush of the moments: fp32 per-lane sums -> shuffle over the warp's groups -> fp64 shared atomics
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
      mine += (C.n - first) < cpc ? (C.n - first) : cpc;
    }
    atomicAdd(st.cnt + 1, (unsigned long long)(mine * C.n_steps));
  }
  cta_stats_finish(st, C.stats, C.d);
}

// ---------------------------------------------------------------------------------------------------------
// FAST kernel: exact layout, Philox noise, identity mass.  Same arithmetic as mala_kernel (bit-identical results), but
// the state is held as float2 {lo[e], hi[e]} pairs and the float work goes through the packed FFMA2 / FADD2 / FMUL2
// instructions: the kernel is bound by instruction issue, and a packed instruction does two fp32 operations per slot.
// ---------------------------------------------------------------------------------------------------------
template <int POT, int E>
__global__ void __launch_bounds__(kThreads, NFMC_MALA_MINB) mala_fast_kernel(const LocalArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ChainArgs& C = A.c;
  const Geom g = make_geom(C.d, C.gs);
  CtaStats st = cta_stats_init(smem, C.d);
  const size_t off = (cta_stats_bytes(C.d) + 15) & ~size_t(15);
  float4* mom = reinterpret_cast<float4*>(smem + off) + threadIdx.x;  // slot e at mom[e * kThreads]
  const int cpc = kThreads / C.gs;
  const long long tiles = (C.n + cpc - 1) / cpc;
  const float inv4tau = __fdiv_rn(1.f, 4.f * A.tau);
  const PhiloxKeys PK = philox_keys(C.rng.seed);
  const float2 S2T = splat2(A.sqrt_2tau), TAU = splat2(A.tau), NTAU = splat2(-A.tau);
  unsigned int n_acc = 0, n_bad = 0;
  constexpr int NQ = (E + 2) / 2;
  // only the last slot can be invalid in an exact layout
  const int k_last = g.j + g.gs * (E - 1);
  const bool vl_last = k_last < g.da, vh_last = k_last < g.db;

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
    auto grad = [&](const PotCtx& c, int e, float2 v) {
      if constexpr (POT == NFMC_POT_ISO_GAUSSIAN) {
        return mul2(splat2(C.pot.s0), v);
      } else {
        float glo, ghi;
        if (e < E - 1) pot_grad<POT, true>(C.pot, c, g, g.j + g.gs * e, v.x, v.y, glo, ghi);
        else pot_grad<POT, false>(C.pot, c, g, g.j + g.gs * e, v.x, v.y, glo, ghi);
        return make_float2(glo, ghi);
      }
    };
    PotCtx ctx = prepare(x);

    for (int k = 0; k < C.n_steps; ++k) {
      const RngKey key = make_rng_key(C.rng.seed, 0u, C.rng.step0 + (uint64_t)k, (uint64_t)(C.chain0 + chain));
      float2 p[E];
      float2 qf = make_float2(0.f, 0.f);
      uint32_t ubits = 0;
      // ---- noise (langevin.py:63) consumed quad by quad; proposal x' = x - tau grad U + sqrt(2 tau) xi (:74-76) --------
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const uint4 w = rng_quad(PK, key, q, g.j);
        if (q == 0) ubits = w.x;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int e = 2 * q + hh - 1;
          if (e < 0 || e >= E) continue;
          const float2 nz = box_muller2(hh ? w.z : w.x, hh ? w.w : w.y);
          const float2 gv = grad(ctx, e, x[e]);
          float2 pv = fma2(S2T, nz, fma2(NTAU, gv, x[e]));
          float2 tv = fma2(TAU, gv, sub2(pv, x[e]));                               // langevin.py:41
          if (e == E - 1) {
            pv.x = vl_last ? pv.x : 0.f; pv.y = vh_last ? pv.y : 0.f;
            tv.x = vl_last ? tv.x : 0.f; tv.y = vh_last ? tv.y : 0.f;
          }
          qf = fma2(tv, tv, qf);
          p[e] = pv;
        }
      }
      bool accept = true;
      const PotCtx ctxp = prepare(p);                                              // U(x') (langevin.py:80-82)
      if (A.adjusted) {
        float2 qr = make_float2(0.f, 0.f);
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const float2 gv = grad(ctxp, e, p[e]);
          float2 tv = fma2(TAU, gv, sub2(x[e], p[e]));                             // langevin.py:91-97
          if (e == E - 1) { tv.x = vl_last ? tv.x : 0.f; tv.y = vh_last ? tv.y : 0.f; }
          qr = fma2(tv, tv, qr);
        }
        const float qfs = group_sum(qf.x + qf.y, g.gs) * inv4tau;
        const float qrs = group_sum(qr.x + qr.y, g.gs) * inv4tau;
        const float log_ratio = (-ctxp.u) - (-ctx.u) + (-qrs) - (-qfs);            // util.py:392, langevin.py:88-105
        const float u = uniform_from_bits(__shfl_sync(0xffffffffu, ubits, g.grp_base));
        accept = logf(u) < log_ratio;                                              // langevin.py:106
        if (!(fabsf(log_ratio) <= 3.0e38f) && g.j == 0 && active) ++n_bad;
      }
#pragma unroll
      for (int e = 0; e < E; ++e) {                                                // mcmc/base.py:77,86
        x[e].x = accept ? p[e].x : x[e].x;
        x[e].y = accept ? p[e].y : x[e].y;
        const float4 m = mom[e * kThreads];
        const float2 m1 = add2(make_float2(m.x, m.y), x[e]), m2 = fma2(x[e], x[e], make_float2(m.z, m.w));
        mom[e * kThreads] = make_float4(m1.x, m1.y, m2.x, m2.y);
      }
      ctx = select_ctx(accept, ctxp, ctx);
      if (accept && g.j == 0 && active) ++n_acc;
      if (C.sink.samples && active) {                                              // mcmc/base.py:90
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
int launch_mala(int pot_kind, bool exact, const LocalArgs& A, int grid, size_t smem, cudaStream_t s) {
  NFMC_DISPATCH_POT(pot_kind, {
    if (exact && !A.c.rng.normals && !A.imd && !A.random_walk) {
      NFMC_SET_SMEM_RET((mala_fast_kernel<POT, E>), smem);
      mala_fast_kernel<POT, E><<<occupancy_grid(mala_fast_kernel<POT, E>, smem, A.c.n, A.c.gs), kThreads, smem, s>>>(A);
    } else {
      NFMC_SET_SMEM_RET((mala_kernel<POT, E, false>), smem);
      mala_kernel<POT, E, false><<<occupancy_grid(mala_kernel<POT, E, false>, smem, A.c.n, A.c.gs), kThreads, smem, s>>>(A);
    }
  });
  return check_cuda(cudaGetLastError(), "mala_kernel launch");
}
template int launch_mala<NFMC_ONLY_E>(int, bool, const LocalArgs&, int, size_t, cudaStream_t);

}  // namespace nfmc
