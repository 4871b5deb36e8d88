// neutra_kernels.cu -- NeuTra HMC: HMC in the latent space of the flow, gradient through the flow by a
// hand-written input-VJP (the flow is frozen, so only d/dz is needed).  One translation unit per E.
//
// Replaces NeuTra.adjusted_target (nfmc/neutra.py:58-68) evaluated and differentiated by autograd inside
// HMC.propose (mcmc/hmc.py:40-48,96-126):  U~(z) = U(T^-1 z) - log|det dT^-1/dz|.
// The flow is invertible, so no activation is stored: the backward sweep walks x -> z again, re-deriving
// each layer's input from its output while it pulls the gradient back (flow.cuh: flow_unwind).
// The reference evaluates grad U~ 2L times per step, twice at each interior point; here L+1 times.
// Outputs (samples, moments) are latent-space quantities, as in the reference.
#include "launchers.cuh"

#ifndef NFMC_ONLY_E
#error "compile with -DNFMC_ONLY_E=<slots per half>"
#endif

#ifndef NFMC_NEUTRA_MINB
#define NFMC_NEUTRA_MINB 2
#endif

namespace nfmc {

// U~(z) and its gradient.  (zlo, zhi) physical-order latent; (glo, ghi) receives dU~/dz.
template <int E, bool SB, bool X, bool SM>
__device__ __forceinline__ float neutra_value_grad(const FlowDesc& F, int pot_kind, const PotParams& P, const Geom& g, const float (&zlo)[E],
                                                   const float (&zhi)[E], float (&glo)[E], float (&ghi)[E], float* scr,
                                                   bool want_grad, float* stash = nullptr) {
  float xlo[E], xhi[E];
#pragma unroll
  for (int e = 0; e < E; ++e) { xlo[e] = zlo[e]; xhi[e] = zhi[e]; }
  const float ld_inv = flow_inverse<E, SB, X, SM>(F, g, xlo, xhi, scr, want_grad ? stash : nullptr);   // neutra.py:60
  const PotCtx c = pot_prepare_rt<E>(pot_kind, P, g, xlo, xhi);
  const float value = -((-c.u) + ld_inv);                                          // neutra.py:62-64
  if (want_grad) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int kk = g.j + g.gs * e;
      pot_grad_rt(pot_kind, P, c, g, kk, xlo[e], xhi[e], glo[e], ghi[e]);
      if (kk >= g.da) glo[e] = 0.f;
      if (kk >= g.db) ghi[e] = 0.f;
    }
    flow_unwind<E, SB, X, SM>(F, g, xlo, xhi, glo, ghi, scr, stash);
  }
  return value;
}



// Register budget: the live per-lane state is the latent z, the momentum p and the gradient g (2E each) plus the flow's
// working copy; the start state of a step is NOT kept (on rejection it is re-read from global memory, which always holds
// the current state), the running moments live in shared memory, and a non-identity mass is read from shared memory.
template <int E, bool SB, bool X, bool SM>
__global__ void __launch_bounds__(kThreads, NFMC_NEUTRA_MINB) neutra_hmc_kernel(const NeutraArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ChainArgs& C = A.c;
  const Geom g = make_geom(C.d, C.gs);
  FlowSmem S = flow_smem_init<SB>(smem, A.f, true);
  const bool flip = (A.f.Lc & 1) != 0;
  const bool unit_mass = (A.imd == nullptr);
  float* smass = reinterpret_cast<float*>(S.mom - threadIdx.x + (size_t)E * kThreads);   // [d] inverse mass, PHYSICAL order
  if (!unit_mass) {
    for (int i = threadIdx.x; i < C.d; i += blockDim.x) smass[i] = __ldg(A.imd + (flip ? C.d - 1 - i : i));
    __syncthreads();
  }
  // conditioner stash (flow.cuh): Lc x cond_stash_floats x 128 floats behind the mass table, when the host found room
  float* stash = (SM && A.stash) ? smass + ((C.d + 3) & ~3) + threadIdx.x : nullptr;
  const int cpc = kThreads / C.gs;
  const long long tiles = (C.n + cpc - 1) / cpc;
  const float half_tau = A.tau / 2;
  unsigned int n_acc = 0, n_bad = 0;

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long chain_raw = tile * cpc + threadIdx.x / C.gs;
    const bool active = chain_raw < C.n;
    const long long chain = active ? chain_raw : C.n - 1;
    float* row = C.x + chain * (long long)C.d;

    float zlo[E], zhi[E];
    if (flip) load_chain_flipped(row, g, zlo, zhi); else load_chain(row, g, zlo, zhi);
#pragma unroll
    for (int e = 0; e < E; ++e) S.mom[e * kThreads] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int k = 0; k < C.n_steps; ++k) {
      float plo[E], phi[E];
      uint32_t ubits = 0;
      float kin0 = 0.f;
      {
        StepNoise<E> nz;
        if (C.rng.normals) {
          const float* nr = C.rng.normals + ((long long)k * C.n + chain) * (long long)C.d;
          if (flip) load_chain_flipped(nr, g, nz.lo, nz.hi); else load_chain(nr, g, nz.lo, nz.hi);
          nz.ubits = 0;
        } else {
          const RngKey key = make_rng_key(C.rng.seed, 0u, C.rng.step0 + (uint64_t)k, (uint64_t)(C.chain0 + chain));
          draw_step_noise<E>(key, g.j, nz);
        }
        ubits = nz.ubits;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int kk = g.j + g.gs * e;
          const bool vl = slot_ok<X, E>(e, kk, g.da), vh = slot_ok<X, E>(e, kk, g.db);
          float pl = vl ? nz.lo[e] : 0.f, ph = vh ? nz.hi[e] : 0.f;
          float ml = 1.f, mh = 1.f;
          if (!unit_mass) {                                                         // hmc.py:100
            ml = smass[vl ? kk : 0]; mh = smass[g.da + (vh ? kk : 0)];
            pl *= __fdiv_rn(1.f, sqrtf(ml));
            ph *= __fdiv_rn(1.f, sqrtf(mh));
          }
          plo[e] = pl; phi[e] = ph;
          kin0 = fmaf(pl * pl, ml, fmaf(ph * ph, mh, kin0));
        }
      }
      float glo[E], ghi[E];
      // leapfrog with a single gradient call site: iteration 0 only evaluates at the start point
      float u0 = 0.f, u1 = 0.f;
      for (int l = 0; l <= A.n_leapfrog; ++l) {
        if (l > 0) {
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const int kk = g.j + g.gs * e;
            const bool vl = slot_ok<X, E>(e, kk, g.da), vh = slot_ok<X, E>(e, kk, g.db);
            float ml = 1.f, mh = 1.f;
            if (!unit_mass) { ml = smass[vl ? kk : 0]; mh = smass[g.da + (vh ? kk : 0)]; }
            plo[e] = fmaf(-half_tau, glo[e], plo[e]);                               // hmc.py:51-53
            phi[e] = fmaf(-half_tau, ghi[e], phi[e]);
            zlo[e] = vl ? fmaf(A.tau, unit_mass ? plo[e] : plo[e] * ml, zlo[e]) : 0.f;   // hmc.py:56-58
            zhi[e] = vh ? fmaf(A.tau, unit_mass ? phi[e] : phi[e] * mh, zhi[e]) : 0.f;
          }
        }
        u1 = neutra_value_grad<E, SB, X, SM>(S.F, A.pot_kind, C.pot, g, zlo, zhi, glo, ghi, S.scr, true, stash);
        if (l == 0) u0 = u1;
        else {
#pragma unroll
          for (int e = 0; e < E; ++e) {
            plo[e] = fmaf(-half_tau, glo[e], plo[e]);
            phi[e] = fmaf(-half_tau, ghi[e], phi[e]);
          }
        }
      }
      float kin1 = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        float ml = 1.f, mh = 1.f;
        if (!unit_mass) {
          const int kk = g.j + g.gs * e;
          ml = smass[slot_ok<X, E>(e, kk, g.da) ? kk : 0]; mh = smass[g.da + (slot_ok<X, E>(e, kk, g.db) ? kk : 0)];
        }
        kin1 = fmaf(plo[e] * plo[e], ml, fmaf(phi[e] * phi[e], mh, kin1));
      }
      const float h0 = u0 + 0.5f * group_sum(kin0, g.gs);                           // hmc.py:103-106
      const float h1 = u1 + 0.5f * group_sum(kin1, g.gs);                           // hmc.py:107-110
      const float log_acc = -h1 - (-h0);
      float u;
      if (C.rng.uniforms) u = __ldg(C.rng.uniforms + (long long)k * C.n + chain);
      else u = uniform_from_bits(__shfl_sync(0xffffffffu, ubits, g.grp_base));
      const bool accept = logf(u) < log_acc;                                        // hmc.py:112-113
      if (!(fabsf(log_acc) <= 3.0e38f) && g.j == 0 && active) ++n_bad;
      // global memory always holds the current state: write the new one if accepted, re-read the old one if rejected
      if (accept) {
        if (active) { if (flip) store_chain_flipped(row, g, zlo, zhi); else store_chain(row, g, zlo, zhi); }
        if (g.j == 0 && active) ++n_acc;
      } else {
        if (flip) load_chain_flipped(row, g, zlo, zhi); else load_chain(row, g, zlo, zhi);
      }
#pragma unroll
      for (int e = 0; e < E; ++e) {
        float4 m = S.mom[e * kThreads];
        m.x += zlo[e]; m.y += zhi[e]; m.z = fmaf(zlo[e], zlo[e], m.z); m.w = fmaf(zhi[e], zhi[e], m.w);
        S.mom[e * kThreads] = m;
      }
      if (C.sink.samples && active) {
        const long long idx = C.sink.seen0 + k;
        if (idx % C.sink.thinning == 0) {
          const long long first = (C.sink.seen0 + C.sink.thinning - 1) / C.sink.thinning;
          float* dst = C.sink.samples + ((idx / C.sink.thinning - first) * C.n + chain) * (long long)C.d;
          if (flip) store_chain_flipped(dst, g, zlo, zhi); else store_chain(dst, g, zlo, zhi);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      float4 m = S.mom[e * kThreads];
      if (!active) m = make_float4(0.f, 0.f, 0.f, 0.f);
      const int kk = g.j + g.gs * e;
      const float a = across_groups_sum(m.x, g.gs), b = across_groups_sum(m.y, g.gs);
      const float c = across_groups_sum(m.z, g.gs), dd = across_groups_sum(m.w, g.gs);
      if (g.lane < g.gs) {
        const int il = flip ? g.d - 1 - kk : kk, ih = flip ? g.d - 1 - (g.da + kk) : g.da + kk;
        if (kk < g.da) { atomicAdd(S.st.sx + il, (double)a); atomicAdd(S.st.sx2 + il, (double)c); }
        if (kk < g.db) { atomicAdd(S.st.sx + ih, (double)b); atomicAdd(S.st.sx2 + ih, (double)dd); }
      }
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

template <int E, bool SB, bool X, bool SM>
__global__ void __launch_bounds__(kThreads) neutra_potential_kernel(FlowArgs FA, int pot_kind, PotParams P, const float* __restrict__ z,
                                                                   float* __restrict__ u, float* __restrict__ grad, long long n) {
  extern __shared__ __align__(16) unsigned char smem[];
  const Geom g = make_geom(FA.d, FA.gs);
  FlowSmem S = flow_smem_init<SB>(smem, FA, false);
  const bool flip = (FA.Lc & 1) != 0;
  const int cpc = kThreads / FA.gs;
  const long long tiles = (n + cpc - 1) / cpc;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long chain_raw = tile * cpc + threadIdx.x / FA.gs;
    const bool active = chain_raw < n;
    const long long chain = active ? chain_raw : n - 1;
    float zlo[E], zhi[E], glo[E], ghi[E];
    const float* src = z + chain * (long long)FA.d;
    if (flip) load_chain_flipped(src, g, zlo, zhi); else load_chain(src, g, zlo, zhi);
    const float v = neutra_value_grad<E, SB, X, SM>(S.F, pot_kind, P, g, zlo, zhi, glo, ghi, S.scr, grad != nullptr);
    if (active) {
      if (g.j == 0) u[chain] = v;
      if (grad) {
        float* dst = grad + chain * (long long)FA.d;
        if (flip) store_chain_flipped(dst, g, glo, ghi); else store_chain(dst, g, glo, ghi);
      }
    }
  }
}


// Pull a data-space cotangent back to the latent space for an EXTERNAL target (a Python callable differentiated by autograd,
// nfmc_b200/external.py): given z and gx = grad U(x) at x = T^-1(z), writes grad_z [ U(T^-1 z) - log|det dT^-1/dz| ] -- the
// gradient of NeuTra.adjusted_target (neutra.py:58-68) -- and, optionally, log|det dT^-1/dz| itself.  Same inverse pass and
// reversible backward sweep as neutra_value_grad, with the seed read from memory instead of pot_grad.
template <int E, bool SB, bool X, bool SM>
__global__ void __launch_bounds__(kThreads) neutra_pullback_kernel(FlowArgs FA, const float* __restrict__ z, const float* __restrict__ gx,
                                                                  float* __restrict__ gz, float* __restrict__ ld_out, long long n) {
  extern __shared__ __align__(16) unsigned char smem[];
  const Geom g = make_geom(FA.d, FA.gs);
  FlowSmem S = flow_smem_init<SB>(smem, FA, false);
  const bool flip = (FA.Lc & 1) != 0;
  const int cpc = kThreads / FA.gs;
  const long long tiles = (n + cpc - 1) / cpc;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long chain_raw = tile * cpc + threadIdx.x / FA.gs;
    const bool active = chain_raw < n;
    const long long chain = active ? chain_raw : n - 1;
    float xlo[E], xhi[E], glo[E], ghi[E];
    const float* src = z + chain * (long long)FA.d;
    if (flip) load_chain_flipped(src, g, xlo, xhi); else load_chain(src, g, xlo, xhi);
    const float ld_inv = flow_inverse<E, SB, X, SM>(S.F, g, xlo, xhi, S.scr, nullptr);
    load_chain(gx + chain * (long long)FA.d, g, glo, ghi);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int kk = g.j + g.gs * e;
      if (kk >= g.da) glo[e] = 0.f;
      if (kk >= g.db) ghi[e] = 0.f;
    }
    flow_unwind<E, SB, X, SM>(S.F, g, xlo, xhi, glo, ghi, S.scr, nullptr);
    if (active) {
      if (ld_out && g.j == 0) ld_out[chain] = ld_inv;
      float* dst = gz + chain * (long long)FA.d;
      if (flip) store_chain_flipped(dst, g, glo, ghi); else store_chain(dst, g, glo, ghi);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------
// NeuTra MH (nfmc/neutra.py:147-159): random-walk Metropolis (mcmc/mh.py:44-73) in the latent space on U~.  One inverse
// pass + one potential per step, no gradient; U~ of the current state is carried (the reference re-evaluates it, to the
// same value).  z' = z + inv_mass_diag * xi; accept iff log u < U~(z) - U~(z').
// ---------------------------------------------------------------------------------------------------------
template <int E, bool SB, bool X, bool SM>
__global__ void __launch_bounds__(kThreads, 3) neutra_mh_kernel(const NeutraArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ChainArgs& C = A.c;
  const Geom g = make_geom(C.d, C.gs);
  FlowSmem S = flow_smem_init<SB>(smem, A.f, true);
  const bool flip = (A.f.Lc & 1) != 0;
  const bool unit_mass = (A.imd == nullptr);
  float* smass = reinterpret_cast<float*>(S.mom - threadIdx.x + (size_t)E * kThreads);   // [d] proposal scale, PHYSICAL order
  if (!unit_mass) {
    for (int i = threadIdx.x; i < C.d; i += blockDim.x) smass[i] = __ldg(A.imd + (flip ? C.d - 1 - i : i));
    __syncthreads();
  }
  const int cpc = kThreads / C.gs;
  const long long tiles = (C.n + cpc - 1) / cpc;
  unsigned int n_acc = 0, n_bad = 0;

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long chain_raw = tile * cpc + threadIdx.x / C.gs;
    const bool active = chain_raw < C.n;
    const long long chain = active ? chain_raw : C.n - 1;
    float* row = C.x + chain * (long long)C.d;
    float zlo[E], zhi[E], dlo[E], dhi[E];
    if (flip) load_chain_flipped(row, g, zlo, zhi); else load_chain(row, g, zlo, zhi);
#pragma unroll
    for (int e = 0; e < E; ++e) S.mom[e * kThreads] = make_float4(0.f, 0.f, 0.f, 0.f);
    float u_cur = neutra_value_grad<E, SB, X, SM>(S.F, A.pot_kind, C.pot, g, zlo, zhi, dlo, dhi, S.scr, false);

    for (int k = 0; k < C.n_steps; ++k) {
      StepNoise<E> nz;
      if (C.rng.normals) {
        const float* nr = C.rng.normals + ((long long)k * C.n + chain) * (long long)C.d;
        if (flip) load_chain_flipped(nr, g, nz.lo, nz.hi); else load_chain(nr, g, nz.lo, nz.hi);
        nz.ubits = 0;
      } else {
        const RngKey key = make_rng_key(C.rng.seed, 0u, C.rng.step0 + (uint64_t)k, (uint64_t)(C.chain0 + chain));
        draw_step_noise<E>(key, g.j, nz);
      }
      float plo[E], phi[E];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int kk = g.j + g.gs * e;
        const bool vl = slot_ok<X, E>(e, kk, g.da), vh = slot_ok<X, E>(e, kk, g.db);
        const float ml = unit_mass ? 1.f : smass[vl ? kk : 0], mh = unit_mass ? 1.f : smass[g.da + (vh ? kk : 0)];
        plo[e] = vl ? zlo[e] + nz.lo[e] * ml : 0.f;                                  // mh.py:52-56
        phi[e] = vh ? zhi[e] + nz.hi[e] * mh : 0.f;
      }
      bool accept = true;
      float u_p = u_cur;
      if (A.adjusted) {
        u_p = neutra_value_grad<E, SB, X, SM>(S.F, A.pot_kind, C.pot, g, plo, phi, dlo, dhi, S.scr, false);
        const float log_ratio = (-u_p) - (-u_cur) + 0.f - 0.f;                       // mh.py:59
        float u;
        if (C.rng.uniforms) u = __ldg(C.rng.uniforms + (long long)k * C.n + chain);
        else u = uniform_from_bits(__shfl_sync(0xffffffffu, nz.ubits, g.grp_base));
        accept = logf(u) < log_ratio;                                                // mh.py:60
        if (!(fabsf(log_ratio) <= 3.0e38f) && g.j == 0 && active) ++n_bad;
      }
#pragma unroll
      for (int e = 0; e < E; ++e) {
        zlo[e] = accept ? plo[e] : zlo[e];
        zhi[e] = accept ? phi[e] : zhi[e];
        float4 m = S.mom[e * kThreads];
        m.x += zlo[e]; m.y += zhi[e]; m.z = fmaf(zlo[e], zlo[e], m.z); m.w = fmaf(zhi[e], zhi[e], m.w);
        S.mom[e * kThreads] = m;
      }
      u_cur = accept ? u_p : u_cur;
      if (accept && g.j == 0 && active) ++n_acc;
      if (C.sink.samples && active) {
        const long long idx = C.sink.seen0 + k;
        if (idx % C.sink.thinning == 0) {
          const long long first = (C.sink.seen0 + C.sink.thinning - 1) / C.sink.thinning;
          float* dst = C.sink.samples + ((idx / C.sink.thinning - first) * C.n + chain) * (long long)C.d;
          if (flip) store_chain_flipped(dst, g, zlo, zhi); else store_chain(dst, g, zlo, zhi);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      float4 m = S.mom[e * kThreads];
      if (!active) m = make_float4(0.f, 0.f, 0.f, 0.f);
      const int kk = g.j + g.gs * e;
      const float a = across_groups_sum(m.x, g.gs), b = across_groups_sum(m.y, g.gs);
      const float c = across_groups_sum(m.z, g.gs), dd = across_groups_sum(m.w, g.gs);
      if (g.lane < g.gs) {
        const int il = flip ? g.d - 1 - kk : kk, ih = flip ? g.d - 1 - (g.da + kk) : g.da + kk;
        if (kk < g.da) { atomicAdd(S.st.sx + il, (double)a); atomicAdd(S.st.sx2 + il, (double)c); }
        if (kk < g.db) { atomicAdd(S.st.sx + ih, (double)b); atomicAdd(S.st.sx2 + ih, (double)dd); }
      }
    }
    if (active) { if (flip) store_chain_flipped(row, g, zlo, zhi); else store_chain(row, g, zlo, zhi); }
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

template <int E>
int launch_neutra_hmc(const NeutraArgs& A, int grid, size_t smem, cudaStream_t s) {
  // exact-layout specialisations are instantiated for the production shapes (E = 13, 16) only; the generic
  // (non-small) conditioner path is a separate, single variant per blob placement
  constexpr bool XE = (E == 13 || E == 16);
  const bool small = flow_is_small(A.f.M, A.f.H);
  const bool xl = A.f.exact && XE;
#define NFMC_LAUNCH(SBv, Xv, Sv)                                                              \
  do {                                                                                        \
    NFMC_SET_SMEM_RET((neutra_hmc_kernel<E, SBv, Xv, Sv>), smem);                                        \
    neutra_hmc_kernel<E, SBv, Xv, Sv><<<occupancy_grid(neutra_hmc_kernel<E, SBv, Xv, Sv>, smem, A.c.n, A.c.gs), kThreads, smem, s>>>(A);                              \
  } while (0)
  if (!small) { if (A.f.stage_blob) NFMC_LAUNCH(true, false, false); else NFMC_LAUNCH(false, false, false); }
  else if (A.f.stage_blob && xl) NFMC_LAUNCH(true, XE, true);
  else if (A.f.stage_blob) NFMC_LAUNCH(true, false, true);
  else if (xl) NFMC_LAUNCH(false, XE, true);
  else NFMC_LAUNCH(false, false, true);
#undef NFMC_LAUNCH
  return check_cuda(cudaGetLastError(), "neutra_hmc_kernel launch");
}
template <int E>
int launch_neutra_potential(const FlowArgs& FA, int pot_kind, const PotParams& P, const float* z, float* u, float* grad,
                            long long n, int grid, size_t smem, cudaStream_t s) {
  // exact-layout specialisations are instantiated for the production shapes (E = 13, 16) only; the generic
  // (non-small) conditioner path is a separate, single variant per blob placement
  constexpr bool XE = (E == 13 || E == 16);
  const bool small = flow_is_small(FA.M, FA.H);
  const bool xl = FA.exact && XE;
#define NFMC_LAUNCH(SBv, Xv, Sv)                                                              \
  do {                                                                                        \
    NFMC_SET_SMEM_RET((neutra_potential_kernel<E, SBv, Xv, Sv>), smem);                                        \
    neutra_potential_kernel<E, SBv, Xv, Sv><<<occupancy_grid(neutra_potential_kernel<E, SBv, Xv, Sv>, smem, n, FA.gs), kThreads, smem, s>>>(FA, pot_kind, P, z, u, grad, n);                              \
  } while (0)
  if (!small) { if (FA.stage_blob) NFMC_LAUNCH(true, false, false); else NFMC_LAUNCH(false, false, false); }
  else if (FA.stage_blob && xl) NFMC_LAUNCH(true, XE, true);
  else if (FA.stage_blob) NFMC_LAUNCH(true, false, true);
  else if (xl) NFMC_LAUNCH(false, XE, true);
  else NFMC_LAUNCH(false, false, true);
#undef NFMC_LAUNCH
  return check_cuda(cudaGetLastError(), "neutra_potential_kernel launch");
}
template int launch_neutra_hmc<NFMC_ONLY_E>(const NeutraArgs&, int, size_t, cudaStream_t);

template <int E>
int launch_neutra_mh(const NeutraArgs& A, int grid, size_t smem, cudaStream_t s) {
  constexpr bool XE = (E == 13 || E == 16);
  const bool small = flow_is_small(A.f.M, A.f.H);
  const bool xl = A.f.exact && XE;
#define NFMC_LAUNCH(SBv, Xv, Sv)                                                   \
  do {                                                                             \
    NFMC_SET_SMEM_RET((neutra_mh_kernel<E, SBv, Xv, Sv>), smem);                   \
    neutra_mh_kernel<E, SBv, Xv, Sv><<<occupancy_grid(neutra_mh_kernel<E, SBv, Xv, Sv>, smem, A.c.n, A.c.gs), kThreads, smem, s>>>(A);              \
  } while (0)
  if (!small) { if (A.f.stage_blob) NFMC_LAUNCH(true, false, false); else NFMC_LAUNCH(false, false, false); }
  else if (A.f.stage_blob && xl) NFMC_LAUNCH(true, XE, true);
  else if (A.f.stage_blob) NFMC_LAUNCH(true, false, true);
  else if (xl) NFMC_LAUNCH(false, XE, true);
  else NFMC_LAUNCH(false, false, true);
#undef NFMC_LAUNCH
  return check_cuda(cudaGetLastError(), "neutra_mh_kernel launch");
}
template int launch_neutra_mh<NFMC_ONLY_E>(const NeutraArgs&, int, size_t, cudaStream_t);
template int launch_neutra_potential<NFMC_ONLY_E>(const FlowArgs&, int, const PotParams&, const float*, float*, float*, long long, int, size_t, cudaStream_t);

template <int E>
int launch_neutra_pullback(const FlowArgs& FA, const float* z, const float* gx, float* gz, float* ld, long long n, int grid,
                           size_t smem, cudaStream_t s) {
  constexpr bool XE = (E == 13 || E == 16);
  const bool small = flow_is_small(FA.M, FA.H);
  const bool xl = FA.exact && XE;
#define NFMC_LAUNCH(SBv, Xv, Sv)                                                              \
  do {                                                                                        \
    NFMC_SET_SMEM_RET((neutra_pullback_kernel<E, SBv, Xv, Sv>), smem);                                        \
    neutra_pullback_kernel<E, SBv, Xv, Sv><<<occupancy_grid(neutra_pullback_kernel<E, SBv, Xv, Sv>, smem, n, FA.gs), kThreads, smem, s>>>(FA, z, gx, gz, ld, n);                              \
  } while (0)
  if (!small) { if (FA.stage_blob) NFMC_LAUNCH(true, false, false); else NFMC_LAUNCH(false, false, false); }
  else if (FA.stage_blob && xl) NFMC_LAUNCH(true, XE, true);
  else if (FA.stage_blob) NFMC_LAUNCH(true, false, true);
  else if (xl) NFMC_LAUNCH(false, XE, true);
  else NFMC_LAUNCH(false, false, true);
#undef NFMC_LAUNCH
  return check_cuda(cudaGetLastError(), "neutra_pullback_kernel launch");
}
template int launch_neutra_pullback<NFMC_ONLY_E>(const FlowArgs&, const float*, const float*, float*, float*, long long, int, size_t, cudaStream_t);

}  // namespace nfmc
