// potentials.cuh -- analytic target potentials U and grad U on the register-resident chain layout.
//
// The reference evaluates `target(x)` and autograd of it (mcmc/langevin.py:66-68,80-82; mcmc/hmc.py:40-48);
// here each built-in potential has a closed-form value and gradient.  A potential is split into
//   pot_prepare : one pass over the chain + group reductions -> PotCtx (U and a few chain-level scalars)
//   pot_grad    : elementwise gradient of slot e's (lo, hi) pair from PotCtx
// so the gradient vector never has to be stored.
#pragma once
#include "common.cuh"
#include "../../include/nfmc_b200.h"

namespace nfmc {

struct PotParams {
  const float* params;  // per-dimension parameters (device) or nullptr
  float s0, s1, s2, s3;
};

struct PotCtx {
  float u;        // U(x)
  float a, b, c;  // potential-specific chain-level scalars
};

__device__ __forceinline__ PotCtx select_ctx(bool take_new, const PotCtx& n, const PotCtx& o) {
  PotCtx r;
  r.u = take_new ? n.u : o.u;
  r.a = take_new ? n.a : o.a;
  r.b = take_new ? n.b : o.b;
  r.c = take_new ? n.c : o.c;
  return r;
}

// value of physical element i (i in {0, 1}) broadcast to every lane of the group
template <int E>
__device__ __forceinline__ float elem_bcast(const float (&lo)[E], const float (&hi)[E], const Geom& g, int i) {
  const bool in_hi = i >= g.da;
  const int k = in_hi ? i - g.da : i;
  const int owner = k & (g.gs - 1);
  const int slot = k / g.gs;  // 0 or 1
  const float v0 = in_hi ? hi[0] : lo[0];
  const float v1 = in_hi ? hi[E > 1 ? 1 : 0] : lo[E > 1 ? 1 : 0];
  return group_bcast(slot == 0 ? v0 : v1, g, owner);
}

template <int POT, int E>
__device__ __forceinline__ PotCtx pot_prepare(const PotParams& P, const Geom& g, const float (&lo)[E], const float (&hi)[E]) {
  PotCtx c;
  c.a = c.b = c.c = 0.f;
  if constexpr (POT == NFMC_POT_ISO_GAUSSIAN) {
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) s = fmaf(lo[e], lo[e], fmaf(hi[e], hi[e], s));  // invalid slots hold 0
    c.u = 0.5f * P.s0 * group_sum(s, g.gs);
  } else if constexpr (POT == NFMC_POT_DIAG_GAUSSIAN) {
    const float2* wm = reinterpret_cast<const float2*>(P.params);
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int k = g.j + g.gs * e;
      if (k < g.da) { const float2 p = __ldg(wm + k); const float t = lo[e] - p.y; s = fmaf(p.x * t, t, s); }
      if (k < g.db) { const float2 p = __ldg(wm + g.da + k); const float t = hi[e] - p.y; s = fmaf(p.x * t, t, s); }
    }
    c.u = 0.5f * group_sum(s, g.gs);
  } else if constexpr (POT == NFMC_POT_FUNNEL) {
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) s = fmaf(lo[e], lo[e], fmaf(hi[e], hi[e], s));
    const float x0 = elem_bcast(lo, hi, g, 0);
    const float S = group_sum(s, g.gs) - x0 * x0;
    const float ex = __expf(-x0);
    c.a = x0; c.b = ex; c.c = S;
    c.u = x0 * x0 * P.s1 + 0.5f * (float)(g.d - 1) * x0 + 0.5f * ex * S;  // P.s1 = 1/(2 s^2)
  } else if constexpr (POT == NFMC_POT_ROSENBROCK) {
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int k = g.j + g.gs * e;
      if (k < g.da) {
        const float t = lo[e] - 1.f, r = hi[e] - lo[e] * lo[e];
        s += t * t + P.s0 * r * r;
      }
    }
    c.u = group_sum(s, g.gs);
  } else if constexpr (POT == NFMC_POT_MIXTURE4) {
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) s = fmaf(lo[e], lo[e], fmaf(hi[e], hi[e], s));
    const float S = group_sum(s, g.gs);
    const float x0 = elem_bcast(lo, hi, g, 0), x1 = elem_bcast(lo, hi, g, 1);
    const float a = P.s0;
    const float base = -0.5f * (S + 2.f * a * a);
    const float e0 = base + a * (x0 + x1), e1 = base + a * (x0 - x1), e2 = base + a * (-x0 + x1), e3 = base + a * (-x0 - x1);
    const float m = fmaxf(fmaxf(e0, e1), fmaxf(e2, e3));
    const float w0 = __expf(e0 - m), w1 = __expf(e1 - m), w2 = __expf(e2 - m), w3 = __expf(e3 - m);
    const float z = w0 + w1 + w2 + w3;
    c.u = -(m + __logf(z));
    const float rz = 1.f / z;
    c.a = a * (w0 + w1 - w2 - w3) * rz;  // sum_k r_k mu_k[0]
    c.b = a * (w0 - w1 + w2 - w3) * rz;  // sum_k r_k mu_k[1]
  }
  return c;
}

// gradient of slot e: glo = dU/dx[k], ghi = dU/dx[da + k]  (k = j + gs*e); values at invalid slots are ignored.
// KV: the caller knows the slot is valid in both halves (no clamping of parameter indices needed).
template <int POT, bool KV = false>
__device__ __forceinline__ void pot_grad(const PotParams& P, const PotCtx& c, const Geom& g, int k, float vlo, float vhi,
                                         float& glo, float& ghi) {
  if constexpr (POT == NFMC_POT_ISO_GAUSSIAN) {
    glo = P.s0 * vlo;
    ghi = P.s0 * vhi;
  } else if constexpr (POT == NFMC_POT_DIAG_GAUSSIAN) {
    const float2* wm = reinterpret_cast<const float2*>(P.params);
    const int kl = KV ? k : min(k, g.da > 0 ? g.da - 1 : 0);
    const int kh = g.da + (KV ? k : min(k, g.db - 1));
    const float2 pl = __ldg(wm + kl), ph = __ldg(wm + kh);
    glo = pl.x * (vlo - pl.y);
    ghi = ph.x * (vhi - ph.y);
  } else if constexpr (POT == NFMC_POT_FUNNEL) {
    glo = c.b * vlo;
    ghi = c.b * vhi;
    const float g0 = 2.f * P.s1 * c.a + 0.5f * (float)(g.d - 1) - 0.5f * c.b * c.c;
    if (k == 0) { if (g.da > 0) glo = g0; else ghi = g0; }
  } else if constexpr (POT == NFMC_POT_ROSENBROCK) {
    const float r = vhi - vlo * vlo;
    glo = 2.f * (vlo - 1.f) - 4.f * P.s0 * vlo * r;
    ghi = 2.f * P.s0 * r;
  } else if constexpr (POT == NFMC_POT_MIXTURE4) {
    glo = vlo;
    ghi = vhi;
    // elements 0 and 1 carry the mode offsets
    if (k == 0) { if (g.da > 0) glo = vlo - c.a; else ghi = vhi - c.a; }
    if (g.da >= 2) { if (k == 1) glo = vlo - c.b; }
    else if (g.da == 1) { if (k == 0) ghi = vhi - c.b; }
    else { if (k == 1) ghi = vhi - c.b; }
  }
}

// does grad U need the chain-level scalars of pot_prepare (i.e. a reduction over the chain)?
template <int POT>
constexpr bool pot_grad_needs_ctx() { return POT == NFMC_POT_FUNNEL || POT == NFMC_POT_MIXTURE4; }

// runtime-dispatched versions for the flow-heavy kernels (jump / IMH / NeuTra), where the potential is a small
// part of the work and templating on it would only multiply compile time
template <int E>
__device__ __forceinline__ PotCtx pot_prepare_rt(int kind, const PotParams& P, const Geom& g, const float (&lo)[E], const float (&hi)[E]) {
  switch (kind) {
    case NFMC_POT_ISO_GAUSSIAN: return pot_prepare<NFMC_POT_ISO_GAUSSIAN, E>(P, g, lo, hi);
    case NFMC_POT_DIAG_GAUSSIAN: return pot_prepare<NFMC_POT_DIAG_GAUSSIAN, E>(P, g, lo, hi);
    case NFMC_POT_FUNNEL: return pot_prepare<NFMC_POT_FUNNEL, E>(P, g, lo, hi);
    case NFMC_POT_ROSENBROCK: return pot_prepare<NFMC_POT_ROSENBROCK, E>(P, g, lo, hi);
    default: return pot_prepare<NFMC_POT_MIXTURE4, E>(P, g, lo, hi);
  }
}
__device__ __forceinline__ void pot_grad_rt(int kind, const PotParams& P, const PotCtx& c, const Geom& g, int k, float vlo,
                                            float vhi, float& glo, float& ghi) {
  switch (kind) {
    case NFMC_POT_ISO_GAUSSIAN: pot_grad<NFMC_POT_ISO_GAUSSIAN>(P, c, g, k, vlo, vhi, glo, ghi); break;
    case NFMC_POT_DIAG_GAUSSIAN: pot_grad<NFMC_POT_DIAG_GAUSSIAN>(P, c, g, k, vlo, vhi, glo, ghi); break;
    case NFMC_POT_FUNNEL: pot_grad<NFMC_POT_FUNNEL>(P, c, g, k, vlo, vhi, glo, ghi); break;
    case NFMC_POT_ROSENBROCK: pot_grad<NFMC_POT_ROSENBROCK>(P, c, g, k, vlo, vhi, glo, ghi); break;
    default: pot_grad<NFMC_POT_MIXTURE4>(P, c, g, k, vlo, vhi, glo, ghi); break;
  }
}

#define NFMC_DISPATCH_POT(kind, ...)                                                            \
  switch (kind) {                                                                               \
    case NFMC_POT_ISO_GAUSSIAN: { constexpr int POT = NFMC_POT_ISO_GAUSSIAN; __VA_ARGS__; } break;   \
    case NFMC_POT_DIAG_GAUSSIAN: { constexpr int POT = NFMC_POT_DIAG_GAUSSIAN; __VA_ARGS__; } break; \
    case NFMC_POT_FUNNEL: { constexpr int POT = NFMC_POT_FUNNEL; __VA_ARGS__; } break;               \
    case NFMC_POT_ROSENBROCK: { constexpr int POT = NFMC_POT_ROSENBROCK; __VA_ARGS__; } break;       \
    case NFMC_POT_MIXTURE4: { constexpr int POT = NFMC_POT_MIXTURE4; __VA_ARGS__; } break;           \
    default: return set_error("unknown potential kind");                                        \
  }

#define NFMC_DISPATCH_E(E_rt, ...)                                    \
  switch (E_rt) {                                                     \
    case 4: { constexpr int E = 4; __VA_ARGS__; } break;              \
    case 7: { constexpr int E = 7; __VA_ARGS__; } break;              \
    case 13: { constexpr int E = 13; __VA_ARGS__; } break;            \
    case 16: { constexpr int E = 16; __VA_ARGS__; } break;            \
    default: return set_error("unsupported slots-per-half");          \
  }

}  // namespace nfmc
