// flow.cuh -- RealNVP forward / inverse / log-det / input-VJP on the register-resident chain layout.
//
// Specification = oracle/realnvp_ref.py (the torchflows surface the reference calls: sampling/base.py:26,
// util.py:280-281, jump.py:205,218, imh.py:214,221, neutra.py:60).  Generic CUDA-core path: any number of
// coupling layers Lc, conditioner linear layers M >= 1 and hidden width H.  (Wide conditioners go to the
// tcgen05 path in cond_tc.cu.)
//
// Physical coordinates.  The flow's reverse permutations are folded into the packed parameters: the chain
// state is never permuted; instead coupling l (0-based) has its source/target halves swapped when l is even
// (an odd number of reversals precede it) and every parameter is stored in the coordinate the state actually
// lives in.  Only the latent z of a flow with odd Lc is flipped, on load/store.
//
// Blob layout (fp32), da = d/2, db = d - da, Hs = (M == 1 ? da : H):
//   [ (Lc+3) elementwise affines ] each 3*d: alpha[d], beta[d], 1/alpha[d]      (order: A0, AN_0..AN_{Lc-1}, A_last, AN_last)
//   [ 4 floats ] [0] = sum over all elementwise affines of sum_i log alpha_i
//   [ Lc couplings ] each:
//        M >= 2:  W1[H][da] b1[H] | (M-2) x { Wm[H_in][H_out] bm[H] } | Wl[H][2][db] bl[2][db]
//        M == 1:  Wl[da][2][db] bl[2][db]
//   (W1 indexed by packed source index, Wl by packed target index; see pack_realnvp in nfmc_b200/flow.py)
#pragma once
#include "common.cuh"

namespace nfmc {

struct FlowDesc {
  const float* blob;
  int d, da, db, Lc, M, H;
  int Hs;               // width of the last layer's input
  int off_const;        // (Lc+3)*3*d
  int off_coupling;     // off_const + 4
  int coupling_stride;  // floats per coupling
  int scratch;          // floats of scratch per chain group
};

__host__ __device__ inline int flow_coupling_floats(int d, int M, int H) {
  const int da = d / 2, db = d - da;
  if (M == 1) return da * 2 * db + 2 * db;
  return (da * H + H) + (M - 2) * (H * H + H) + (H * 2 * db + 2 * db);
}
__host__ __device__ inline long long flow_blob_floats(int d, int Lc, int M, int H) {
  return (long long)(Lc + 3) * 3 * d + 4 + (long long)Lc * flow_coupling_floats(d, M, H);
}
__host__ __device__ inline FlowDesc make_flow_desc(const float* blob, int d, int Lc, int M, int H) {
  FlowDesc F;
  F.blob = blob; F.d = d; F.da = d / 2; F.db = d - F.da; F.Lc = Lc; F.M = M; F.H = H;
  F.Hs = (M == 1) ? F.da : H;
  F.off_const = (Lc + 3) * 3 * d;
  F.off_coupling = F.off_const + 4;
  F.coupling_stride = flow_coupling_floats(d, M, H);
  F.scratch = ((M - 1 > 1 ? M - 1 : 1) + 2) * F.Hs;
  return F;
}

__device__ __forceinline__ float precise_tanh(float v) { return tanhf(v); }

// ---- elementwise affine (A0, act-norms, A_last) ---------------------------------------------------------
template <int E, bool INV>
__device__ __forceinline__ void affine_apply(const FlowDesc& F, const Geom& g, int a, float (&lo)[E], float (&hi)[E]) {
  const float* al = F.blob + (long long)a * 3 * F.d;
  const float* be = al + F.d;
  const float* ra = be + F.d;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int k = g.j + g.gs * e;
    if (k < g.da) lo[e] = INV ? (lo[e] - be[k]) * ra[k] : fmaf(al[k], lo[e], be[k]);
    if (k < g.db) hi[e] = INV ? (hi[e] - be[g.da + k]) * ra[g.da + k] : fmaf(al[g.da + k], hi[e], be[g.da + k]);
  }
}
// undo an inverse affine (state <- alpha*state + beta) while pulling the gradient back (grad <- grad / alpha)
template <int E>
__device__ __forceinline__ void affine_unwind(const FlowDesc& F, const Geom& g, int a, float (&lo)[E], float (&hi)[E],
                                              float (&glo)[E], float (&ghi)[E]) {
  const float* al = F.blob + (long long)a * 3 * F.d;
  const float* be = al + F.d;
  const float* ra = be + F.d;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int k = g.j + g.gs * e;
    if (k < g.da) { lo[e] = fmaf(al[k], lo[e], be[k]); glo[e] *= ra[k]; }
    if (k < g.db) { hi[e] = fmaf(al[g.da + k], hi[e], be[g.da + k]); ghi[e] *= ra[g.da + k]; }
  }
}

// ---- conditioner MLP ---------------------------------------------------------------------------------------
// src[e] holds source element ks = j + gs*e - shift (valid iff 0 <= ks < da).  Targets: slot e of the other half
// has packed index t = j + gs*e (valid iff t < nt_main), plus, when has_x, one extra target t = da (the middle
// element of an odd-d chain, which lives in hi[0] of lane 0).  Hidden activations stay in scr for the backward.
template <int E>
__device__ __forceinline__ void cond_forward(const FlowDesc& F, const Geom& g, int l, const float (&src)[E], int shift,
                                             int nt_main, bool has_x, float* scr, float (&ua)[E], float (&ub)[E],
                                             float& ua_x, float& ub_x) {
  const float* W = F.blob + F.off_coupling + (long long)l * F.coupling_stride;
  const int da = F.da, db = F.db, H = F.H;
  __syncwarp();
  const float* last;
  const float* Wl;
  if (F.M >= 2) {
    const float* W1 = W;
    const float* b1 = W + H * da;
    for (int h0 = 0; h0 < H; h0 += 4) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int ks = g.j + g.gs * e - shift;
        if (ks >= 0 && ks < da) {
          const float v = src[e];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (h0 + u < H) acc[u] = fmaf(W1[(h0 + u) * da + ks], v, acc[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float s = group_sum(acc[u], g.gs);
        const int h = h0 + u;
        if (h < H && g.j == (h & (g.gs - 1))) scr[h] = precise_tanh(s + b1[h]);
      }
    }
    __syncwarp();
    const float* Wm = b1 + H;
    for (int m = 1; m <= F.M - 2; ++m) {
      const float* in = scr + (m - 1) * H;
      float* out = scr + m * H;
      const float* bm = Wm + H * H;
      for (int hq = g.j; hq < H; hq += g.gs) {
        float acc = bm[hq];
        for (int h = 0; h < H; ++h) acc = fmaf(Wm[h * H + hq], in[h], acc);
        out[hq] = precise_tanh(acc);
      }
      __syncwarp();
      Wm = bm + H;
    }
    last = scr + (F.M - 2) * H;
    Wl = Wm;
  } else {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int ks = g.j + g.gs * e - shift;
      if (ks >= 0 && ks < da) scr[ks] = src[e];
    }
    __syncwarp();
    last = scr;
    Wl = W;
  }
  const int Hs = F.Hs;
  const float* bl = Wl + (long long)Hs * 2 * db;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int t = g.j + g.gs * e;
    const int tc = t < nt_main ? t : 0;
    ua[e] = bl[tc];
    ub[e] = bl[db + tc];
  }
  ua_x = has_x ? bl[da] : 0.f;
  ub_x = has_x ? bl[db + da] : 0.f;
  for (int h = 0; h < Hs; ++h) {
    const float hv = last[h];
    const float* wa = Wl + (long long)h * 2 * db;
    const float* wb = wa + db;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int t = g.j + g.gs * e;
      const int tc = t < nt_main ? t : 0;
      ua[e] = fmaf(wa[tc], hv, ua[e]);
      ub[e] = fmaf(wb[tc], hv, ub[e]);
    }
    if (has_x) {
      ua_x = fmaf(wa[da], hv, ua_x);
      ub_x = fmaf(wb[da], hv, ub_x);
    }
  }
}

// input-VJP of the conditioner: dsrc[e] += sum_out d_out * d out / d src[e].  Needs scr as left by cond_forward.
template <int E>
__device__ __forceinline__ void cond_backward(const FlowDesc& F, const Geom& g, int l, int shift, int nt_main, bool has_x,
                                              float* scr, const float (&dua)[E], const float (&dub)[E], float dua_x,
                                              float dub_x, float (&dsrc)[E]) {
  const float* W = F.blob + F.off_coupling + (long long)l * F.coupling_stride;
  const int da = F.da, db = F.db, H = F.H, Hs = F.Hs;
  float* gA = scr + (F.M - 1 > 1 ? F.M - 1 : 1) * Hs;
  float* gB = gA + Hs;
  __syncwarp();
  if (F.M >= 2) {
    const float* W1 = W;
    const float* Wm_first = W + H * da + H;
    const float* Wl = Wm_first + (long long)(F.M - 2) * (H * H + H);
    const float* act_last = scr + (F.M - 2) * H;
    for (int h0 = 0; h0 < H; h0 += 4) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int t = g.j + g.gs * e;
        if (t < nt_main) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (h0 + u < H) {
              const float* wa = Wl + (long long)(h0 + u) * 2 * db;
              acc[u] = fmaf(wa[t], dua[e], fmaf(wa[db + t], dub[e], acc[u]));
            }
        }
      }
      if (has_x && g.j == 0) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (h0 + u < H) {
            const float* wa = Wl + (long long)(h0 + u) * 2 * db;
            acc[u] = fmaf(wa[da], dua_x, fmaf(wa[db + da], dub_x, acc[u]));
          }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float s = group_sum(acc[u], g.gs);
        const int h = h0 + u;
        if (h < H && g.j == (h & (g.gs - 1))) { const float a = act_last[h]; gA[h] = s * (1.f - a * a); }
      }
    }
    __syncwarp();
    for (int m = F.M - 2; m >= 1; --m) {
      const float* Wm = Wm_first + (long long)(m - 1) * (H * H + H);
      const float* act_in = scr + (m - 1) * H;
      for (int hq = g.j; hq < H; hq += g.gs) {
        float acc = 0.f;
        for (int h2 = 0; h2 < H; ++h2) acc = fmaf(Wm[hq * H + h2], gA[h2], acc);
        const float a = act_in[hq];
        gB[hq] = acc * (1.f - a * a);
      }
      __syncwarp();
      float* tmp = gA; gA = gB; gB = tmp;
    }
    for (int h = 0; h < H; ++h) {
      const float gv = gA[h];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int ks = g.j + g.gs * e - shift;
        if (ks >= 0 && ks < da) dsrc[e] = fmaf(W1[h * da + ks], gv, dsrc[e]);
      }
    }
  } else {
    const float* Wl = W;
    for (int ks0 = 0; ks0 < da; ++ks0) {
      const float* wa = Wl + (long long)ks0 * 2 * db;
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int t = g.j + g.gs * e;
        if (t < nt_main) acc = fmaf(wa[t], dua[e], fmaf(wa[db + t], dub[e], acc));
      }
      if (has_x && g.j == 0) acc = fmaf(wa[da], dua_x, fmaf(wa[db + da], dub_x, acc));
      acc = group_sum(acc, g.gs);
#pragma unroll
      for (int e = 0; e < E; ++e)
        if (g.j + g.gs * e - shift == ks0) dsrc[e] += acc;
    }
  }
}

__device__ __forceinline__ void affine_coef(float ua, float ub, float& alpha, float& beta) {
  alpha = __expf(kLogOneMinusM + 0.5f * ua) + kMinScale;
  beta = 0.5f * ub;
}

template <int E>
__device__ __forceinline__ void swap_halves(float (&a)[E], float (&b)[E], bool doit) {
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const float t = a[e];
    a[e] = doit ? b[e] : t;
    b[e] = doit ? t : b[e];
  }
}

// ---- affine coupling layer l; INV = false: x -> z direction, true: z -> x.  ld accumulates THIS LANE's share of
//      sum_t log alpha_t (caller group-sums once per pass).  When the source is the high half the two register
//      arrays are swapped around the layer so that there is a single conditioner call site.
template <int E, bool INV>
__device__ __forceinline__ void coupling_apply(const FlowDesc& F, const Geom& g, int l, float (&lo)[E], float (&hi)[E],
                                               float* scr, float& ld) {
  float ua[E], ub[E], ua_x, ub_x;
  const bool src_is_hi = (l & 1) == 0;
  const int shift = src_is_hi ? F.db - F.da : 0;
  const bool has_x = shift > 0;
  const int nt_main = src_is_hi ? F.da : F.db;
  swap_halves(lo, hi, src_is_hi);  // lo = source array, hi = target array
  cond_forward<E>(F, g, l, lo, shift, nt_main, has_x, scr, ua, ub, ua_x, ub_x);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    if (g.j + g.gs * e < nt_main) {
      float al, be;
      affine_coef(ua[e], ub[e], al, be);
      hi[e] = INV ? __fdividef(hi[e] - be, al) : fmaf(al, hi[e], be);
      ld += __logf(al);
    }
  }
  if (has_x && g.j == 0) {  // middle element of an odd-d chain: physically hi[0], here slot 0 of the source array
    float al, be;
    affine_coef(ua_x, ub_x, al, be);
    lo[0] = INV ? __fdividef(lo[0] - be, al) : fmaf(al, lo[0], be);
    ld += __logf(al);
  }
  swap_halves(lo, hi, src_is_hi);
}

// Undo inverse-coupling l: (lo, hi) currently hold the coupling's OUTPUT of the z->x pass (a, b) and (glo, ghi)
// the gradient of U~ with respect to it.  On return they hold its INPUT (a, b') and the gradient with respect to
// that input, where U~ = U(x) - log|det dx/dz| so each coupling contributes + sum log alpha to U~:
//   b = (b' - beta)/alpha  =>  dU~/dalpha = (1 - gb*b)/alpha,  dU~/dbeta = -gb/alpha,  dU~/db' = gb/alpha.
template <int E>
__device__ __forceinline__ void coupling_unwind(const FlowDesc& F, const Geom& g, int l, float (&lo)[E], float (&hi)[E],
                                                float (&glo)[E], float (&ghi)[E], float* scr) {
  float ua[E], ub[E], ua_x = 0.f, ub_x = 0.f;
  float dua_x = 0.f, dub_x = 0.f;
  const bool src_is_hi = (l & 1) == 0;
  const int shift = src_is_hi ? F.db - F.da : 0;
  const bool has_x = shift > 0;
  const int nt_main = src_is_hi ? F.da : F.db;
  swap_halves(lo, hi, src_is_hi);
  swap_halves(glo, ghi, src_is_hi);
  cond_forward<E>(F, g, l, lo, shift, nt_main, has_x, scr, ua, ub, ua_x, ub_x);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    float da_ = 0.f, db_ = 0.f;
    if (g.j + g.gs * e < nt_main) {
      float al, be;
      affine_coef(ua[e], ub[e], al, be);
      const float ra = __fdividef(1.f, al);
      da_ = (1.f - ghi[e] * hi[e]) * (al - kMinScale) * 0.5f * ra;
      db_ = -0.5f * ghi[e] * ra;
      hi[e] = fmaf(al, hi[e], be);
      ghi[e] *= ra;
    }
    ua[e] = da_;  // reuse as d/d u_a, d/d u_b
    ub[e] = db_;
  }
  if (has_x && g.j == 0) {
    float al, be;
    affine_coef(ua_x, ub_x, al, be);
    const float ra = __fdividef(1.f, al);
    dua_x = (1.f - glo[0] * lo[0]) * (al - kMinScale) * 0.5f * ra;
    dub_x = -0.5f * glo[0] * ra;
    lo[0] = fmaf(al, lo[0], be);
    glo[0] *= ra;
  }
  cond_backward<E>(F, g, l, shift, nt_main, has_x, scr, ua, ub, dua_x, dub_x, glo);
  swap_halves(lo, hi, src_is_hi);
  swap_halves(glo, ghi, src_is_hi);
}

// ---- whole-flow passes on physical coordinates -------------------------------------------------------------
// forward: x -> z (physical), returns log|det dz/dx| (group-summed)
template <int E>
__device__ __forceinline__ float flow_forward(const FlowDesc& F, const Geom& g, float (&lo)[E], float (&hi)[E], float* scr) {
  float ld = 0.f;
  affine_apply<E, false>(F, g, 0, lo, hi);
  for (int l = 0; l < F.Lc; ++l) {
    coupling_apply<E, false>(F, g, l, lo, hi, scr, ld);
    affine_apply<E, false>(F, g, 1 + l, lo, hi);
  }
  affine_apply<E, false>(F, g, F.Lc + 1, lo, hi);
  affine_apply<E, false>(F, g, F.Lc + 2, lo, hi);
  return group_sum(ld, g.gs) + F.blob[F.off_const];
}
// inverse: z (physical) -> x, returns log|det dx/dz|
template <int E>
__device__ __forceinline__ float flow_inverse(const FlowDesc& F, const Geom& g, float (&lo)[E], float (&hi)[E], float* scr) {
  float ld = 0.f;
  affine_apply<E, true>(F, g, F.Lc + 2, lo, hi);
  affine_apply<E, true>(F, g, F.Lc + 1, lo, hi);
  for (int l = F.Lc - 1; l >= 0; --l) {
    affine_apply<E, true>(F, g, 1 + l, lo, hi);
    coupling_apply<E, true>(F, g, l, lo, hi, scr, ld);
  }
  affine_apply<E, true>(F, g, 0, lo, hi);
  return -(group_sum(ld, g.gs) + F.blob[F.off_const]);
}
// Given x = T^-1(z) in (lo, hi) and dU/dx in (glo, ghi): walk x -> z, leaving z in (lo, hi) and
// d/dz [ U(T^-1 z) - log|det dT^-1/dz| ] in (glo, ghi).
template <int E>
__device__ __forceinline__ void flow_unwind(const FlowDesc& F, const Geom& g, float (&lo)[E], float (&hi)[E],
                                            float (&glo)[E], float (&ghi)[E], float* scr) {
  affine_unwind<E>(F, g, 0, lo, hi, glo, ghi);
  for (int l = 0; l < F.Lc; ++l) {
    coupling_unwind<E>(F, g, l, lo, hi, glo, ghi, scr);
    affine_unwind<E>(F, g, 1 + l, lo, hi, glo, ghi);
  }
  affine_unwind<E>(F, g, F.Lc + 1, lo, hi, glo, ghi);
  affine_unwind<E>(F, g, F.Lc + 2, lo, hi, glo, ghi);
}

// log N(z; 0, I)  (oracle FlowRef.base_log_prob)
template <int E>
__device__ __forceinline__ float base_log_prob(const Geom& g, const float (&lo)[E], const float (&hi)[E]) {
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) s = fmaf(lo[e], lo[e], fmaf(hi[e], hi[e], s));
  return -0.5f * group_sum(s, g.gs) - 0.5f * (float)g.d * 1.8378770664093453f;
}

}  // namespace nfmc
