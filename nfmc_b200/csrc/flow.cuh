// flow.cuh -- RealNVP forward / inverse / log-det / input-VJP on the register-resident chain layout.
//
// Specification = oracle/realnvp_ref.py (the torchflows surface the reference calls: sampling/base.py:26,
// util.py:280-281, jump.py:205,218, imh.py:214,221, neutra.py:60).
//
// Two conditioner code paths:
//   * small  (M == 2, H <= 8 -- every default conditioner for d <= 1024): hidden units live in registers, weights are
//     packed hidden-minor and padded to 8 so that one 128-bit load brings four weights; no scratch memory.
//   * generic (any M >= 1, any H): hidden activations in per-group shared scratch; out-of-line, correctness path.
//   (Wide conditioners are meant for the tcgen05 path, cond_tc.cu.)
//
// Physical coordinates.  The flow's reverse permutations are folded into the packed parameters: the chain state is
// never permuted; coupling l (0-based) has its source/target halves swapped when l is even (an odd number of
// reversals precede it) and every parameter is stored in the coordinate the state actually lives in.  Only the
// latent z of a flow with odd Lc is flipped, on load/store.
//
// Blob layout (fp32), da = d/2, db = d - da:
//   [ (Lc+1) elementwise affines ]  affine 0 = the leading ElementwiseAffine; affine l+1 = the act-norm after coupling l;
//                                   the last one also absorbs the trailing ElementwiseAffine + ActNorm (composed at
//                                   pack time).  Each: fwd float2[d] {alpha, beta}, inv float2[d] {1/alpha, -beta/alpha}.
//   [ 4 floats ] [0] = sum over all elementwise affines of sum_i log alpha_i
//   [ Lc couplings ]
//        small  :  W1T[da][8] | b1[8] | WlT[db][2][8] | bl[db][2] (+pad to 4)  (zero padded in h)
//        generic:  M >= 2:  W1[H][da] b1[H] | (M-2) x { Wm[H_in][H_out] bm[H] } | Wl[H][2][db] bl[2][db]
//                  M == 1:  Wl[da][2][db] bl[2][db]
//   (source / target indices are PACKED physical indices; see pack_realnvp in nfmc_b200/flow.py)
#pragma once
#include "common.cuh"

namespace nfmc {

constexpr int kSmallH = 8;

struct FlowDesc {
  const float* blob;
  unsigned sbase;       // shared-window address of blob[0] when the blob is staged in shared memory
  int d, da, db, Lc, M, H;
  int small;            // M == 2 && H <= 8
  int Hs;               // generic path: width of the last layer's input
  int off_const;        // (Lc+1)*4*d
  int off_coupling;     // off_const + 4
  int coupling_stride;  // floats per coupling
  int scratch;          // generic path: floats of scratch per chain group (0 for small)
};

__host__ __device__ inline bool flow_is_small(int M, int H) { return M == 2 && H <= kSmallH; }
__host__ __device__ inline int flow_coupling_floats(int d, int M, int H) {
  const int da = d / 2, db = d - da;
  if (flow_is_small(M, H)) return da * kSmallH + kSmallH + db * 2 * kSmallH + ((db * 2 + 3) & ~3);
  if (M == 1) return da * 2 * db + 2 * db;
  return (da * H + H) + (M - 2) * (H * H + H) + (H * 2 * db + 2 * db);
}
__host__ __device__ inline long long flow_blob_floats(int d, int Lc, int M, int H) {
  return (long long)(Lc + 1) * 4 * d + 4 + (long long)Lc * flow_coupling_floats(d, M, H);
}
__host__ __device__ inline FlowDesc make_flow_desc(const float* blob, int d, int Lc, int M, int H) {
  FlowDesc F;
  F.blob = blob; F.sbase = 0u; F.d = d; F.da = d / 2; F.db = d - F.da; F.Lc = Lc; F.M = M; F.H = H;
  F.small = flow_is_small(M, H) ? 1 : 0;
  F.Hs = (M == 1) ? F.da : H;
  F.off_const = (Lc + 1) * 4 * d;
  F.off_coupling = F.off_const + 4;
  F.coupling_stride = flow_coupling_floats(d, M, H);
  F.scratch = F.small ? 0 : ((M - 1 > 1 ? M - 1 : 1) + 2) * F.Hs;
  return F;
}

// ---- parameter loads by float offset into the blob.  SB = the blob was staged into shared memory: explicit
//      ld.shared with 32-bit addresses (the compiler cannot prove the address space through FlowDesc). ------------------
template <bool SB>
__device__ __forceinline__ float ldp(const FlowDesc& F, int off) {
  if (SB) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(F.sbase + 4u * (unsigned)off));
    return v;
  }
  return __ldg(F.blob + off);
}
template <bool SB>
__device__ __forceinline__ float2 ldp2(const FlowDesc& F, int off) {
  if (SB) {
    float2 v;
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(F.sbase + 4u * (unsigned)off));
    return v;
  }
  return __ldg(reinterpret_cast<const float2*>(F.blob + off));
}
template <bool SB>
__device__ __forceinline__ float4 ldp4(const FlowDesc& F, int off) {
  if (SB) {
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(F.sbase + 4u * (unsigned)off));
    return v;
  }
  return __ldg(reinterpret_cast<const float4*>(F.blob + off));
}

// tanh with absolute error ~1e-7 (the conditioner output feeds a linear layer, so absolute error is what matters)
__device__ __forceinline__ float fast_tanh(float v) {
  const float t = __expf(2.f * v);
  return 1.f - __fdividef(2.f, t + 1.f);
}

// ---- elementwise affine (branch-free: clamped parameter index, invalid slots stay 0).  Both directions are the same
//      fma: forward parameters {alpha, beta}, inverse parameters {1/alpha, -beta/alpha}; `inv` only selects the table.
template <int E, bool SB, bool X>
__device__ __forceinline__ void affine_apply(const FlowDesc& F, const Geom& g, int a, bool inv, float (&lo)[E], float (&hi)[E]) {
  const int base = a * 4 * F.d + (inv ? 2 * F.d : 0);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int k = g.j + g.gs * e;
    const bool vl = slot_ok<X, E>(e, k, g.da), vh = slot_ok<X, E>(e, k, g.db);
    const float2 pl = ldp2<SB>(F, base + 2 * (vl ? k : 0));
    const float2 ph = ldp2<SB>(F, base + 2 * (g.da + (vh ? k : 0)));
    lo[e] = vl ? fmaf(pl.x, lo[e], pl.y) : 0.f;
    hi[e] = vh ? fmaf(ph.x, hi[e], ph.y) : 0.f;
  }
}
// undo an inverse affine (state <- alpha*state + beta) while pulling the gradient back (grad <- grad / alpha)
template <int E, bool SB, bool X>
__device__ __forceinline__ void affine_unwind(const FlowDesc& F, const Geom& g, int a, float (&lo)[E], float (&hi)[E],
                                              float (&glo)[E], float (&ghi)[E]) {
  const int fw = a * 4 * F.d, iv = fw + 2 * F.d;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int k = g.j + g.gs * e;
    const bool vl = slot_ok<X, E>(e, k, g.da), vh = slot_ok<X, E>(e, k, g.db);
    const int il = 2 * (vl ? k : 0), ih = 2 * (g.da + (vh ? k : 0));
    const float2 pl = ldp2<SB>(F, fw + il), ph = ldp2<SB>(F, fw + ih);
    const float rl = ldp<SB>(F, iv + il), rh = ldp<SB>(F, iv + ih);
    lo[e] = vl ? fmaf(pl.x, lo[e], pl.y) : 0.f;
    hi[e] = vh ? fmaf(ph.x, hi[e], ph.y) : 0.f;
    glo[e] = vl ? glo[e] * rl : 0.f;
    ghi[e] = vh ? ghi[e] * rh : 0.f;
  }
}

// ---- small conditioner (M == 2, H <= 8): registers only ---------------------------------------------------------------
// src[e] holds source element ks = j + gs*e - shift (valid iff 0 <= ks < da).  Targets: slot e of the other half has
// packed index t = j + gs*e (valid iff t < nt_main), plus, when has_x, one extra target t = da (the middle element of
// an odd-d chain, which physically is hi[0] of lane 0).  hid[] returns the hidden activations (for the backward).
template <int E, bool SB, bool X>
__device__ __forceinline__ void cond_forward_small(const FlowDesc& F, const Geom& g, int W, const float (&src)[E],
                                                   int shift, int nt_main, bool has_x, float (&hid)[kSmallH], float (&ua)[E],
                                                   float (&ub)[E], float& ua_x, float& ub_x) {
  const int da = F.da, db = F.db;
  float acc[kSmallH];
#pragma unroll
  for (int h = 0; h < kSmallH; ++h) acc[h] = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int ks = g.j + g.gs * e - shift;
    const bool ok = (e > 0 || ks >= 0) && slot_ok<X, E>(e, ks, da);   // in an exact layout only slot 0 / the last slot can be invalid
    const float v = ok ? src[e] : 0.f;
    const int w = W + (ok ? ks : 0) * kSmallH;
    const float4 w0 = ldp4<SB>(F, w), w1 = ldp4<SB>(F, w + 4);
    acc[0] = fmaf(w0.x, v, acc[0]); acc[1] = fmaf(w0.y, v, acc[1]); acc[2] = fmaf(w0.z, v, acc[2]); acc[3] = fmaf(w0.w, v, acc[3]);
    acc[4] = fmaf(w1.x, v, acc[4]); acc[5] = fmaf(w1.y, v, acc[5]); acc[6] = fmaf(w1.z, v, acc[6]); acc[7] = fmaf(w1.w, v, acc[7]);
  }
  const int b1 = W + da * kSmallH;
#pragma unroll
  for (int h = 0; h < kSmallH; ++h) hid[h] = fast_tanh(group_sum(acc[h], g.gs) + ldp<SB>(F, b1 + h));   // padded h: tanh(0) = 0
  const int Wl = b1 + kSmallH;
  const int bl = Wl + db * 2 * kSmallH;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int t = g.j + g.gs * e;
    const int tc = slot_ok<X, E>(e, t, nt_main) ? t : 0;
    const int w = Wl + tc * 2 * kSmallH;
    const float4 a0 = ldp4<SB>(F, w), a1 = ldp4<SB>(F, w + 4), c0 = ldp4<SB>(F, w + 8), c1 = ldp4<SB>(F, w + 12);
    const float2 b = ldp2<SB>(F, bl + 2 * tc);
    float sa = b.x, sb = b.y;
    sa = fmaf(a0.x, hid[0], sa); sa = fmaf(a0.y, hid[1], sa); sa = fmaf(a0.z, hid[2], sa); sa = fmaf(a0.w, hid[3], sa);
    sa = fmaf(a1.x, hid[4], sa); sa = fmaf(a1.y, hid[5], sa); sa = fmaf(a1.z, hid[6], sa); sa = fmaf(a1.w, hid[7], sa);
    sb = fmaf(c0.x, hid[0], sb); sb = fmaf(c0.y, hid[1], sb); sb = fmaf(c0.z, hid[2], sb); sb = fmaf(c0.w, hid[3], sb);
    sb = fmaf(c1.x, hid[4], sb); sb = fmaf(c1.y, hid[5], sb); sb = fmaf(c1.z, hid[6], sb); sb = fmaf(c1.w, hid[7], sb);
    ua[e] = sa;
    ub[e] = sb;
  }
  ua_x = ub_x = 0.f;
  if (has_x) {
    const int w = Wl + da * 2 * kSmallH;
    const float2 b = ldp2<SB>(F, bl + 2 * da);
    float sa = b.x, sb = b.y;
#pragma unroll
    for (int h = 0; h < kSmallH; ++h) {
      sa = fmaf(ldp<SB>(F, w + h), hid[h], sa);
      sb = fmaf(ldp<SB>(F, w + kSmallH + h), hid[h], sb);
    }
    ua_x = sa;
    ub_x = sb;
  }
}

// input-VJP of the small conditioner: dsrc[e] += sum_out d_out * d out / d src[e]
template <int E, bool SB, bool X>
__device__ __forceinline__ void cond_backward_small(const FlowDesc& F, const Geom& g, int W, int shift, int nt_main,
                                                    bool has_x, const float (&hid)[kSmallH], const float (&dua)[E],
                                                    const float (&dub)[E], float dua_x, float dub_x, float (&dsrc)[E]) {
  const int da = F.da, db = F.db;
  const int b1 = W + da * kSmallH;
  const int Wl = b1 + kSmallH;
  float acc[kSmallH];
#pragma unroll
  for (int h = 0; h < kSmallH; ++h) acc[h] = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int t = g.j + g.gs * e;
    const bool ok = slot_ok<X, E>(e, t, nt_main);
    const float va = ok ? dua[e] : 0.f, vb = ok ? dub[e] : 0.f;
    const int w = Wl + (ok ? t : 0) * 2 * kSmallH;
    const float4 a0 = ldp4<SB>(F, w), a1 = ldp4<SB>(F, w + 4), c0 = ldp4<SB>(F, w + 8), c1 = ldp4<SB>(F, w + 12);
    acc[0] = fmaf(a0.x, va, fmaf(c0.x, vb, acc[0])); acc[1] = fmaf(a0.y, va, fmaf(c0.y, vb, acc[1]));
    acc[2] = fmaf(a0.z, va, fmaf(c0.z, vb, acc[2])); acc[3] = fmaf(a0.w, va, fmaf(c0.w, vb, acc[3]));
    acc[4] = fmaf(a1.x, va, fmaf(c1.x, vb, acc[4])); acc[5] = fmaf(a1.y, va, fmaf(c1.y, vb, acc[5]));
    acc[6] = fmaf(a1.z, va, fmaf(c1.z, vb, acc[6])); acc[7] = fmaf(a1.w, va, fmaf(c1.w, vb, acc[7]));
  }
  if (has_x && g.j == 0) {
    const int w = Wl + da * 2 * kSmallH;
#pragma unroll
    for (int h = 0; h < kSmallH; ++h) acc[h] = fmaf(ldp<SB>(F, w + h), dua_x, fmaf(ldp<SB>(F, w + kSmallH + h), dub_x, acc[h]));
  }
  float dpre[kSmallH];
#pragma unroll
  for (int h = 0; h < kSmallH; ++h) dpre[h] = group_sum(acc[h], g.gs) * (1.f - hid[h] * hid[h]);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int ks = g.j + g.gs * e - shift;
    const bool ok = (e > 0 || ks >= 0) && slot_ok<X, E>(e, ks, da);
    const int w = W + (ok ? ks : 0) * kSmallH;
    const float4 w0 = ldp4<SB>(F, w), w1 = ldp4<SB>(F, w + 4);
    float s = w0.x * dpre[0];
    s = fmaf(w0.y, dpre[1], s); s = fmaf(w0.z, dpre[2], s); s = fmaf(w0.w, dpre[3], s);
    s = fmaf(w1.x, dpre[4], s); s = fmaf(w1.y, dpre[5], s); s = fmaf(w1.z, dpre[6], s); s = fmaf(w1.w, dpre[7], s);
    if (ok) dsrc[e] += s;
  }
}

// ---- generic conditioner MLP (any M, H): out of line, activations in shared scratch ------------------------------------
template <int E>
__device__ __noinline__ void cond_forward_generic(const FlowDesc& F, const Geom& g, const float* W, const float (&src)[E],
                                                  int shift, int nt_main, bool has_x, float* scr, float (&ua)[E], float (&ub)[E],
                                                  float& ua_x, float& ub_x) {
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
        if (h < H && g.j == (h & (g.gs - 1))) scr[h] = tanhf(s + b1[h]);
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
        out[hq] = tanhf(acc);
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
  const float* bl = Wl + Hs * 2 * db;
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
    const float* wa = Wl + h * 2 * db;
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

template <int E>
__device__ __noinline__ void cond_backward_generic(const FlowDesc& F, const Geom& g, const float* W, int shift, int nt_main,
                                                   bool has_x, float* scr, const float (&dua)[E], const float (&dub)[E],
                                                   float dua_x, float dub_x, float (&dsrc)[E]) {
  const int da = F.da, db = F.db, H = F.H, Hs = F.Hs;
  float* gA = scr + (F.M - 1 > 1 ? F.M - 1 : 1) * Hs;
  float* gB = gA + Hs;
  __syncwarp();
  if (F.M >= 2) {
    const float* W1 = W;
    const float* Wm_first = W + H * da + H;
    const float* Wl = Wm_first + (F.M - 2) * (H * H + H);
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
              const float* wa = Wl + (h0 + u) * 2 * db;
              acc[u] = fmaf(wa[t], dua[e], fmaf(wa[db + t], dub[e], acc[u]));
            }
        }
      }
      if (has_x && g.j == 0) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (h0 + u < H) {
            const float* wa = Wl + (h0 + u) * 2 * db;
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
      const float* Wm = Wm_first + (m - 1) * (H * H + H);
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
      const float* wa = Wl + ks0 * 2 * db;
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

// ---- affine coupling layer l; inv = false: x -> z direction, true: z -> x (one code instance for both: the target
//      update is fma(r, b, s) with (r, s) = (alpha, beta) or (1/alpha, -beta/alpha)).  ld accumulates THIS LANE's share
//      of sum_t log alpha_t (the caller group-sums once per pass and applies the sign).  When the source is the high
//      half the two register arrays are swapped around the layer so that there is a single conditioner call site.
// Conditioner stash (NeuTra): the inverse pass keeps each coupling's conditioner outputs {u_a[E], u_b[E], hid[8], u_a_x,
// u_b_x} in per-thread shared memory (element i at stash[i * kThreads]) and the backward sweep reads them back instead of
// evaluating the conditioner a second time on the same input.
template <int E>
__host__ __device__ constexpr int cond_stash_floats() { return 2 * E + kSmallH + 2; }

template <int E, bool SB, bool X, bool SM>
__device__ __forceinline__ void coupling_apply(const FlowDesc& F, const Geom& g, int l, bool inv, float (&lo)[E], float (&hi)[E],
                                               float* scr, float& ld, float* stash = nullptr) {
  float ua[E], ub[E], ua_x, ub_x;
  const bool src_is_hi = (l & 1) == 0;
  const int shift = src_is_hi ? F.db - F.da : 0;
  const bool has_x = shift > 0;
  const int nt_main = src_is_hi ? F.da : F.db;
  const int Woff = F.off_coupling + l * F.coupling_stride;
  swap_halves(lo, hi, src_is_hi);  // lo = source array, hi = target array
  if constexpr (SM) {
    float hid[kSmallH];
    cond_forward_small<E, SB, X>(F, g, Woff, lo, shift, nt_main, has_x, hid, ua, ub, ua_x, ub_x);
    if (stash) {
#pragma unroll
      for (int e = 0; e < E; ++e) { stash[e * kThreads] = ua[e]; stash[(E + e) * kThreads] = ub[e]; }
#pragma unroll
      for (int h = 0; h < kSmallH; ++h) stash[(2 * E + h) * kThreads] = hid[h];
      stash[(2 * E + kSmallH) * kThreads] = ua_x;
      stash[(2 * E + kSmallH + 1) * kThreads] = ub_x;
    }
  } else {
    cond_forward_generic<E>(F, g, F.blob + Woff, lo, shift, nt_main, has_x, scr, ua, ub, ua_x, ub_x);
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const bool ok = slot_ok<X, E>(e, g.j + g.gs * e, nt_main);
    float al, be;
    affine_coef(ua[e], ub[e], al, be);
    const float ra = __fdividef(1.f, al);
    const float r = inv ? ra : al, s = inv ? -be * ra : be;
    hi[e] = ok ? fmaf(r, hi[e], s) : hi[e];
    ld += ok ? __logf(al) : 0.f;
  }
  if (has_x && g.j == 0) {  // middle element of an odd-d chain: physically hi[0], here slot 0 of the source array
    float al, be;
    affine_coef(ua_x, ub_x, al, be);
    const float ra = __fdividef(1.f, al);
    lo[0] = fmaf(inv ? ra : al, lo[0], inv ? -be * ra : be);
    ld += __logf(al);
  }
  swap_halves(lo, hi, src_is_hi);
}

// Undo inverse-coupling l: (lo, hi) currently hold the coupling's OUTPUT of the z->x pass (a, b) and (glo, ghi)
// the gradient of U~ with respect to it.  On return they hold its INPUT (a, b') and the gradient with respect to
// that input, where U~ = U(x) - log|det dx/dz| so each coupling contributes + sum log alpha to U~:
//   b = (b' - beta)/alpha  =>  dU~/dalpha = (1 - gb*b)/alpha,  dU~/dbeta = -gb/alpha,  dU~/db' = gb/alpha.
template <int E, bool SB, bool X, bool SM>
__device__ __forceinline__ void coupling_unwind(const FlowDesc& F, const Geom& g, int l, float (&lo)[E], float (&hi)[E],
                                                float (&glo)[E], float (&ghi)[E], float* scr, const float* stash = nullptr) {
  float ua[E], ub[E], ua_x = 0.f, ub_x = 0.f;
  float dua_x = 0.f, dub_x = 0.f;
  float hid[kSmallH];
  const bool src_is_hi = (l & 1) == 0;
  const int shift = src_is_hi ? F.db - F.da : 0;
  const bool has_x = shift > 0;
  const int nt_main = src_is_hi ? F.da : F.db;
  const int Woff = F.off_coupling + l * F.coupling_stride;
  swap_halves(lo, hi, src_is_hi);
  swap_halves(glo, ghi, src_is_hi);
  if constexpr (SM) {
    if (stash) {
#pragma unroll
      for (int e = 0; e < E; ++e) { ua[e] = stash[e * kThreads]; ub[e] = stash[(E + e) * kThreads]; }
#pragma unroll
      for (int h = 0; h < kSmallH; ++h) hid[h] = stash[(2 * E + h) * kThreads];
      ua_x = stash[(2 * E + kSmallH) * kThreads];
      ub_x = stash[(2 * E + kSmallH + 1) * kThreads];
    } else {
      cond_forward_small<E, SB, X>(F, g, Woff, lo, shift, nt_main, has_x, hid, ua, ub, ua_x, ub_x);
    }
  } else {
    cond_forward_generic<E>(F, g, F.blob + Woff, lo, shift, nt_main, has_x, scr, ua, ub, ua_x, ub_x);
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    float da_ = 0.f, db_ = 0.f;
    if (slot_ok<X, E>(e, g.j + g.gs * e, nt_main)) {
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
  if constexpr (SM) cond_backward_small<E, SB, X>(F, g, Woff, shift, nt_main, has_x, hid, ua, ub, dua_x, dub_x, glo);
  else cond_backward_generic<E>(F, g, F.blob + Woff, shift, nt_main, has_x, scr, ua, ub, dua_x, dub_x, glo);
  swap_halves(lo, hi, src_is_hi);
  swap_halves(glo, ghi, src_is_hi);
}

// ---- whole-flow pass on physical coordinates: one loop over the 2*Lc+1 layers, walked forwards (x -> z, returns
//      log|det dz/dx|) or backwards (z -> x, returns log|det dx/dz|).  A single code instance of each layer type, so the
//      instruction footprint stays small (the jump executes this code once per launch per warp: fetch matters).
template <int E, bool SB, bool X, bool SM>
__device__ __forceinline__ float flow_pass(const FlowDesc& F, const Geom& g, bool inv, float (&lo)[E], float (&hi)[E], float* scr,
                                           float* stash = nullptr) {
  float ld = 0.f;
  const int n_ops = 2 * F.Lc + 1;
#pragma unroll 1
  for (int i = 0; i < n_ops; ++i) {
    const int op = inv ? n_ops - 1 - i : i;
    if (op & 1)
      coupling_apply<E, SB, X, SM>(F, g, op >> 1, inv, lo, hi, scr, ld,
                                   stash ? stash + (size_t)(op >> 1) * cond_stash_floats<E>() * kThreads : nullptr);
    else affine_apply<E, SB, X>(F, g, op >> 1, inv, lo, hi);
  }
  const float tot = group_sum(ld, g.gs) + ldp<SB>(F, F.off_const);
  return inv ? -tot : tot;
}
template <int E, bool SB, bool X, bool SM>
__device__ __forceinline__ float flow_forward(const FlowDesc& F, const Geom& g, float (&lo)[E], float (&hi)[E], float* scr) {
  return flow_pass<E, SB, X, SM>(F, g, false, lo, hi, scr);
}
template <int E, bool SB, bool X, bool SM>
__device__ __forceinline__ float flow_inverse(const FlowDesc& F, const Geom& g, float (&lo)[E], float (&hi)[E], float* scr,
                                              float* stash = nullptr) {
  return flow_pass<E, SB, X, SM>(F, g, true, lo, hi, scr, stash);
}
// Given x = T^-1(z) in (lo, hi) and dU/dx in (glo, ghi): walk x -> z, leaving z in (lo, hi) and
// d/dz [ U(T^-1 z) - log|det dT^-1/dz| ] in (glo, ghi).
template <int E, bool SB, bool X, bool SM>
__device__ __forceinline__ void flow_unwind(const FlowDesc& F, const Geom& g, float (&lo)[E], float (&hi)[E],
                                            float (&glo)[E], float (&ghi)[E], float* scr, const float* stash = nullptr) {
  affine_unwind<E, SB, X>(F, g, 0, lo, hi, glo, ghi);
  for (int l = 0; l < F.Lc; ++l) {
    coupling_unwind<E, SB, X, SM>(F, g, l, lo, hi, glo, ghi, scr,
                                  stash ? stash + (size_t)l * cond_stash_floats<E>() * kThreads : nullptr);
    affine_unwind<E, SB, X>(F, g, 1 + l, lo, hi, glo, ghi);
  }
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
