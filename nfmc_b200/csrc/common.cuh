// common.cuh -- chain layout, group collectives, Philox4x32-10, Box-Muller.   (sm_100a only)
//
// Chain layout ("halves" layout).  A chain's d values are split at da = d/2 into a low half [0, da) and a
// high half [da, d) (db = d - da elements) -- the RealNVP source/target split, so coupling layers never
// move data between lanes.  A chain is owned by a group of `gs` adjacent lanes (gs = 1,2,4,..,32); lane j
// of the group holds, for slot e = 0..E-1,
//     lo[e] = x[j + gs*e]          (valid iff j + gs*e < da)
//     hi[e] = x[da + j + gs*e]     (valid iff j + gs*e < db)
// in registers for the whole launch.  E is a template parameter (slots per half), gs is a runtime value.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace nfmc {

constexpr int kThreads = 128;          // threads per CTA for every chain kernel
constexpr float kMinScale = 1e-3f;     // affine-map floor m  (oracle/realnvp_ref.py: MIN_SCALE)
constexpr float kLogOneMinusM = -1.0005003335835344e-3f;  // log(1 - 1e-3)

struct Geom {
  int d, da, db;  // event size and the half split
  int gs;         // lanes per chain
  int j;          // this lane's index within its group
  int lane;       // lane in warp
  int grp_base;   // first lane of this group in the warp
};

__device__ __forceinline__ Geom make_geom(int d, int gs) {
  Geom g;
  g.d = d; g.da = d / 2; g.db = d - g.da; g.gs = gs;
  g.lane = threadIdx.x & 31;
  g.j = g.lane & (gs - 1);
  g.grp_base = g.lane - g.j;
  return g;
}

// sum over the lanes of a group (butterfly; every lane gets the total). All 32 lanes must call.
__device__ __forceinline__ float group_sum(float v, int gs) {
  for (int o = gs >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float group_bcast(float v, const Geom& g, int src_j) {
  return __shfl_sync(0xffffffffu, v, g.grp_base + src_j);
}
// sum over the groups of a warp for the same j (lanes j, j+gs, ...); result valid on every lane
__device__ __forceinline__ float across_groups_sum(float v, int gs) {
  for (int o = gs; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2): one issue slot for two IEEE fp32 operations.  The
//      results are bit-identical to the scalar fmaf / * / +; the point is issue bandwidth: these kernels are bound by
//      instruction issue (Philox integer work competes with the float work for the same slots).
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }

// ---- Philox4x32-10 (Salmon et al. 2011; same round function / constants as cuRAND and torch) -------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// round keys precomputed once per kernel (the seed is launch-constant), so the key schedule costs no issue slots
struct PhiloxKeys {
  uint32_t k0[10], k1[10];
};
__device__ __forceinline__ PhiloxKeys philox_keys(uint64_t seed) {
  PhiloxKeys K;
  uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    K.k0[r] = a; K.k1[r] = b;
    a += 0x9E3779B9u; b += 0xBB67AE85u;
  }
  return K;
}
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxKeys& K) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ K.k0[r], lo1, hi0 ^ c.w ^ K.k1[r], lo0);
  }
  return c;
}

// Counter convention (DESIGN.md "random numbers"):
//   c.x = quad * 32 + j      quad = index of the 4-word block within this (chain, step, stream), j = lane in group
//   c.y = stream id | (step >> 32) << 8
//   c.z = step (low 32 bits, global step index)
//   c.w = global chain index
//   key = seed
// Word usage per (chain, step, stream): pair p = 0 -> (accept uniform on lane j = 0, spare);
//   pair p = e + 1 (quad p/2, words 2(p%2), 2(p%2)+1) -> Box-Muller -> (lo[e], hi[e]).
struct RngKey {
  uint2 key;
  uint32_t cy, cz, cw;
};
__device__ __forceinline__ RngKey make_rng_key(uint64_t seed, uint32_t stream_id, uint64_t step, uint64_t chain) {
  RngKey k;
  k.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  k.cy = stream_id | ((uint32_t)(step >> 32) << 8);
  k.cz = (uint32_t)step;
  k.cw = (uint32_t)chain;
  return k;
}
__device__ __forceinline__ uint4 rng_quad(const RngKey& k, int quad, int j) {
  return philox4x32_10(make_uint4((uint32_t)(quad * 32 + j), k.cy, k.cz, k.cw), k.key);
}
__device__ __forceinline__ uint4 rng_quad(const PhiloxKeys& K, const RngKey& k, int quad, int j) {
  return philox4x32_10(make_uint4((uint32_t)(quad * 32 + j), k.cy, k.cz, k.cw), K);
}

__device__ __forceinline__ float uniform_from_bits(uint32_t b) {  // [0,1), 24 bits (torch.rand convention)
  return (float)(b >> 8) * 5.9604644775390625e-8f;
}

__device__ __forceinline__ float fast_lg2(float v) {  // v is a normal float: no denormal fix-up needed
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float fast_sqrt(float v) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// two N(0,1) from two 32-bit words.  u1 in (0,1] from all 32 bits (tail to 6.7 sigma), angle from 23 bits.
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  const float u1 = fmaf(__uint2float_rn(a), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  const float th = (__uint_as_float(0x3f800000u | (b >> 9)) - 1.5f) * 6.283185307179586f;  // [-pi, pi)
  const float r = fast_sqrt(-1.3862943611198906f * fast_lg2(u1));                             // sqrt(-2 ln u1)
  float s, c;
  __sincosf(th, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

// packed variant: returns (z0, z1) with the final scaling as one FMUL2
__device__ __forceinline__ float2 box_muller2(uint32_t a, uint32_t b) {
  const float u1 = fmaf(__uint2float_rn(a), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  const float th = (__uint_as_float(0x3f800000u | (b >> 9)) - 1.5f) * 6.283185307179586f;
  const float r = fast_sqrt(-1.3862943611198906f * fast_lg2(u1));
  float s, c;
  __sincosf(th, &s, &c);
  return mul2(splat2(r), make_float2(c, s));
}

// Fill the per-step normals for this lane (E slots per half) and return the accept uniform bits
// (meaningful on every lane: all lanes of a group compute lane 0's quad 0 only if they are lane 0;
// the caller broadcasts from j = 0).
template <int E>
struct StepNoise {
  float lo[E];
  float hi[E];
  uint32_t ubits;  // word 0 of quad 0 (accept uniform when taken from lane j = 0)
};

template <int E>
__device__ __forceinline__ void draw_step_noise(const RngKey& k, int j, StepNoise<E>& nz) {
  constexpr int NQ = (E + 2) / 2;  // pairs 0..E -> quads 0..E/2
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const uint4 w = rng_quad(k, q, j);
    if (q == 0) {
      nz.ubits = w.x;
      if (E >= 1) box_muller(w.z, w.w, nz.lo[0], nz.hi[0]);
    } else {
      const int e0 = 2 * q - 1;
      if (e0 < E) box_muller(w.x, w.y, nz.lo[e0], nz.hi[e0]);
      if (e0 + 1 < E) box_muller(w.z, w.w, nz.lo[e0 + 1], nz.hi[e0 + 1]);
    }
  }
}

// slot validity.  In an EXACT layout (ceil(db/gs) == E) slots e < E-1 are valid in both halves for every lane,
// so only the last slot needs a runtime test; the compiler folds the rest away.
template <bool EXACT, int E>
__device__ __forceinline__ bool slot_ok(int e, int k, int n_half) {
  if (EXACT && e < E - 1) return true;
  return k < n_half;
}

// ---- loads / stores between the row-major [n, d] tensor and the register layout -----------------------
template <int E>
__device__ __forceinline__ void load_chain(const float* __restrict__ row, const Geom& g, float (&lo)[E], float (&hi)[E]) {
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int k = g.j + g.gs * e;
    lo[e] = (k < g.da) ? __ldg(row + k) : 0.f;
    hi[e] = (k < g.db) ? __ldg(row + g.da + k) : 0.f;
  }
}
template <int E>
__device__ __forceinline__ void store_chain(float* __restrict__ row, const Geom& g, const float (&lo)[E], const float (&hi)[E]) {
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int k = g.j + g.gs * e;
    if (k < g.da) row[k] = lo[e];
    if (k < g.db) row[g.da + k] = hi[e];
  }
}
// reversed order (z <-> physical coordinates when the flow has an odd number of reversals)
template <int E>
__device__ __forceinline__ void load_chain_flipped(const float* __restrict__ row, const Geom& g, float (&lo)[E], float (&hi)[E]) {
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int k = g.j + g.gs * e;
    lo[e] = (k < g.da) ? __ldg(row + (g.d - 1 - k)) : 0.f;
    hi[e] = (k < g.db) ? __ldg(row + (g.d - 1 - (g.da + k))) : 0.f;
  }
}
template <int E>
__device__ __forceinline__ void store_chain_flipped(float* __restrict__ row, const Geom& g, const float (&lo)[E], const float (&hi)[E]) {
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int k = g.j + g.gs * e;
    if (k < g.da) row[g.d - 1 - k] = lo[e];
    if (k < g.db) row[g.d - 1 - (g.da + k)] = hi[e];
  }
}

// ---- per-CTA statistics: fp32 per-thread partials -> warp shuffle over groups -> smem double -> global ----
struct StatsSmem {
  double* sx;   // [d]
  double* sx2;  // [d]
};

template <int E>
__device__ __forceinline__ void flush_moments(const Geom& g, float (&m1lo)[E], float (&m1hi)[E], float (&m2lo)[E],
                                              float (&m2hi)[E], double* sx, double* sx2, bool flip = false) {
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int k = g.j + g.gs * e;
    const float a = across_groups_sum(m1lo[e], g.gs), b = across_groups_sum(m1hi[e], g.gs);
    const float c = across_groups_sum(m2lo[e], g.gs), dd = across_groups_sum(m2hi[e], g.gs);
    if (g.lane < g.gs) {
      const int il = flip ? g.d - 1 - k : k, ih = flip ? g.d - 1 - (g.da + k) : g.da + k;
      if (k < g.da) { atomicAdd(sx + il, (double)a); atomicAdd(sx2 + il, (double)c); }
      if (k < g.db) { atomicAdd(sx + ih, (double)b); atomicAdd(sx2 + ih, (double)dd); }
    }
    m1lo[e] = m1hi[e] = m2lo[e] = m2hi[e] = 0.f;
  }
}

}  // namespace nfmc
