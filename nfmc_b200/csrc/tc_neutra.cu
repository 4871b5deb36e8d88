// tc_neutra.cu -- NeuTra HMC for wide flows with the conditioner MLP AND ITS INPUT-VJP on the tensor cores (tcgen05 + TMEM).
//
// Reference: NeuTra.adjusted_target (nfmc/neutra.py:58-68)  U~(z) = U(T^-1 z) - log|det dT^-1/dz|, differentiated with
// respect to z by autograd inside HMC.propose (mcmc/hmc.py:40-48,61-77,96-126).  Here one gradient evaluation is
//   1. flow_tc_kernel<INV> (tc_flow.cu):            z -> x, log|det dx/dz|
//   2. neutra_unwind_tc_kernel (this file):         U(x), grad U(x) (closed forms), then the reversible backward sweep
//      x -> z that re-derives every layer's input from its output while it pulls the cotangent back (flow.cuh:
//      flow_unwind) -- with, per coupling and 128-chain tile, FOUR GEMMs on tcgen05:
//        G1  Hpre[128 x Hp]   = [S | 1 | 1][128 x K1] . W1^T              (conditioner forward, as in tc_flow.cu)
//        G2  U'  [128 x N2p]  = tanh(Hpre)[128 x Hp] . Wl'^T
//        G3  dH  [128 x Hp]   = dU'[128 x N2p] . Wl'                        (dgrad of the last layer; two N-halves)
//        G4  dS  [128 x K1]   = (dH * (1 - tanh^2))[128 x Hp] . W1          (dgrad of the first layer)
//      bf16 operands, fp32 accumulators in tensor memory; tanh(Hpre), dU' and dH*(1-tanh^2) are written back to tensor
//      memory as packed bf16 and consumed from there as the A operands of G2 / G3 / G4 (never through shared memory).
//      The leapfrog update (hmc.py:51-58: half-kicks p <- p - tau/2 grad, drift z <- z + tau p m) is fused into the kernel's
//      output stage: the momentum and latent tiles are staged by TMA into the weight buffers the last coupling has released,
//      updated in place and stored back by TMA -- no gradient tensor is ever written.
//   3. small kernels for the momentum draw and the accept step (hmc.py:100,103-113).
// The reference evaluates the gradient 2L times per step; consecutive half-kicks share an evaluation here (L + 1).
//
// Tensor-memory plan of the unwind kernel (one tile at a time, 512 columns):
//   [0, Hp)               Hpre (fp32)  ->  [0, Hp/2) tanh as packed bf16 pairs (K-step s = hidden [8s, 8s+8) and
//                                           [Hp/2 + 8s, Hp/2 + 8s + 8), the split layout of tc_flow.cu)
//   [Hp/2, Hp)            dH of hidden units [0, Hp/2) (fp32)  ->  in place: dH*(1-tanh^2) as packed bf16 (A of G4)
//   [256, 256 + N2p)      U' (fp32)  ->  in place: dU' as packed bf16 (A of G3; thread g owns columns [32 g, 32 g + 32))
//                         later dS [256, 256 + K1) (fp32, D of G4)
//   [384, 384 + Hp/2)     dH of hidden units [Hp/2, Hp) (fp32)
// Shared memory: the A1 image, one buffer per weight image (W1, Wl', Wl'^T, W1^T; refilled by TMA for the next coupling as
// soon as the GEMM that read them has completed), second-layer biases and affine tables.  The Wl'^T buffer doubles as the
// tile buffer: x arrives there by TMA before the first coupling needs Wl'^T, and the gradient tile leaves from there.
//
// Exact HMC: U~ and its gradient come from the same deterministic bf16 flow, the leapfrog map is volume preserving and
// reversible for ANY force field, and the accept test uses U~ itself -- so the chain targets exp(-U~) exactly and
// x = T^-1(z) has the target's distribution whatever the rounding of the flow.
#include <cuda_bf16.h>
#include "host_common.cuh"
#include "chain_kernel.cuh"
#include "tc_common.cuh"

namespace nfmc {

#ifdef NFMC_NU_TRACE
__device__ long long* g_nu_trace = nullptr;   // [4096] {event id, clock} of epilogue thread 0 of CTA 0 (tools/nu_trace.py)
#define NU_EV(id) do { if (tr_p && tr_n < 2000) { tr_p[2 * tr_n] = (id); tr_p[2 * tr_n + 1] = clock64(); ++tr_n; } } while (0)
#else
#define NU_EV(id)
#endif

constexpr int kNuColU = 256, kNuColDhHi = 384;
enum { kNuBarA1 = 0, kNuBarG1, kNuBarHid, kNuBarG2, kNuBarDu, kNuBarG3, kNuBarDpre, kNuBarG4, kNuBarW1, kNuBarWl, kNuBarWlT,
       kNuBarW1T, kNuBarXFull, kNuBarXRead, kNuBarGOut, kNuBarPZFull, kNuBarPZOut, kNuBarBufFree, kNuNumBars };

struct NuArgs {
  const unsigned char* blob;    // tc blob: affines | per coupling {W1 image, Wl' image, bl'}
  const unsigned char* blobT;   // per coupling {Wl'^T image [N2p/8][Hp][8], W1^T image [Hp/8][K1][8]} bf16
  TcShape S;
  int pot_kind;
  PotParams pot;
  const float* x;        // [n, d]  x = T^-1(z) from the inverse pass
  const float* ld_inv;   // [n]     log|det dx/dz|
  float* grad;           // [n, d]  dU~/dz (logical order), or nullptr when the leapfrog update is fused (p != nullptr)
  float* value;          // [n]     U~(z)
  long long n;
  // fused leapfrog update (hmc.py:51-58): p <- p - kicks * tau/2 * grad;  if drift: z <- z + tau * (p * m)
  float* p;              // [n, d] momentum, updated in place (nullptr: gradient output mode)
  float* zw;             // [n, d] latent state, updated in place when drift
  const float* imd;      // [d] inverse mass diagonal or nullptr (= ones)
  float half_tau, tau;
  int kicks, drift;
};

struct NuSmem {
  unsigned char *a1, *w1, *wl, *wlT, *w1T;
  float* bl;
  float4* aff4;
  float* red;          // [2][6][128] row scratch of the potential (two sets, alternating by tile)
  uint64_t* bars;
  uint32_t* tmem_slot;
};
__host__ __device__ inline size_t nu_wlT_bytes(const TcShape& S) {
  const size_t w = (size_t)S.N2p * S.Hp * 2, t = tc_tile_bytes(S);
  return ((w > t ? w : t) + 127) & ~size_t(127);
}
__host__ __device__ inline size_t nu_w1T_bytes(const TcShape& S) { return (size_t)S.Hp * S.K1 * 2; }
// the Wl' buffer doubles as the z tile buffer of the fused leapfrog update
__host__ __device__ inline size_t nu_wl_bytes(const TcShape& S) {
  const size_t w = tc_wl_bytes(S), t = tc_tile_bytes(S);
  return ((w > t ? w : t) + 127) & ~size_t(127);
}
__host__ __device__ inline size_t nu_smem_total(const TcShape& S) {
  return tc_a1_bytes(S) + tc_w1_bytes(S) + nu_wl_bytes(S) + nu_wlT_bytes(S) + nu_w1T_bytes(S) + (size_t)S.Lc * S.N2p * 4 + tc_aff4_bytes(S) +
         (size_t)2 * 6 * kTcRows * 4 + (size_t)kNuNumBars * 8 + 16;
}
__device__ __forceinline__ NuSmem nu_carve(unsigned char* p, const TcShape& S) {
  NuSmem m;
  m.a1 = p; p += tc_a1_bytes(S);
  m.w1 = p; p += tc_w1_bytes(S);
  m.wl = p; p += nu_wl_bytes(S);
  m.wlT = p; p += nu_wlT_bytes(S);
  m.w1T = p; p += nu_w1T_bytes(S);
  m.bl = reinterpret_cast<float*>(p); p += (size_t)S.Lc * S.N2p * 4;
  m.aff4 = reinterpret_cast<float4*>(p); p += tc_aff4_bytes(S);
  m.red = reinterpret_cast<float*>(p); p += (size_t)2 * 6 * kTcRows * 4;
  m.bars = reinterpret_cast<uint64_t*>(p);
  m.tmem_slot = reinterpret_cast<uint32_t*>(p + (size_t)kNuNumBars * 8);
  return m;
}

__device__ __forceinline__ void tmem_st16u(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
               "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld3(uint32_t (&a)[8], uint32_t (&b)[8], uint32_t (&c)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(b[0]), "+r"(b[1]),
                 "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]), "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]),
                 "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7])
               :
               : "memory");
}

// ---- potentials in the (row, column group) layout: value (as tc_jump.cu) and gradient -----------------------------------
__device__ __forceinline__ float nu_pot_partial(int kind, const PotParams& P, int da, int e0, const float (&lo)[kTcOwn], const float (&hi)[kTcOwn]) {
  float s = 0.f;
  if (kind == NFMC_POT_DIAG_GAUSSIAN) {
    const float2* wm = reinterpret_cast<const float2*>(P.params);
#pragma unroll
    for (int i = 0; i < kTcOwn; ++i) {
      const int k = e0 + i;
      if (k < da) {
        const float2 pl = __ldg(wm + k), ph = __ldg(wm + da + k);
        const float a = lo[i] - pl.y, b = hi[i] - ph.y;
        s = fmaf(pl.x * a, a, fmaf(ph.x * b, b, s));
      }
    }
  } else if (kind == NFMC_POT_ROSENBROCK) {
#pragma unroll
    for (int i = 0; i < kTcOwn; ++i)
      if (e0 + i < da) {
        const float t = lo[i] - 1.f, r = hi[i] - lo[i] * lo[i];
        s += t * t + P.s0 * r * r;
      }
  } else {
#pragma unroll
    for (int i = 0; i < kTcOwn; ++i)
      if (e0 + i < da) s = fmaf(lo[i], lo[i], fmaf(hi[i], hi[i], s));
  }
  return s;
}
// S = sum of the four partials; x0, x1 = the first two coordinates of the chain.  Returns U and fills this thread's part
// of grad U (potentials.cuh: pot_prepare / pot_grad).
__device__ __forceinline__ float nu_pot_value_grad(int kind, const PotParams& P, int d, int da, int e0, float S, float x0, float x1,
                                                   const float (&lo)[kTcOwn], const float (&hi)[kTcOwn], float (&glo)[kTcOwn],
                                                   float (&ghi)[kTcOwn]) {
  float u;
  switch (kind) {
    case NFMC_POT_ISO_GAUSSIAN: {
      u = 0.5f * P.s0 * S;
#pragma unroll
      for (int i = 0; i < kTcOwn; ++i) { glo[i] = P.s0 * lo[i]; ghi[i] = P.s0 * hi[i]; }
    } break;
    case NFMC_POT_DIAG_GAUSSIAN: {
      u = 0.5f * S;
      const float2* wm = reinterpret_cast<const float2*>(P.params);
#pragma unroll
      for (int i = 0; i < kTcOwn; ++i) {
        const int k = e0 + i;
        glo[i] = ghi[i] = 0.f;
        if (k < da) {
          const float2 pl = __ldg(wm + k), ph = __ldg(wm + da + k);
          glo[i] = pl.x * (lo[i] - pl.y);
          ghi[i] = ph.x * (hi[i] - ph.y);
        }
      }
    } break;
    case NFMC_POT_ROSENBROCK: {
      u = S;
#pragma unroll
      for (int i = 0; i < kTcOwn; ++i) {
        const float r = hi[i] - lo[i] * lo[i];
        glo[i] = 2.f * (lo[i] - 1.f) - 4.f * P.s0 * lo[i] * r;
        ghi[i] = 2.f * P.s0 * r;
      }
    } break;
    case NFMC_POT_FUNNEL: {
      const float ex = __expf(-x0), rest = S - x0 * x0;
      u = x0 * x0 * P.s1 + 0.5f * (float)(d - 1) * x0 + 0.5f * ex * rest;
#pragma unroll
      for (int i = 0; i < kTcOwn; ++i) { glo[i] = ex * lo[i]; ghi[i] = ex * hi[i]; }
      if (e0 == 0) glo[0] = 2.f * P.s1 * x0 + 0.5f * (float)(d - 1) - 0.5f * ex * rest;
    } break;
    default: {
      const float a = P.s0, base = -0.5f * (S + 2.f * a * a);
      const float q0 = base + a * (x0 + x1), q1 = base + a * (x0 - x1), q2 = base + a * (-x0 + x1), q3 = base + a * (-x0 - x1);
      const float m = fmaxf(fmaxf(q0, q1), fmaxf(q2, q3));
      const float w0 = __expf(q0 - m), w1 = __expf(q1 - m), w2 = __expf(q2 - m), w3 = __expf(q3 - m);
      const float z = w0 + w1 + w2 + w3, rz = 1.f / z;
      u = -(m + __logf(z));
#pragma unroll
      for (int i = 0; i < kTcOwn; ++i) { glo[i] = lo[i]; ghi[i] = hi[i]; }
      if (e0 == 0) {
        glo[0] = lo[0] - a * (w0 + w1 - w2 - w3) * rz;
        if (da >= 2) glo[1] = lo[1] - a * (w0 - w1 + w2 - w3) * rz;
        else ghi[0] = hi[0] - a * (w0 - w1 + w2 - w3) * rz;
      }
    } break;
  }
#pragma unroll
  for (int i = 0; i < kTcOwn; ++i)
    if (e0 + i >= da) { glo[i] = 0.f; ghi[i] = 0.f; }
  return u;
}

// undo an elementwise affine of the z -> x pass: state <- alpha state + beta, cotangent <- cotangent / alpha
__device__ __forceinline__ void nu_affine_unwind(const float4* aff4, int idx, int e0, float (&lo)[kTcOwn], float (&hi)[kTcOwn],
                                                 float (&glo)[kTcOwn], float (&ghi)[kTcOwn]) {
  const float4* fw = aff4 + (idx * 2) * 64 + e0;
  const float4* iv = aff4 + (idx * 2 + 1) * 64 + e0;
#pragma unroll
  for (int q = 0; q < kTcOwn; ++q) {
    const float4 p = fw[q], r = iv[q];
    lo[q] = fmaf(p.x, lo[q], p.y);
    hi[q] = fmaf(p.z, hi[q], p.w);
    glo[q] *= r.x;
    ghi[q] *= r.z;
  }
}

// epilogue 2 of the unwind: U' -> (alpha, beta); the target half goes back to its z-side value, its cotangent is divided
// by alpha, and dL/dU' replaces U' in tensor memory as packed bf16 (flow.cuh: coupling_unwind with the constants folded
// as in tc_flow.cu: u_a' = log2(e) (log(1-m) + u_a / 2), u_b' = u_b / 2):
//   dU~/du_a' = (1 - g b)(alpha - m)/alpha * ln 2,   dU~/du_b' = -g / alpha
__device__ __forceinline__ void nu_epi2(uint32_t tcol_u, const float* bl, int N2p, int g, int da, float (&tgt)[kTcOwn], float (&gt)[kTcOwn]) {
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = 0u;
  const int e0 = g * kTcOwn;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = 4 * g + 2 * h;
    if (c * 8 < N2p) {
      uint32_t v0[8], v1[8];
      tmem_ld8(tcol_u + c * 8, v0);
      tmem_ld8(tcol_u + c * 8 + 8, v1);
      const float4* b4 = reinterpret_cast<const float4*>(bl + c * 8);
      const float4 bb[4] = {b4[0], b4[1], b4[2], b4[3]};
      tmem_wait_ld(v0, v1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t* v = j < 4 ? v0 : v1;
        const int jj = j & 3;
        const float4 b = bb[j >> 1];
        const float ba = (j & 1) ? b.z : b.x, bbv = (j & 1) ? b.w : b.y;
        const int i = 8 * h + j;
        const float al = fast_ex2(__uint_as_float(v[2 * jj]) + ba) + kMinScale;
        const float be = __uint_as_float(v[2 * jj + 1]) + bbv;
        const float ra = fast_rcp(al);
        const bool ok = e0 + i < da;
        const float xt = tgt[i], gv = gt[i];
        const float dua = ok ? (1.f - gv * xt) * (al - kMinScale) * ra * 0.6931471805599453f : 0.f;
        const float dub = ok ? -gv * ra : 0.f;
        tgt[i] = fmaf(al, xt, be);
        gt[i] = gv * ra;
        w[i] = pack_bf16(dua, dub);
      }
    }
  }
  tmem_st16u(tcol_u + 32 * g, w);
  tmem_wait_st();
}

// epilogue 3: dH * (1 - tanh^2) -> packed bf16, in place over the low half of dH (the A operand of G4, split layout)
__device__ __forceinline__ void nu_epi3(uint32_t trow, int Hp, int g) {
  const int nsteps = Hp >> 4;
  const uint32_t lo_col = (uint32_t)(Hp >> 1);
  for (int s = g; s < nsteps; s += kTcGroups) {
    uint32_t hd[8], a[8], b[8], q[8];
    tmem_ld8(trow + 8 * s, hd);
    tmem_ld8(trow + lo_col + 8 * s, a);
    tmem_ld8(trow + kNuColDhHi + 8 * s, b);
    tmem_wait_ld3(hd, a, b);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float h0 = __uint_as_float(hd[i] << 16), h1 = __uint_as_float(hd[i] & 0xffff0000u);
      const float k0 = __uint_as_float(hd[4 + i] << 16), k1 = __uint_as_float(hd[4 + i] & 0xffff0000u);
      q[i] = pack_bf16(__uint_as_float(a[2 * i]) * (1.f - h0 * h0), __uint_as_float(a[2 * i + 1]) * (1.f - h1 * h1));
      q[4 + i] = pack_bf16(__uint_as_float(b[2 * i]) * (1.f - k0 * k0), __uint_as_float(b[2 * i + 1]) * (1.f - k1 * k1));
    }
    tmem_st8(trow + lo_col + 8 * s, q);
  }
  tmem_wait_st();
}

// ---- the kernel ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1) neutra_unwind_tc_kernel(const __grid_constant__ NuArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  const TcShape& S = A.S;
  const int d = S.d, da = d / 2, Lc = S.Lc, Hp = S.Hp;
  NuSmem sm = nu_carve(smem, S);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bars = smem_u32(sm.bars);
  auto bar = [&](int which) { return bars + 8u * (uint32_t)which; };

  // ---- prologue -----------------------------------------------------------------------------------------------------------
  if (tid == 0) {
    const int full[] = {kNuBarA1, kNuBarHid, kNuBarDu, kNuBarDpre, kNuBarXRead, kNuBarGOut, kNuBarPZOut};
    for (int i = 0; i < kNuNumBars; ++i) {
      bool wide = false;
      for (int j = 0; j < 7; ++j) wide = wide || (full[j] == i);
      mbar_init(bar(i), wide ? kTcEpiThreads : 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(sm.tmem_slot), 512);
  const float* gaff = reinterpret_cast<const float*>(A.blob);
  for (int i = tid; i < (Lc + 1) * 2 * 64; i += kTcThreads) {
    const int k = i & 63, dir = (i >> 6) & 1, idx = i >> 7;
    float4 v = make_float4(1.f, 0.f, 1.f, 0.f);
    if (k < da) {
      const float* tab = gaff + idx * 4 * d + dir * 2 * d;
      v = make_float4(__ldg(tab + 2 * k), __ldg(tab + 2 * k + 1), __ldg(tab + 2 * (da + k)), __ldg(tab + 2 * (da + k) + 1));
    }
    sm.aff4[i] = v;
  }
  const unsigned char* wblob = A.blob + tc_affine_bytes(d, Lc);
  const size_t cb = tc_coupling_bytes(S), bl_off = tc_w1_bytes(S) + tc_wl_bytes(S);
  for (int i = tid; i < Lc * S.N2p; i += kTcThreads) {
    const int l = i / S.N2p, k = i % S.N2p;
    sm.bl[i] = __ldg(reinterpret_cast<const float*>(wblob + (size_t)l * cb + bl_off) + k);
  }
  auto init_a1 = [&](int t0, int nthreads) {     // zero + the constant-one columns that no epilogue thread owns (k >= 64)
    uint32_t* w = reinterpret_cast<uint32_t*>(sm.a1);
    const int words = (int)(tc_a1_bytes(S) / 4);
    for (int i = t0; i < words; i += nthreads) {
      const int kg = i / (kTcRows * 4), k0 = 8 * kg + 2 * (i & 3);
      uint32_t v = 0;
      if (kg >= 8) {
        if (k0 == da || k0 == da + 1) v |= 0x3F80u;
        if (k0 + 1 == da || k0 + 1 == da + 1) v |= 0x3F800000u;
      }
      w[i] = v;
    }
  };
  init_a1(tid, kTcThreads);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (*sm.tmem_slot != 0u) __trap();

  const long long tiles = (A.n + kTcRows - 1) / kTcRows;
  long long my_tiles = 0;
  if ((long long)blockIdx.x < tiles) my_tiles = (tiles - 1 - blockIdx.x) / gridDim.x + 1;
  const uint32_t total_uses = (uint32_t)(my_tiles * Lc);
  const size_t cbT = (size_t)S.N2p * Hp * 2 + nu_w1T_bytes(S);
  const bool fused = A.p != nullptr;

  if (warp >= kTcEpiWarps) {
    reg_dealloc<kTcRegsService>();
    const bool lead = elect_one();
    if (warp == kTcWarpMma) {
      // ================================ MMA issuer ===========================================================================
      const uint32_t idesc1 = umma_idesc(kTcRows, Hp), idesc2 = umma_idesc(kTcRows, S.N2p);
      const uint32_t idesc3 = umma_idesc(kTcRows, Hp / 2), idesc4 = umma_idesc(kTcRows, S.K1);
      for (uint32_t u = 0; u < total_uses; ++u) {
        const uint32_t par = u & 1;
        // G1: Hpre = A1 . W1^T
        {
          const uint64_t ad = umma_desc(smem_u32(sm.a1), kTcRows * 16, 128), bd = umma_desc(smem_u32(sm.w1), Hp * 16, 128);
          uint32_t a_lo = (uint32_t)ad, b_lo = (uint32_t)bd;
          const uint32_t a_hi = (uint32_t)(ad >> 32), b_hi = (uint32_t)(bd >> 32);
          const uint32_t astep = (2 * (kTcRows * 16)) >> 4, bstep = (uint32_t)(2 * (Hp * 16)) >> 4;
          mbar_wait(bar(kNuBarW1), par);
          mbar_wait(bar(kNuBarA1), par);
          tc_fence_after();
          const int ks = S.K1 / 16;
          if (lead) umma_ss<false>(0u, a_lo, a_hi, b_lo, b_hi, idesc1);
          for (int kk = 1; kk < ks; ++kk) {
            a_lo += astep; b_lo += bstep;
            if (lead) umma_ss<true>(0u, a_lo, a_hi, b_lo, b_hi, idesc1);
          }
          if (lead) umma_commit(bar(kNuBarG1));
        }
        // G2: U' = tanh(Hpre) . Wl'^T   (A from tensor memory, split K layout)
        {
          const uint64_t bd = umma_desc(smem_u32(sm.wl), (uint32_t)(Hp / 16) * S.N2p * 16, 128);
          uint32_t b_lo = (uint32_t)bd, acol = 0u;
          const uint32_t b_hi = (uint32_t)(bd >> 32), bstep = (uint32_t)(S.N2p * 16) >> 4;
          mbar_wait(bar(kNuBarWl), par);
          mbar_wait(bar(kNuBarHid), par);
          tc_fence_after();
          const int ks = Hp / 16;
          if (lead) umma_ts<false>(kNuColU, acol, b_lo, b_hi, idesc2);
          for (int s = 1; s < ks; ++s) {
            acol += 8; b_lo += bstep;
            if (lead) umma_ts<true>(kNuColU, acol, b_lo, b_hi, idesc2);
          }
          if (lead) umma_commit(bar(kNuBarG2));
        }
        // G3: dH = dU' . Wl'   (A = packed dU' in the U' columns; B = Wl'^T image; two N-halves)
        {
          mbar_wait(bar(kNuBarWlT), par);
          mbar_wait(bar(kNuBarDu), par);
          tc_fence_after();
          const int ks = S.N2p / 16;
          const uint32_t bstep = (uint32_t)(2 * (Hp * 16)) >> 4;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint64_t bd = umma_desc(smem_u32(sm.wlT) + (uint32_t)half * (uint32_t)(Hp / 2) * 16u, Hp * 16, 128);
            uint32_t b_lo = (uint32_t)bd;
            const uint32_t b_hi = (uint32_t)(bd >> 32);
            const uint32_t dcol = half ? (uint32_t)kNuColDhHi : (uint32_t)(Hp / 2);
            for (int s = 0; s < ks; ++s) {
              const uint32_t acol = (uint32_t)kNuColU + 32u * (uint32_t)(s >> 1) + 8u * (uint32_t)(s & 1);
              if (lead) { if (s == 0) umma_ts<false>(dcol, acol, b_lo, b_hi, idesc3); else umma_ts<true>(dcol, acol, b_lo, b_hi, idesc3); }
              b_lo += bstep;
            }
          }
          if (lead) umma_commit(bar(kNuBarG3));
        }
        // G4: dS = (dH (1 - tanh^2)) . W1   (A = packed, split K layout at [Hp/2, Hp); B = W1^T image)
        {
          const uint64_t bd = umma_desc(smem_u32(sm.w1T), (uint32_t)(Hp / 16) * S.K1 * 16, 128);
          uint32_t b_lo = (uint32_t)bd, acol = (uint32_t)(Hp / 2);
          const uint32_t b_hi = (uint32_t)(bd >> 32), bstep = (uint32_t)(S.K1 * 16) >> 4;
          mbar_wait(bar(kNuBarW1T), par);
          mbar_wait(bar(kNuBarDpre), par);
          tc_fence_after();
          const int ks = Hp / 16;
          if (lead) umma_ts<false>(kNuColU, acol, b_lo, b_hi, idesc4);
          for (int s = 1; s < ks; ++s) {
            acol += 8; b_lo += bstep;
            if (lead) umma_ts<true>(kNuColU, acol, b_lo, b_hi, idesc4);
          }
          if (lead) umma_commit(bar(kNuBarG4));
        }
      }
    } else if (warp == kTcWarpWeights) {
      // ================================ weight loader =========================================================================
      if (lead) {
        const uint32_t b1 = (uint32_t)tc_w1_bytes(S), b2 = (uint32_t)tc_wl_bytes(S), b3 = (uint32_t)((size_t)S.N2p * Hp * 2),
                       b4 = (uint32_t)nu_w1T_bytes(S);
        for (uint32_t u = 0; u < total_uses; ++u) {
          const int l = (int)(u % (uint32_t)Lc);
          const uint32_t prev = (u - 1) & 1;
          const unsigned char* w = wblob + (size_t)l * cb;
          const unsigned char* wT = A.blobT + (size_t)l * cbT;
          if (u > 0) mbar_wait_relaxed(bar(kNuBarG1), prev);
          mbar_expect_tx(bar(kNuBarW1), b1);
          tma_bulk_load(smem_u32(sm.w1), w, b1, bar(kNuBarW1));
          if (u > 0) {
            if (l == 0 && fused) mbar_wait_relaxed(bar(kNuBarBufFree), (uint32_t)((u / (uint32_t)Lc - 1) & 1));   // the buffer held the z tile
            else mbar_wait_relaxed(bar(kNuBarG2), prev);
          }
          mbar_expect_tx(bar(kNuBarWl), b2);
          tma_bulk_load(smem_u32(sm.wl), w + b1, b2, bar(kNuBarWl));
          if (l == 0) mbar_wait_relaxed(bar(kNuBarXRead), (uint32_t)((u / (uint32_t)Lc) & 1));   // the buffer still holds the x tile
          else mbar_wait_relaxed(bar(kNuBarG3), prev);
          mbar_expect_tx(bar(kNuBarWlT), b3);
          tma_bulk_load(smem_u32(sm.wlT), wT, b3, bar(kNuBarWlT));
          if (u > 0) mbar_wait_relaxed(bar(kNuBarG4), prev);
          mbar_expect_tx(bar(kNuBarW1T), b4);
          tma_bulk_load(smem_u32(sm.w1T), wT + b3, b4, bar(kNuBarW1T));
        }
      }
    } else if (warp == kTcWarpTiles) {
      // ================================ tile loader: x in, gradient out, through the Wl'^T buffer ================================
      if (lead) {
        for (long long p = 0; p < my_tiles; ++p) {
          const long long tile = (long long)blockIdx.x + p * gridDim.x;
          long long rows = A.n - tile * kTcRows;
          if (rows > kTcRows) rows = kTcRows;
          const uint32_t bytes = (uint32_t)(rows * d * 4);
          const long long off = tile * kTcRows * (long long)d;
          if (!fused || p == 0) {             // fused mode: later tiles are requested at the end of the previous iteration
            mbar_expect_tx(bar(kNuBarXFull), bytes);
            tma_bulk_load(smem_u32(sm.wlT), A.x + off, bytes, bar(kNuBarXFull));
          }
          if (fused) {
            // the momentum tile (and the latent tile, if it drifts) arrive under the last coupling: z into the Wl' buffer once
            // its last GEMM 2 has completed, p into the Wl'^T buffer once its last GEMM 3 has
            // (this lane follows the GEMM barriers phase by phase: a parity wait cannot tell phases two apart)
            mbar_expect_tx(bar(kNuBarPZFull), A.drift ? 2 * bytes : bytes);
            for (int l = 0; l < Lc; ++l) {
              const uint32_t par = (uint32_t)((p * Lc + l) & 1);
              mbar_wait_relaxed(bar(kNuBarG2), par);
              if (l == Lc - 1 && A.drift) tma_bulk_load(smem_u32(sm.wl), A.zw + off, bytes, bar(kNuBarPZFull));
              mbar_wait_relaxed(bar(kNuBarG3), par);
            }
            tma_bulk_load(smem_u32(sm.wlT), A.p + off, bytes, bar(kNuBarPZFull));
            mbar_wait_relaxed(bar(kNuBarPZOut), (uint32_t)(p & 1));
            tma_bulk_store(A.p + off, smem_u32(sm.wlT), bytes);
            if (A.drift) tma_bulk_store(A.zw + off, smem_u32(sm.wl), bytes);
            // the next x tile goes into the Wl'^T buffer as soon as the momentum store (the OLDER bulk group) has read it
            if (A.drift) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
            if (p + 1 < my_tiles) {
              const long long tile2 = tile + gridDim.x;
              long long rows2 = A.n - tile2 * kTcRows;
              if (rows2 > kTcRows) rows2 = kTcRows;
              const uint32_t bytes2 = (uint32_t)(rows2 * d * 4);
              mbar_expect_tx(bar(kNuBarXFull), bytes2);
              tma_bulk_load(smem_u32(sm.wlT), A.x + tile2 * kTcRows * (long long)d, bytes2, bar(kNuBarXFull));
            }
            tma_store_wait_read<0>();
            mbar_arrive(bar(kNuBarBufFree));
          } else {
            mbar_wait_relaxed(bar(kNuBarGOut), (uint32_t)(p & 1));
            tma_bulk_store(A.grad + off, smem_u32(sm.wlT), bytes);
            tma_store_wait_read<0>();
          }
        }
        tma_store_wait_all();
      }
    }
    __syncwarp();
  } else {
    // ================================ epilogue warps ==============================================================================
    reg_alloc<kTcRegsEpi>();
    const int r = tid & (kTcRows - 1), g = tid >> 7, q = (tid >> 5) & 3;
    const int e0 = g * kTcOwn;
    const uint32_t trow = (uint32_t)(q * 32) << 16;
    const bool flip = (Lc & 1) != 0;
    float* tile_row = reinterpret_cast<float*>(sm.wlT) + (size_t)r * d;
    uint32_t use = 0;
    float st[2][kTcOwn], gr[2][kTcOwn];
#ifdef NFMC_NU_TRACE
    long long* tr_p = (blockIdx.x == 0 && tid == 0) ? g_nu_trace : nullptr;
    int tr_n = 0;
#endif

    for (long long p = 0; p < my_tiles; ++p) {
      const long long tile = (long long)blockIdx.x + p * gridDim.x;
      const long long row_g = tile * kTcRows + r;
      // ---- x out of the tile buffer; U(x), grad U(x) -----------------------------------------------------------------------
      const float ld_inv_r = (g == 0 && row_g < A.n) ? __ldg(A.ld_inv + row_g) : 0.f;   // in flight while the tile lands
      NU_EV(0);
      mbar_wait_relaxed(bar(kNuBarXFull), (uint32_t)(p & 1));
      NU_EV(1);
      tc_row_read_half<0>(tile_row, d, da, e0, false, st[0]);
      tc_row_read_half<1>(tile_row, d, da, e0, false, st[1]);
      mbar_arrive(bar(kNuBarXRead));
      float* red = sm.red + (size_t)(p & 1) * 6 * kTcRows;      // the other set may still be read by slower warps of the last tile
      red[g * kTcRows + r] = nu_pot_partial(A.pot_kind, A.pot, da, e0, st[0], st[1]);
      if (g == 0) {
        red[4 * kTcRows + r] = st[0][0];
        red[5 * kTcRows + r] = da >= 2 ? st[0][1] : st[1][0];
      }
      tc_epi_barrier();
      {
        const float Ssum = red[r] + red[kTcRows + r] + red[2 * kTcRows + r] + red[3 * kTcRows + r];
        const float u_val = nu_pot_value_grad(A.pot_kind, A.pot, d, da, e0, Ssum, red[4 * kTcRows + r], red[5 * kTcRows + r], st[0], st[1],
                                              gr[0], gr[1]);
        if (g == 0 && row_g < A.n) A.value[row_g] = u_val - ld_inv_r;                    // neutra.py:62-64
      }
      // ---- backward sweep x -> z -----------------------------------------------------------------------------------------------
      nu_affine_unwind(sm.aff4, 0, e0, st[0], st[1], gr[0], gr[1]);
      NU_EV(2);
#pragma unroll 1
      for (int l = 0; l < Lc; ++l) {
        const uint32_t par = use & 1;
        const int src = (l & 1) == 0 ? 1 : 0;
        if (src) tc_write_a1(sm.a1, r, g, st[1]); else tc_write_a1(sm.a1, r, g, st[0]);
        fence_async_smem();
        tc_fence_before();
        mbar_arrive(bar(kNuBarA1));
        NU_EV(10);
        mbar_wait(bar(kNuBarG1), par);
        NU_EV(11);
        tc_fence_after();
        tc_epi1(trow, Hp, g);
        tc_fence_before();
        mbar_arrive(bar(kNuBarHid));
        NU_EV(12);
        mbar_wait(bar(kNuBarG2), par);
        NU_EV(13);
        tc_fence_after();
        if (src) nu_epi2(trow + kNuColU, sm.bl + (size_t)l * S.N2p, S.N2p, g, da, st[0], gr[0]);
        else nu_epi2(trow + kNuColU, sm.bl + (size_t)l * S.N2p, S.N2p, g, da, st[1], gr[1]);
        tc_fence_before();
        mbar_arrive(bar(kNuBarDu));
        NU_EV(14);
        mbar_wait(bar(kNuBarG3), par);
        NU_EV(15);
        tc_fence_after();
        nu_epi3(trow, Hp, g);
        tc_fence_before();
        mbar_arrive(bar(kNuBarDpre));
        NU_EV(16);
        mbar_wait(bar(kNuBarG4), par);
        NU_EV(17);
        tc_fence_after();
        {
          uint32_t v[16];
          tmem_ld16(trow + kNuColU + 16 * g, v);
          tmem_wait_ld(v);
#pragma unroll
          for (int i = 0; i < kTcOwn; ++i)
            if (e0 + i < da) { if (src) gr[1][i] += __uint_as_float(v[i]); else gr[0][i] += __uint_as_float(v[i]); }
        }
        tc_fence_before();
        nu_affine_unwind(sm.aff4, l + 1, e0, st[0], st[1], gr[0], gr[1]);
        NU_EV(18);
        use += 1;
      }
      if (fused) {
        // ---- leapfrog update in place on the staged momentum / latent tiles (logical order: flipped when Lc is odd) -----------
        float* zrow = reinterpret_cast<float*>(sm.wl) + (size_t)r * d;
        NU_EV(3);
        mbar_wait_relaxed(bar(kNuBarPZFull), (uint32_t)(p & 1));
        NU_EV(4);
        // In place on the staged tiles, in 16-byte units where the row layout allows (d/2 is even, so elements come in valid pairs;
        // the high half starts 8 bytes off a 16-byte boundary when d/2 % 4 == 2: pair, three quads, pair).  Shared-memory
        // wavefronts are what this stage costs -- rows are 4 d bytes apart, so a warp-wide access is spread over 32 rows --
        // and whole-row register arrays spilled to local memory (an L2 round trip at this carve-out): 16.7 k cycles per tile
        // in the first r02 timeline, 10.4 k with 8-byte units.
        auto unit2 = [&](int half, int c, float g0, float g1) {
          const int k = e0 + 2 * c;
          if (k >= da) return;
          const int pos = half * da + k;
          const int at = flip ? d - 2 - pos : pos;
          float2 pv = *reinterpret_cast<const float2*>(tile_row + at);
          float2 zv = A.drift ? *reinterpret_cast<const float2*>(zrow + at) : make_float2(0.f, 0.f);
          if (flip) { pv = make_float2(pv.y, pv.x); zv = make_float2(zv.y, zv.x); }
          pv.x = fmaf(-A.half_tau, g0, pv.x);                                       // hmc.py:51-53
          pv.y = fmaf(-A.half_tau, g1, pv.y);
          if (A.kicks == 2) { pv.x = fmaf(-A.half_tau, g0, pv.x); pv.y = fmaf(-A.half_tau, g1, pv.y); }
          if (A.drift) {                                                            // hmc.py:56-58
            float m0 = 1.f, m1 = 1.f;
            if (A.imd) { m0 = __ldg(A.imd + (flip ? d - 1 - pos : pos)); m1 = __ldg(A.imd + (flip ? d - 2 - pos : pos + 1)); }
            zv.x = fmaf(A.tau, A.imd ? pv.x * m0 : pv.x, zv.x);
            zv.y = fmaf(A.tau, A.imd ? pv.y * m1 : pv.y, zv.y);
          }
          *reinterpret_cast<float2*>(tile_row + at) = flip ? make_float2(pv.y, pv.x) : pv;
          if (A.drift) *reinterpret_cast<float2*>(zrow + at) = flip ? make_float2(zv.y, zv.x) : zv;
        };
        auto unit4 = [&](int half, int c, float g0, float g1, float g2, float g3) {       // unflipped, 16-byte aligned, c even / odd as planned
          const int k = e0 + 2 * c;
          if (k + 4 > da) { unit2(half, c, g0, g1); unit2(half, c + 1, g2, g3); return; }
          const int at = half * da + k;
          float4 pv = *reinterpret_cast<const float4*>(tile_row + at);
          float4 zv = A.drift ? *reinterpret_cast<const float4*>(zrow + at) : make_float4(0.f, 0.f, 0.f, 0.f);
          pv.x = fmaf(-A.half_tau, g0, pv.x); pv.y = fmaf(-A.half_tau, g1, pv.y);
          pv.z = fmaf(-A.half_tau, g2, pv.z); pv.w = fmaf(-A.half_tau, g3, pv.w);
          if (A.kicks == 2) {
            pv.x = fmaf(-A.half_tau, g0, pv.x); pv.y = fmaf(-A.half_tau, g1, pv.y);
            pv.z = fmaf(-A.half_tau, g2, pv.z); pv.w = fmaf(-A.half_tau, g3, pv.w);
          }
          if (A.drift) {
            float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
            if (A.imd) m = make_float4(__ldg(A.imd + at), __ldg(A.imd + at + 1), __ldg(A.imd + at + 2), __ldg(A.imd + at + 3));
            zv.x = fmaf(A.tau, A.imd ? pv.x * m.x : pv.x, zv.x); zv.y = fmaf(A.tau, A.imd ? pv.y * m.y : pv.y, zv.y);
            zv.z = fmaf(A.tau, A.imd ? pv.z * m.z : pv.z, zv.z); zv.w = fmaf(A.tau, A.imd ? pv.w * m.w : pv.w, zv.w);
          }
          *reinterpret_cast<float4*>(tile_row + at) = pv;
          if (A.drift) *reinterpret_cast<float4*>(zrow + at) = zv;
        };
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const float* gh = gr[half];
          const bool shifted = half == 1 && (da & 2) != 0;     // this half starts 8 bytes off a 16-byte boundary
          if (flip) {
#pragma unroll
            for (int c = 0; c < kTcOwn / 2; ++c) unit2(half, c, gh[2 * c], gh[2 * c + 1]);
          } else if (!shifted) {
#pragma unroll
            for (int c = 0; c < kTcOwn / 2; c += 2) { unit4(half, c, gh[2 * c], gh[2 * c + 1], gh[2 * c + 2], gh[2 * c + 3]); NU_EV(50 + c); }
          } else {
            unit2(half, 0, gh[0], gh[1]);
#pragma unroll
            for (int c = 1; c < kTcOwn / 2 - 1; c += 2) unit4(half, c, gh[2 * c], gh[2 * c + 1], gh[2 * c + 2], gh[2 * c + 3]);
            unit2(half, kTcOwn / 2 - 1, gh[kTcOwn - 2], gh[kTcOwn - 1]);
          }
          NU_EV(42 + half);
        }
        fence_async_smem();
        NU_EV(44);
        mbar_arrive(bar(kNuBarPZOut));
        NU_EV(5);
      } else {
        // ---- dU~/dz leaves through the tile buffer (logical order: flipped when Lc is odd) ----------------------------------------
        tc_row_write_half<0>(tile_row, d, da, e0, flip, gr[0]);
        tc_row_write_half<1>(tile_row, d, da, e0, flip, gr[1]);
        fence_async_smem();
        mbar_arrive(bar(kNuBarGOut));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(0u, 512);
}

// ---- elementwise pieces of the HMC step ----------------------------------------------------------------------------------------
// p = xi / sqrt(m) (hmc.py:100), kin0 = 1/2 sum p^2 m (hmc.py:103-106), z_work = z.  One warp per chain.
__global__ void neutra_tc_init_kernel(const float* __restrict__ xi, const float* __restrict__ imd, const float* __restrict__ z, float* __restrict__ p,
                                      float* __restrict__ zw, float* __restrict__ kin0, long long n, int d) {
  const int lane = threadIdx.x & 31;
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n; row += (long long)gridDim.x * (blockDim.x >> 5)) {
    float s = 0.f;
    for (int i = lane; i < d; i += 32) {
      const float m = imd ? __ldg(imd + i) : 1.f;
      float v = __ldg(xi + row * d + i);
      if (imd) v *= __fdiv_rn(1.f, sqrtf(m));
      p[row * d + i] = v;
      zw[row * d + i] = __ldg(z + row * d + i);
      s = fmaf(v * v, m, s);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) kin0[row] = 0.5f * s;
  }
}
// accept test (hmc.py:107-113), masked overwrite (mcmc/base.py:77), moments and counters of the post-accept state, sink
__global__ void __launch_bounds__(256) neutra_tc_accept_kernel(float* __restrict__ z, const float* __restrict__ zw, const float* __restrict__ p,
                                                               const float* __restrict__ imd, const float* __restrict__ u0, const float* __restrict__ u1,
                                                               const float* __restrict__ kin0, const float* __restrict__ unif, int adjusted,
                                                               long long n, int d, StatsArgs stats, float* __restrict__ sink_row) {
  extern __shared__ double acc_s[];      // [2 d] column sums of this CTA
  __shared__ unsigned int cnt_s[3];
  for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) acc_s[i] = 0.0;
  if (threadIdx.x < 3) cnt_s[threadIdx.x] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int per = (d + 31) / 32;        // <= 4 columns per lane for d <= 128
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  unsigned int n_acc = 0, n_rows = 0, n_bad = 0;
  for (long long row = (long long)blockIdx.x * nw + wid; row < n; row += (long long)gridDim.x * nw) {
    float s = 0.f;
    for (int i = lane; i < d; i += 32) {
      const float pv = __ldg(p + row * d + i), m = imd ? __ldg(imd + i) : 1.f;
      s = fmaf(pv * pv, m, s);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float h0 = __ldg(u0 + row) + __ldg(kin0 + row), h1 = __ldg(u1 + row) + 0.5f * s;
    const float log_acc = -h1 - (-h0);
    bool accept = true;
    if (adjusted) {
      accept = logf(__ldg(unif + row)) < log_acc;
      if (lane == 0 && !(fabsf(log_acc) <= 3.0e38f)) ++n_bad;
    }
    if (lane == 0) { ++n_rows; if (accept) ++n_acc; }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = lane + 32 * c;
      if (c < per && i < d) {
        const float v = accept ? __ldg(zw + row * d + i) : z[row * d + i];
        if (accept) z[row * d + i] = v;
        if (sink_row) sink_row[row * d + i] = v;
        s1[c] += v;
        s2[c] = fmaf(v, v, s2[c]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int i = lane + 32 * c;
    if (c < per && i < d) { atomicAdd(acc_s + i, (double)s1[c]); atomicAdd(acc_s + d + i, (double)s2[c]); }
  }
  if (lane == 0) {
    if (n_acc) atomicAdd(cnt_s + 0, n_acc);
    if (n_rows) atomicAdd(cnt_s + 1, n_rows);
    if (n_bad) atomicAdd(cnt_s + 2, n_bad);
  }
  __syncthreads();
  if (stats.sum_x && stats.sum_x2)
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
      atomicAdd(stats.sum_x + i, acc_s[i]);
      atomicAdd(stats.sum_x2 + i, acc_s[d + i]);
    }
  if (stats.counts && threadIdx.x < 3 && cnt_s[threadIdx.x]) atomicAdd(stats.counts + threadIdx.x, (unsigned long long)cnt_s[threadIdx.x]);
}

}  // namespace nfmc

using namespace nfmc;

namespace {
inline size_t al256(size_t v) { return (v + 255) & ~size_t(255); }

int nu_shape(const nfmc_realnvp_tc* flow, const void* blobT, int64_t blobT_bytes, TcShape& S, const char* who) {
  if (int e = tc_validate(flow, S, who)) return e;
  if (S.Hp % 32 != 0) return set_error(std::string(who) + ": the tensor-core NeuTra path needs a hidden width that is a multiple of 32");
  if (S.d % 4 != 0) return set_error(std::string(who) + ": the tensor-core NeuTra path needs d % 4 == 0");
  const int64_t want = (int64_t)S.Lc * ((int64_t)S.N2p * S.Hp * 2 + (int64_t)S.Hp * S.K1 * 2);
  if (!blobT || blobT_bytes != want) return set_error(std::string(who) + ": transposed-weights blob missing or of the wrong size");
  if (nu_smem_total(S) > 227 * 1024) return set_error(std::string(who) + ": shared-memory plan exceeds 227 KB (d x hidden too large)");
  return 0;
}
}  // namespace

#ifdef NFMC_NU_TRACE
extern "C" __attribute__((visibility("default"))) int nfmc_nu_trace_set(long long* buf) {
  return (int)cudaMemcpyToSymbol(g_nu_trace, &buf, sizeof(buf));
}
#endif

extern "C" int64_t nfmc_neutra_tc_transposed_bytes(int32_t d, int32_t n_coupling, int32_t hidden) {
  TcShape S;
  if (!tc_shape(d, n_coupling, hidden, S) || S.Hp % 32 != 0 || d % 4 != 0 || nu_smem_total(S) > 227 * 1024) return -1;
  return (int64_t)S.Lc * ((int64_t)S.N2p * S.Hp * 2 + (int64_t)S.Hp * S.K1 * 2);
}

extern "C" int64_t nfmc_neutra_tc_workspace_bytes(int32_t d, int64_t n) {
  return (int64_t)(3 * al256((size_t)n * d * 4) + 6 * al256((size_t)n * 4));
}

// inverse pass (z -> x, log|det|) followed by the unwind kernel; A carries the output mode (gradient tile or fused leapfrog)
static int nu_value_grad(const nfmc_potential* pot, const nfmc_realnvp_tc* flow, const void* blob_t, NuArgs& A, const float* z, float* x,
                         float* ld, float* u, int64_t n, void* stream) {
  if (int e = nfmc_flow_tc_pass(flow, 1, z, x, ld, n, stream)) return e;
  A.blob = static_cast<const unsigned char*>(flow->blob);
  A.blobT = static_cast<const unsigned char*>(blob_t);
  A.pot_kind = pot->kind; A.pot = pot_params(pot);
  A.x = x; A.ld_inv = ld; A.value = u; A.n = n;
  const size_t smem = nu_smem_total(A.S);
  const long long tiles = (n + kTcRows - 1) / kTcRows;
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    if (int e = check_cuda(cudaFuncSetAttribute(neutra_unwind_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024), "neutra_unwind_tc_kernel attribute")) return e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  neutra_unwind_tc_kernel<<<grid, kTcThreads, smem, (cudaStream_t)stream>>>(A);
  return check_cuda(cudaGetLastError(), "neutra_unwind_tc_kernel launch");
}

// U~(z) and dU~/dz on the tensor cores (x [n, d] and ld [n] are scratch outputs: x = T^-1(z), log|det dx/dz|)
extern "C" int nfmc_neutra_potential_tc(const nfmc_potential* pot, const nfmc_realnvp_tc* flow, const void* blob_t, int64_t blob_t_bytes,
                                        const float* z, float* x, float* ld, float* u, float* grad, int64_t n, void* stream) {
  if (int e = validate_pot(pot)) return e;
  NuArgs A{};
  if (int e = nu_shape(flow, blob_t, blob_t_bytes, A.S, "neutra_potential_tc")) return e;
  if (pot->d != flow->d) return set_error("neutra_potential_tc: potential and flow event sizes differ");
  if (!z || !x || !ld || !u || !grad || n < 1) return set_error("neutra_potential_tc: bad arguments");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(grad) & 15))
    return set_error("neutra_potential_tc: x / grad must be 16-byte aligned");
  A.grad = grad;
  return nu_value_grad(pot, flow, blob_t, A, z, x, ld, u, n, stream);
}

// T NeuTra-HMC iterations on the tensor-core path (same contract as nfmc_neutra_hmc_steps; z in logical order).
extern "C" int nfmc_neutra_hmc_steps_tc(const nfmc_potential* pot, const nfmc_realnvp_tc* flow, const void* blob_t, int64_t blob_t_bytes,
                                        float* z, int64_t n, int32_t n_steps, float step_size, int32_t n_leapfrog,
                                        const float* inv_mass_diag, int32_t adjusted, const nfmc_rng* rng, int64_t chain0,
                                        const nfmc_stats* stats, const nfmc_sink* sink, void* workspace, int64_t workspace_bytes,
                                        void* stream) {
  if (!rng || !z || n < 1 || n_steps < 1 || n_leapfrog < 1) return set_error("neutra_hmc_steps_tc: bad arguments");
  if ((rng->normals == nullptr) != (rng->uniforms == nullptr) && adjusted)
    return set_error("neutra_hmc_steps_tc: inject both normals and uniforms, or neither");
  TcShape S;
  if (int e = validate_pot(pot)) return e;
  if (int e = nu_shape(flow, blob_t, blob_t_bytes, S, "neutra_hmc_steps_tc")) return e;
  if (pot->d != flow->d) return set_error("neutra_hmc_steps_tc: potential and flow event sizes differ");
  const int d = flow->d;
  if (!workspace || workspace_bytes < nfmc_neutra_tc_workspace_bytes(d, n)) return set_error("neutra_hmc_steps_tc: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char* w = static_cast<unsigned char*>(workspace);
  const size_t nd = al256((size_t)n * d * 4), nn = al256((size_t)n * 4);
  float* p = reinterpret_cast<float*>(w);
  float* zw = reinterpret_cast<float*>(w + nd);
  float* x = reinterpret_cast<float*>(w + 2 * nd);
  float* ld = reinterpret_cast<float*>(w + 3 * nd);
  float* u0 = reinterpret_cast<float*>(w + 3 * nd + nn);
  float* u1 = reinterpret_cast<float*>(w + 3 * nd + 2 * nn);
  float* kin0 = reinterpret_cast<float*>(w + 3 * nd + 3 * nn);
  float* unif = reinterpret_cast<float*>(w + 3 * nd + 4 * nn);
  const int row_grid = (int)std::min<long long>((n + 7) / 8, 8ll * sm_count());
  const float half_tau = step_size / 2;
  const StatsArgs st{stats ? stats->sum_x : nullptr, stats ? stats->sum_x2 : nullptr, stats ? stats->counts : nullptr};
  for (int k = 0; k < n_steps; ++k) {
    const float* xi;
    const float* uu;
    if (rng->normals) {
      xi = rng->normals + (size_t)k * n * d;
      uu = rng->uniforms ? rng->uniforms + (size_t)k * n : nullptr;
    } else {
      nfmc_rng r2 = *rng;
      r2.step0 = rng->step0 + (uint64_t)k;
      if (int e = nfmc_rng_fill(&r2, 0, chain0, d, n, 1, x, unif, stream)) return e;     // x is free until the first inverse pass
      xi = x; uu = unif;
    }
    neutra_tc_init_kernel<<<row_grid, 256, 0, s>>>(xi, inv_mass_diag, z, p, zw, kin0, n, d);
    for (int l = 0; l <= n_leapfrog; ++l) {
      // gradient at the current point with the leapfrog update fused into the kernel's output stage: one half-kick at the two
      // ends of the trajectory, two in between (consecutive half-kicks of hmc.py:68-71 share the gradient), then the drift
      NuArgs A{};
      A.S = S;
      A.p = p; A.zw = zw; A.imd = inv_mass_diag; A.half_tau = half_tau; A.tau = step_size;
      A.kicks = (l == 0 || l == n_leapfrog) ? 1 : 2;
      A.drift = l < n_leapfrog ? 1 : 0;
      if (int e = nu_value_grad(pot, flow, blob_t, A, zw, x, ld, l == 0 ? u0 : u1, n, stream)) return e;
    }
    float* sink_row = nullptr;
    if (sink && sink->samples) {
      const int64_t th = sink->thinning > 0 ? sink->thinning : 1;
      const int64_t idx = sink->seen0 + k;
      if (idx % th == 0) sink_row = sink->samples + (size_t)(idx / th - (sink->seen0 + th - 1) / th) * n * d;
    }
    neutra_tc_accept_kernel<<<row_grid, 256, (size_t)2 * d * sizeof(double), s>>>(z, zw, p, inv_mass_diag, u0, u1, kin0, uu, adjusted, n, d, st,
                                                                                  sink_row);
  }
  return check_cuda(cudaGetLastError(), "neutra_hmc_steps_tc launch");
}
