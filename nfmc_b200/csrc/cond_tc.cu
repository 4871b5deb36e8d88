// cond_tc.cu -- RealNVP passes with the conditioner MLP on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// For wide conditioners (M = 2 linear layers, H a multiple of 16, 16 <= H <= 256, even d <= 128) the two GEMMs of a
// coupling layer
//       Hpre[128 x H]  = S[128 x d/2] . W1^T          (K = d/2, zero padded to 64)
//       U   [128 x 2db] = tanh(Hpre + b1)[128 x H] . Wl^T   (K = H)
// run as tcgen05.mma (kind::f16, bf16 operands, fp32 accumulators in tensor memory) on tiles of 128 chains.  One CTA
// of 128 threads owns a tile; thread t owns chain row t, keeps the whole chain state in registers (fp32) and does the
// fused epilogues straight out of TMEM with tcgen05.ld (its TMEM lane = its row):
//   epilogue 1: + b1, tanh, -> bf16 -> shared memory in the UMMA K-major layout (A operand of the second GEMM)
//   epilogue 2: + bl, alpha = exp(c + u_a/2) + m, beta = u_b/2, target <- alpha*target + beta (or the inverse),
//               log-det accumulation.
// The weights of a coupling layer (bf16, pre-arranged on the host in the exact shared-memory image) are staged by the
// TMA engine (cp.async.bulk, mbarrier complete_tx) while the previous epilogue runs.
//
// Same specification as the fp32 path (oracle/realnvp_ref.py); parity tolerance is the bf16 one of the north star
// (rtol 1e-2).  Replaces flow.bijection.forward / inverse / flow.log_prob for wide flows (neutra.py:60, jump.py:218).
//
// UMMA operand layout used everywhere (SWIZZLE_NONE, K-major): an [R rows x K] bf16 operand is stored as
// [K/8][R][8]: 8 consecutive k of one row are 16 contiguous bytes ("core matrix" = 8 rows x 16 B = 128 B contiguous),
// so  SBO (next 8-row group) = 128 B  and  LBO (next 8-column group) = R * 16 B.
#include <cuda_bf16.h>
#include "host_common.cuh"

namespace nfmc {

constexpr int kTcRows = 128;      // chains per tile = UMMA_M
constexpr int kTcK1 = 64;         // padded source width (d/2 <= 64)
constexpr int kTcHalf = 64;       // registers per half of the chain state
constexpr int kTcTmemCols = 512;
constexpr int kTcCol2 = 256;      // TMEM column of the second accumulator

struct TcArgs {
  const unsigned char* blob;  // packed: affines fp32 | const | per coupling { W1 image, b1, Wl image, bl }
  int d, Lc, H, N2p;
  int mode;                   // 0 forward, 1 inverse, 2 log_prob
  const float* in;
  float* out;
  float* aux;
  long long n;
};

__host__ __device__ inline size_t tc_coupling_bytes(int H, int N2p) {
  return (size_t)kTcK1 * H * 2 + (size_t)H * 4 + (size_t)N2p * H * 2 + (size_t)N2p * 4;
}
__host__ __device__ inline size_t tc_affine_bytes(int d, int Lc) { return ((size_t)(Lc + 1) * 4 * d + 4) * 4; }

// ---- PTX wrappers --------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// shared-memory matrix descriptor: SWIZZLE_NONE, K-major (see file header)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address      bits [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16; // leading byte off.  bits [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32; // stride byte off.   bits [32,46)
  d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
  return d;                                         // layout_type (bits 61..63) = 0: no swizzle
}
// instruction descriptor, kind::f16: D = f32, A = B = bf16, both K-major, M x N
__device__ __forceinline__ uint32_t umma_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// tanh of two values at once, straight in the bf16 the next GEMM consumes: one MUFU op per pair instead of two
// (the epilogue is MUFU-bound), and no separate convert/pack.  a -> low half, b -> high half.
__device__ __forceinline__ uint32_t tanh_bf16x2(float a, float b) {
  uint32_t p, y;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(b), "f"(a));
  asm("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(p));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---- the kernel ----------------------------------------------------------------------------------------------------
// 512 threads per CTA: thread (r, g) = (tid % 128, tid / 128) works on chain row r of the tile and owns the 16 elements
// [16 g, 16 g + 16) of each half of that chain (fp32, registers).  Its warp's TMEM lane quarter is (tid % 128) / 32, so
// the four column groups of a row read disjoint column ranges of the same TMEM lane: 16 warps share the epilogue work.
// dynamic smem: [A1: 128 x 64 bf16 = 16 KB][hid: 128 x H bf16][weights of one coupling][row reductions][mbarriers]
constexpr int kTcGroups = 4;
constexpr int kTcOwn = kTcHalf / kTcGroups;   // 16 elements per half per thread
constexpr int kTcThreads = kTcRows * kTcGroups;

// MINB = 2: two CTAs per SM (<= 64 registers per thread, 256 TMEM columns each) so that one tile's epilogue overlaps the
// other tile's MMAs / TMA waits; possible when H + N2p <= 256 and the shared-memory plan fits twice.
template <int MINB>
__global__ void __launch_bounds__(kTcThreads, MINB) flow_tc_kernel(const TcArgs A) {
  const uint32_t tmem_cols = (MINB == 2) ? 256u : (uint32_t)kTcTmemCols;
  const uint32_t col2 = (MINB == 2) ? (uint32_t)A.H : (uint32_t)kTcCol2;   // TMEM column of the second accumulator
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = tid & (kTcRows - 1), g = tid >> 7;
  const int d = A.d, da = d / 2, H = A.H, N2p = A.N2p, Lc = A.Lc;
  unsigned char* sA1 = smem;
  unsigned char* sHid = sA1 + kTcRows * kTcK1 * 2;
  unsigned char* sW = sHid + (size_t)kTcRows * H * 2;
  const size_t wbytes = tc_coupling_bytes(H, N2p);
  float* sRed = reinterpret_cast<float*>(sW + ((wbytes + 15) & ~size_t(15)));      // [2][4][128]
  float* sAff = sRed + 2 * kTcGroups * kTcRows;                                     // (Lc+1)*4*d + 4 floats
  const int n_aff = (Lc + 1) * 4 * d + 4;
  float* sBias = sAff + ((n_aff + 3) & ~3);                                         // Lc x {b1[H], bl[N2p]}: every coupling's biases
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + (size_t)Lc * (H + N2p));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  // weights of a coupling arrive as two TMA transactions -- the W1 image and the Wl image -- each re-issued for the NEXT
  // coupling as soon as the GEMM that reads it has completed, i.e. underneath the epilogues
  const uint32_t bar_w1 = smem_u32(bars), bar_mma = smem_u32(bars + 1), bar_wl = smem_u32(bars + 2);
  const size_t w1_bytes = (size_t)kTcK1 * H * 2, wl_off = w1_bytes + (size_t)H * 4, wl_bytes = (size_t)N2p * H * 2;

  if (tid == 0) {
    mbar_init(bar_w1, 1);
    mbar_init(bar_wl, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_row = tmem_base + ((uint32_t)((r >> 5) * 32) << 16);   // this warp's 32-lane quarter

  for (int i = tid; i < n_aff; i += kTcThreads) sAff[i] = __ldg(reinterpret_cast<const float*>(A.blob) + i);
  const unsigned char* wblob = A.blob + tc_affine_bytes(d, Lc);
  for (int i = tid; i < Lc * (H + N2p); i += kTcThreads) {
    const int l = i / (H + N2p), k = i % (H + N2p);
    const unsigned char* cb = wblob + (size_t)l * wbytes;
    sBias[i] = __ldg(reinterpret_cast<const float*>(k < H ? cb + w1_bytes : cb + wl_off + wl_bytes) + (k < H ? k : k - H));
  }
  __syncthreads();
  const float* aff = sAff;
  const float log_const = aff[(Lc + 1) * 4 * d];
  const bool inv = (A.mode == 1);
  const bool flip = (Lc & 1) != 0;
  const uint32_t idesc1 = umma_idesc(kTcRows, H), idesc2 = umma_idesc(kTcRows, N2p);
  uint32_t ph_w1 = 0, ph_wl = 0, ph_mma = 0;   // weight phases are tracked by the issuing thread only
  const int e0 = g * kTcOwn;   // first owned element index within a half

  const long long tiles = (A.n + kTcRows - 1) / kTcRows;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long row_raw = tile * kTcRows + r;
    const bool active = row_raw < A.n;
    const long long row = active ? row_raw : A.n - 1;
    // ---- this thread's share of the chain, fp32, in registers (physical order: flipped latent when Lc is odd) ----------
    float lo[kTcOwn], hi[kTcOwn];
    {
      const float* src = A.in + row * (long long)d;
      const bool fl = inv && flip;
      if (!fl && (d & 3) == 0 && (da & 1) == 0 && e0 + kTcOwn <= da) {
        // 64 contiguous bytes per half: 4 x 128-bit for the low half, 8 x 64-bit for the high half (da*4 is 8-byte aligned)
        const float4* pl = reinterpret_cast<const float4*>(src + e0);
        const float2* ph = reinterpret_cast<const float2*>(src + da + e0);
#pragma unroll
        for (int q = 0; q < kTcOwn / 4; ++q) {
          const float4 v = __ldg(pl + q);
          lo[4 * q] = v.x; lo[4 * q + 1] = v.y; lo[4 * q + 2] = v.z; lo[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int q = 0; q < kTcOwn / 2; ++q) {
          const float2 v = __ldg(ph + q);
          hi[2 * q] = v.x; hi[2 * q + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int q = 0; q < kTcOwn; ++q) {
          const int k = e0 + q;
          lo[q] = (k < da) ? __ldg(src + (fl ? d - 1 - k : k)) : 0.f;
          hi[q] = (k < da) ? __ldg(src + (fl ? d - 1 - (da + k) : da + k)) : 0.f;
        }
      }
    }
    float ld = 0.f;
    const int n_ops = 2 * Lc + 1;
    // first coupling's weights: issue the TMA bulk copy now (the buffer is free: the previous tile has finished with it)
    const int first_l = inv ? Lc - 1 : 0;
    const bool more_tiles = tile + gridDim.x < tiles;
    if (tid == 0 && Lc > 0 && tile == (long long)blockIdx.x) {       // later tiles: prefetched by the previous tile
      mbar_expect_tx(bar_w1, (uint32_t)w1_bytes);
      tma_bulk_load(smem_u32(sW), wblob + (size_t)first_l * wbytes, (uint32_t)w1_bytes, bar_w1);
      mbar_expect_tx(bar_wl, (uint32_t)wl_bytes);
      tma_bulk_load(smem_u32(sW + wl_off), wblob + (size_t)first_l * wbytes + wl_off, (uint32_t)wl_bytes, bar_wl);
    }
#pragma unroll 1
    for (int i = 0; i < n_ops; ++i) {
      const int op = inv ? n_ops - 1 - i : i;
      if ((op & 1) == 0) {
        // ---- elementwise affine op>>1 (forward {alpha, beta} or inverse {1/alpha, -beta/alpha}: the same fma) -----------
        const float2* tab = reinterpret_cast<const float2*>(aff + (op >> 1) * 4 * d + (inv ? 2 * d : 0));
#pragma unroll
        for (int q = 0; q < kTcOwn; ++q) {
          const int k = e0 + q;
          if (k < da) {
            const float2 pl = tab[k], ph = tab[da + k];
            lo[q] = fmaf(pl.x, lo[q], pl.y);
            hi[q] = fmaf(ph.x, hi[q], ph.y);
          }
        }
        continue;
      }
      // ---- coupling l: source half S, target half T ----------------------------------------------------------------------
      const int l = op >> 1;
      const bool src_is_hi = (l & 1) == 0;
      // the coupling whose weights are staged next: the next one of this tile, else the first one of this CTA's next tile
      const int next_l = (i + 2 < n_ops) ? ((inv ? n_ops - 1 - (i + 2) : i + 2) >> 1) : (more_tiles ? first_l : -1);
      // A operand of GEMM 1: this thread's 16 source values as bf16 into k-groups 2g, 2g+1 of the [K1/8][128][8] image
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        uint4 v;
        const int q0 = h2 * 8;
        v.x = pack_bf16(src_is_hi ? hi[q0 + 0] : lo[q0 + 0], src_is_hi ? hi[q0 + 1] : lo[q0 + 1]);
        v.y = pack_bf16(src_is_hi ? hi[q0 + 2] : lo[q0 + 2], src_is_hi ? hi[q0 + 3] : lo[q0 + 3]);
        v.z = pack_bf16(src_is_hi ? hi[q0 + 4] : lo[q0 + 4], src_is_hi ? hi[q0 + 5] : lo[q0 + 5]);
        v.w = pack_bf16(src_is_hi ? hi[q0 + 6] : lo[q0 + 6], src_is_hi ? hi[q0 + 7] : lo[q0 + 7]);
        *reinterpret_cast<uint4*>(sA1 + ((size_t)(2 * g + h2) * kTcRows + r) * 16) = v;
      }
      fence_async_smem();      // generic-proxy writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        mbar_wait(bar_w1, ph_w1);                                 // W1 image has landed (TMA complete_tx)
        ph_w1 ^= 1;
        const uint32_t a0 = smem_u32(sA1), b0 = smem_u32(sW);     // W1 image: [K1/8][H][8]
#pragma unroll 1
        for (int kk = 0; kk < kTcK1 / 16; ++kk)
          umma_f16(tmem_base, umma_desc(a0 + kk * 2 * (kTcRows * 16), kTcRows * 16, 128),
                   umma_desc(b0 + kk * 2 * (H * 16), H * 16, 128), idesc1, kk > 0);
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, ph_mma);
      ph_mma ^= 1;
      tc_fence_after();
      if (tid == 0 && next_l >= 0) {                              // GEMM 1 is done with the W1 image: stage the next one
        mbar_expect_tx(bar_w1, (uint32_t)w1_bytes);
        tma_bulk_load(smem_u32(sW), wblob + (size_t)next_l * wbytes, (uint32_t)w1_bytes, bar_w1);
      }
      // ---- epilogue 1: hid = tanh(Hpre + b1) -> bf16 -> A operand image of GEMM 2: [H/8][128][8]; 16-column chunks are
      //      dealt round-robin to the four column groups -------------------------------------------------------------------
      {
        const float* b1 = sBias + (size_t)l * (H + N2p);
#pragma unroll 1
        for (int c = g; c < H / 16; c += kTcGroups) {
          float v[16];
          tmem_ld16(tmem_row + c * 16, v);
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] += b1[c * 16 + q];
          uint4 w0, w1;
          w0.x = tanh_bf16x2(v[0], v[1]); w0.y = tanh_bf16x2(v[2], v[3]); w0.z = tanh_bf16x2(v[4], v[5]); w0.w = tanh_bf16x2(v[6], v[7]);
          w1.x = tanh_bf16x2(v[8], v[9]); w1.y = tanh_bf16x2(v[10], v[11]); w1.z = tanh_bf16x2(v[12], v[13]); w1.w = tanh_bf16x2(v[14], v[15]);
          *reinterpret_cast<uint4*>(sHid + ((size_t)(2 * c) * kTcRows + r) * 16) = w0;
          *reinterpret_cast<uint4*>(sHid + ((size_t)(2 * c + 1) * kTcRows + r) * 16) = w1;
        }
      }
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        mbar_wait(bar_wl, ph_wl);                                 // Wl image has landed
        ph_wl ^= 1;
        const uint32_t a0 = smem_u32(sHid), b0 = smem_u32(sW + wl_off);  // Wl image [H/8][N2p][8]
#pragma unroll 1
        for (int kk = 0; kk < H / 16; ++kk)
          umma_f16(tmem_base + col2, umma_desc(a0 + kk * 2 * (kTcRows * 16), kTcRows * 16, 128),
                   umma_desc(b0 + kk * 2 * (N2p * 16), N2p * 16, 128), idesc2, kk > 0);
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, ph_mma);
      ph_mma ^= 1;
      tc_fence_after();
      if (tid == 0 && next_l >= 0) {                              // GEMM 2 is done with the Wl image: stage the next one
        mbar_expect_tx(bar_wl, (uint32_t)wl_bytes);
        tma_bulk_load(smem_u32(sW + wl_off), wblob + (size_t)next_l * wbytes + wl_off, (uint32_t)wl_bytes, bar_wl);
      }
      // ---- epilogue 2: (u_a, u_b) = U + bl -> affine transform of this thread's 16 targets (chunks 2g, 2g+1), log-det ----
      {
        const float* bl = sBias + (size_t)l * (H + N2p) + H;
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int c = 2 * g + h2;                       // 16 columns = 8 targets (u_a, u_b interleaved)
          if (c * 16 < N2p) {
            float v[16];
            tmem_ld16(tmem_row + col2 + c * 16, v);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int t = c * 8 + q;
              if (t < da) {
                const float ua = v[2 * q] + bl[2 * t], ub = v[2 * q + 1] + bl[2 * t + 1];
                const float al = __expf(kLogOneMinusM + 0.5f * ua) + kMinScale, be = 0.5f * ub;
                const float ra = __fdividef(1.f, al);
                const float rr = inv ? ra : al, ss = inv ? -be * ra : be;
                if (src_is_hi) lo[h2 * 8 + q] = fmaf(rr, lo[h2 * 8 + q], ss);
                else hi[h2 * 8 + q] = fmaf(rr, hi[h2 * 8 + q], ss);
                ld += __logf(al);
              }
            }
          }
        }
      }
      // every thread has finished reading TMEM before the next coupling's GEMM 1 overwrites the accumulators
      tc_fence_before();
      __syncthreads();
    }
    // ---- results: the four column groups of a row combine their log-det / base-density shares through shared memory ------
    float sq = 0.f;
    if (A.mode == 2) {
#pragma unroll
      for (int q = 0; q < kTcOwn; ++q) sq = fmaf(lo[q], lo[q], fmaf(hi[q], hi[q], sq));
    }
    sRed[g * kTcRows + r] = ld;
    sRed[(kTcGroups + g) * kTcRows + r] = sq;
    if (active && A.out) {
      float* dst = A.out + row * (long long)d;
      const bool fl = !inv && flip;
      if (!fl && (d & 3) == 0 && (da & 1) == 0 && e0 + kTcOwn <= da) {
        float4* pl = reinterpret_cast<float4*>(dst + e0);
        float2* ph = reinterpret_cast<float2*>(dst + da + e0);
#pragma unroll
        for (int q = 0; q < kTcOwn / 4; ++q) pl[q] = make_float4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
#pragma unroll
        for (int q = 0; q < kTcOwn / 2; ++q) ph[q] = make_float2(hi[2 * q], hi[2 * q + 1]);
      } else {
#pragma unroll
        for (int q = 0; q < kTcOwn; ++q) {
          const int k = e0 + q;
          if (k < da) {
            dst[fl ? d - 1 - k : k] = lo[q];
            dst[fl ? d - 1 - (da + k) : da + k] = hi[q];
          }
        }
      }
    }
    __syncthreads();
    if (g == 0 && active && A.aux) {
      float res = sRed[r] + sRed[kTcRows + r] + sRed[2 * kTcRows + r] + sRed[3 * kTcRows + r] + log_const;
      if (inv) res = -res;
      if (A.mode == 2) {
        const float s = sRed[(kTcGroups + 0) * kTcRows + r] + sRed[(kTcGroups + 1) * kTcRows + r] +
                        sRed[(kTcGroups + 2) * kTcRows + r] + sRed[(kTcGroups + 3) * kTcRows + r];
        res += -0.5f * s - 0.5f * (float)d * 1.8378770664093453f;
      }
      A.aux[row] = res;
    }
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace nfmc

using namespace nfmc;

extern "C" int64_t nfmc_realnvp_tc_blob_bytes(int32_t d, int32_t n_coupling, int32_t hidden) {
  const int N2p = ((d - d / 2) * 2 + 15) & ~15;
  return (int64_t)(tc_affine_bytes(d, n_coupling) + (size_t)n_coupling * tc_coupling_bytes(hidden, N2p));
}

extern "C" int nfmc_flow_tc_pass(const nfmc_realnvp_tc* flow, int32_t mode, const float* in, float* out, float* aux,
                                 int64_t n, void* stream) {
  if (!flow || !flow->blob || !in || n < 1) return set_error("flow_tc_pass: bad arguments");
  const int d = flow->d, H = flow->hidden, Lc = flow->n_coupling;
  if (d < 2 || d > 128 || (d & 1)) return set_error("flow_tc_pass: tensor-core path needs even d <= 128");
  if (H < 16 || H > 256 || (H & 15)) return set_error("flow_tc_pass: hidden width must be a multiple of 16 in [16, 256]");
  if (mode < 0 || mode > 2) return set_error("flow_tc_pass: mode must be 0 (forward), 1 (inverse) or 2 (log_prob)");
  if (flow->blob_bytes != nfmc_realnvp_tc_blob_bytes(d, Lc, H)) return set_error("flow_tc_pass: blob_bytes mismatch");
  TcArgs A;
  A.blob = static_cast<const unsigned char*>(flow->blob);
  A.d = d; A.Lc = Lc; A.H = H; A.N2p = ((d - d / 2) * 2 + 15) & ~15;
  A.mode = mode; A.in = in; A.out = out; A.aux = aux; A.n = n;
  const size_t wbytes = (tc_coupling_bytes(H, A.N2p) + 15) & ~size_t(15);
  const size_t smem = (size_t)kTcRows * kTcK1 * 2 + (size_t)kTcRows * H * 2 + wbytes + 2 * kTcGroups * kTcRows * sizeof(float) +
                      (((size_t)(Lc + 1) * 4 * d + 4 + 3) & ~size_t(3)) * sizeof(float) + (size_t)Lc * (H + A.N2p) * sizeof(float) + 64;
  if (smem > 227 * 1024) return set_error("flow_tc_pass: shared-memory plan exceeds 227 KB");
  const long long tiles = (n + kTcRows - 1) / kTcRows;
  const bool two = (H + A.N2p <= 256) && (2 * (smem + 1024) <= 227 * 1024);
  const long long cap = (long long)sm_count() * (two ? 2 : 1);
  const int grid = (int)(tiles < cap ? tiles : cap);
  if (two) {
    cudaFuncSetAttribute(flow_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    flow_tc_kernel<2><<<grid, kTcThreads, smem, (cudaStream_t)stream>>>(A);
  } else {
    cudaFuncSetAttribute(flow_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    flow_tc_kernel<1><<<grid, kTcThreads, smem, (cudaStream_t)stream>>>(A);
  }
  return check_cuda(cudaGetLastError(), "flow_tc_kernel launch");
}
