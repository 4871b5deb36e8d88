// tc_common.cuh -- building blocks of the tensor-core RealNVP kernels (tc_flow.cu, tc_jump.cu): PTX wrappers for
// tcgen05 / TMEM / TMA / mbarrier, the packed-flow shape, the shared-memory plan, the control-warp issue sequence and
// the epilogue-warp side of a pipelined two-tile flow pass.  The design is described at the top of tc_flow.cu.
#pragma once
#include <cuda_bf16.h>
#include "host_common.cuh"

namespace nfmc {

constexpr int kTcRows = 128;                 // chains per tile = UMMA_M
constexpr int kTcGroups = 4;                 // column groups per chain row
constexpr int kTcOwn = 16;                   // elements per half per thread
constexpr int kTcEpiWarps = 16;
constexpr int kTcEpiThreads = kTcEpiWarps * 32;
constexpr int kTcThreads = kTcEpiThreads + 32;   // + the control warp
constexpr int kTcRegion = 256;               // TMEM columns per tile
constexpr int kTcUCol = 128;                 // column of the U accumulator inside a tile's region
constexpr int kTcMaxSlots = 16;              // weight buffers per image kind (resident mode: one per coupling)
constexpr float kLog2e = 1.4426950408889634f;

// Packed flow shape.  nb1 / nbl: number of shared-memory buffers for the W1 / Wl images; == Lc means every coupling's
// image stays resident for the whole kernel, otherwise the buffers form a ring refilled by the control warp.
struct TcShape {
  int d, Lc, Hp, N2p, K1;
  int nb1, nbl;
};

__host__ __device__ inline size_t tc_w1_bytes(const TcShape& S) { return (size_t)S.K1 * S.Hp * 2; }
__host__ __device__ inline size_t tc_wl_bytes(const TcShape& S) { return (size_t)S.Hp * S.N2p * 2; }
__host__ __device__ inline size_t tc_coupling_bytes(const TcShape& S) { return tc_w1_bytes(S) + tc_wl_bytes(S) + (size_t)S.N2p * 4; }
__host__ __device__ inline size_t tc_affine_bytes(int d, int Lc) { return ((size_t)(Lc + 1) * 4 * d + 4) * 4; }
__host__ __device__ inline size_t tc_a1_bytes(const TcShape& S) { return (size_t)(S.K1 < 64 ? 64 : S.K1) * kTcRows * 2; }

inline bool tc_shape(int d, int Lc, int hidden, TcShape& S) {
  if (d < 2 || d > 128 || (d & 1) || Lc < 1 || hidden < 16 || hidden > 256 || (hidden & 15)) return false;
  S.d = d; S.Lc = Lc; S.Hp = hidden;
  S.N2p = ((d - d / 2) * 2 + 15) & ~15;
  S.K1 = (d / 2 + 2 + 15) & ~15;
  S.nb1 = S.nbl = 1;
  return true;
}

// shared-memory carve-up (bytes, in this order): A1[2] | W1 slots | Wl slots | bl' [Lc][N2p] | affines | row reductions
// [2 tiles][2][4][128] | extra (caller) | mbarriers | tmem slot
struct TcSmem {
  unsigned char* a1[2];
  unsigned char* w1;
  unsigned char* wl;
  float* bl;
  float* aff;
  float* red;
  unsigned char* extra;
  uint64_t* bars;
  uint32_t* tmem_slot;
};
constexpr int kTcBarG1 = 0, kTcBarG2 = 2, kTcBarA1 = 4, kTcBarHid = 6, kTcBarW1 = 8, kTcBarWl = 8 + kTcMaxSlots,
              kTcNumBars = 8 + 2 * kTcMaxSlots;

__host__ __device__ inline size_t tc_smem_fixed(const TcShape& S, size_t extra) {
  const size_t aff = (((size_t)(S.Lc + 1) * 4 * S.d + 4 + 3) & ~size_t(3)) * 4;
  return 2 * tc_a1_bytes(S) + (size_t)S.Lc * S.N2p * 4 + aff + (size_t)2 * 2 * kTcGroups * kTcRows * 4 + ((extra + 15) & ~size_t(15)) +
         (size_t)kTcNumBars * 8 + 16;
}
__host__ __device__ inline size_t tc_smem_total(const TcShape& S, size_t extra) {
  return tc_smem_fixed(S, extra) + (size_t)S.nb1 * tc_w1_bytes(S) + (size_t)S.nbl * tc_wl_bytes(S);
}
// choose the buffer counts: everything resident if it fits, else the deepest rings that fit (Wl first: it is the larger
// image and the one whose reload window is shortest)
inline bool tc_plan_smem(TcShape& S, size_t extra, size_t& total) {
  const size_t cap = 227 * 1024;
  const int opts[4][2] = {{S.Lc, S.Lc}, {2, 2}, {1, 2}, {1, 1}};
  for (int i = 0; i < 4; ++i) {
    if (i == 0 && S.Lc > kTcMaxSlots) continue;
    S.nb1 = opts[i][0]; S.nbl = opts[i][1];
    if (i > 0 && S.Lc <= 2 && S.nb1 >= S.Lc && S.nbl >= S.Lc) continue;   // covered by the resident option
    total = tc_smem_total(S, extra);
    if (total <= cap) return true;
  }
  return false;
}

__device__ inline TcSmem tc_carve(unsigned char* smem, const TcShape& S) {
  TcSmem m;
  unsigned char* p = smem;
  m.a1[0] = p; p += tc_a1_bytes(S);
  m.a1[1] = p; p += tc_a1_bytes(S);
  m.w1 = p; p += (size_t)S.nb1 * tc_w1_bytes(S);
  m.wl = p; p += (size_t)S.nbl * tc_wl_bytes(S);
  m.bl = reinterpret_cast<float*>(p); p += (size_t)S.Lc * S.N2p * 4;
  m.aff = reinterpret_cast<float*>(p); p += (((size_t)(S.Lc + 1) * 4 * S.d + 4 + 3) & ~size_t(3)) * 4;
  m.red = reinterpret_cast<float*>(p); p += (size_t)2 * 2 * kTcGroups * kTcRows * 4;
  m.extra = p;
  return m;
}
__device__ inline void tc_carve_tail(TcSmem& m, size_t extra) {
  unsigned char* p = m.extra + ((extra + 15) & ~size_t(15));
  m.bars = reinterpret_cast<uint64_t*>(p);
  m.tmem_slot = reinterpret_cast<uint32_t*>(p + (size_t)kTcNumBars * 8);
}

// ---- optional timeline trace (build with -DNFMC_TC_TRACE; tools/tc_trace.py) ---------------------------------------
#ifdef NFMC_TC_TRACE
__device__ long long* g_tc_trace = nullptr;     // [2][2048] {event id, clock}: row 0 control lane, row 1 epilogue thread 0
struct TcTrace {                                // lives in registers of the tracing thread: no loads on the traced path
  long long* p;
  int n;
  __device__ void init(int who) { p = (blockIdx.x == 0 && g_tc_trace) ? g_tc_trace + who * 4096 : nullptr; n = 0; }
  __device__ __forceinline__ void ev(int id) {
    if (p && n < 2048) { p[2 * n] = id; p[2 * n + 1] = clock64(); ++n; }
  }
};
#define TC_TRACE_CTL(id) tr.ev(id)
#define TC_TRACE_EPI(id) do { if (threadIdx.x == 0) sy.tr.ev(id); } while (0)
#else
#define TC_TRACE_CTL(id)
#define TC_TRACE_EPI(id)
#endif

// ---- PTX wrappers --------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tma_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// shared-memory matrix descriptor: SWIZZLE_NONE, K-major.  An [R rows x K] bf16 operand is stored as [K/8][R][8]:
// 8 consecutive k of one row are 16 contiguous bytes, a "core matrix" = 8 rows x 16 B = 128 contiguous bytes, so
// SBO (next 8-row group) = 128 B and LBO (next 8-column group) = R * 16 B -- or any other distance: the second k-group
// of a K = 16 instruction may live anywhere (used by GEMM 2, see tc_flow.cu).
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
// D[tmem] (+)= A[smem] . B[smem]
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]: the A operand (128 lanes x 8 columns of packed bf16 pairs per K = 16) is read from
// tensor memory
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 8 / 16 consecutive 32-bit columns of this thread's TMEM lane (issue only; pair with tmem_wait_ld)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// the loaded registers are operands of the wait so that no use of them can be scheduled above it
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&a)[8], uint32_t (&b)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(b[0]),
                 "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&a)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&a)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]),
                 "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// tanh of two values at once, straight in the bf16 the next GEMM consumes: one MUFU op per pair.  a -> low half.
__device__ __forceinline__ uint32_t tanh_bf16x2(float a, float b) {
  uint32_t p, y;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(b), "f"(a));
  asm("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(p));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  uint32_t p;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(b), "f"(a));
  return p;
}
__device__ __forceinline__ float fast_ex2(float v) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float fast_rcp(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float lg2_any(float v) {   // lg2 that also handles denormals / inf
  float r;
  asm("lg2.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// ---- prologue / epilogue of the kernel (all kTcThreads threads) -----------------------------------------------------
__device__ inline uint32_t tc_prologue(TcSmem& sm, const unsigned char* blob, const TcShape& S, size_t extra = 0) {
  tc_carve_tail(sm, extra);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int t = 0; t < 2; ++t) {
      mbar_init(smem_u32(sm.bars + kTcBarG1 + t), 1);
      mbar_init(smem_u32(sm.bars + kTcBarG2 + t), 1);
      mbar_init(smem_u32(sm.bars + kTcBarA1 + t), kTcEpiThreads);
      mbar_init(smem_u32(sm.bars + kTcBarHid + t), kTcEpiThreads);
    }
    for (int s = 0; s < kTcMaxSlots; ++s) {
      mbar_init(smem_u32(sm.bars + kTcBarW1 + s), 1);
      mbar_init(smem_u32(sm.bars + kTcBarWl + s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(sm.tmem_slot), 512);
  // elementwise-affine tables and the second-layer biases of every coupling: resident for the whole kernel
  const int n_aff = (S.Lc + 1) * 4 * S.d + 4;
  for (int i = tid; i < n_aff; i += kTcThreads) sm.aff[i] = __ldg(reinterpret_cast<const float*>(blob) + i);
  const unsigned char* wblob = blob + tc_affine_bytes(S.d, S.Lc);
  const size_t cb = tc_coupling_bytes(S), bl_off = tc_w1_bytes(S) + tc_wl_bytes(S);
  for (int i = tid; i < S.Lc * S.N2p; i += kTcThreads) {
    const int l = i / S.N2p, k = i % S.N2p;
    sm.bl[i] = __ldg(reinterpret_cast<const float*>(wblob + (size_t)l * cb + bl_off) + k);
  }
  // A1 images: zero, plus the two constant-one columns (k = d/2, d/2 + 1: the bias rows of W1) where they fall into
  // k-groups no epilogue thread owns (k >= 64)
  const int da = S.d / 2;
  const int a1_words = (int)(tc_a1_bytes(S) / 4);
  for (int t = 0; t < 2; ++t) {
    uint32_t* w = reinterpret_cast<uint32_t*>(sm.a1[t]);
    for (int i = tid; i < a1_words; i += kTcThreads) {
      // word i: k-group kg = i / (128*4), row = (i / 4) % 128, pair = i % 4 -> k = 8 kg + 2 pair, +1
      const int kg = i / (kTcRows * 4), k0 = 8 * kg + 2 * (i & 3);
      uint32_t v = 0;
      if (kg >= 8) {
        if (k0 == da || k0 == da + 1) v |= 0x3F80u;              // bf16 1.0 in the low half
        if (k0 + 1 == da || k0 + 1 == da + 1) v |= 0x3F800000u;  // ... in the high half
      }
      w[i] = v;
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *sm.tmem_slot;
}
__device__ inline void tc_epilogue_dealloc(uint32_t tmem_base) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) tmem_dealloc(tmem_base, 512);
}

// ---- control warp (one lane) ------------------------------------------------------------------------------------------
// Sequence of couplings ("uses") this CTA will execute, in order; both sides (control / epilogue) count uses to derive
// barrier parities.  seq: 0 = forward passes only (layers 0..Lc-1 per pass), 1 = inverse only (Lc-1..0),
// 2 = forward pass then inverse pass, alternating.
struct TcControl {
  const TcSmem& sm;
  const TcShape& S;
  const unsigned char* wblob;
  uint32_t tmem_base, total_uses, use, seq;
  uint32_t idesc1, idesc2;
  size_t cb;
#ifdef NFMC_TC_TRACE
  TcTrace tr;
#endif

  __device__ TcControl(const TcSmem& sm_, const TcShape& S_, const unsigned char* blob, uint32_t tmem, uint32_t total, uint32_t seq_)
      : sm(sm_), S(S_), wblob(blob + tc_affine_bytes(S_.d, S_.Lc)), tmem_base(tmem), total_uses(total), use(0), seq(seq_) {
    idesc1 = umma_idesc(kTcRows, S.Hp);
    idesc2 = umma_idesc(kTcRows, S.N2p);
    cb = tc_coupling_bytes(S);
#ifdef NFMC_TC_TRACE
    tr.init(0);
#endif
  }
  __device__ int layer_of(uint32_t u) const {
    const uint32_t Lc = (uint32_t)S.Lc;
    if (seq == 0) return (int)(u % Lc);
    if (seq == 1) return (int)(Lc - 1 - u % Lc);
    const uint32_t i = u % (2 * Lc);
    return (int)(i < Lc ? i : 2 * Lc - 1 - i);
  }
  __device__ uint32_t bar(int which) const { return smem_u32(sm.bars + which); }
  __device__ void load_w1(int layer, int slot) {
    const uint32_t b = bar(kTcBarW1 + slot), bytes = (uint32_t)tc_w1_bytes(S);
    mbar_expect_tx(b, bytes);
    tma_bulk_load(smem_u32(sm.w1 + (size_t)slot * bytes), wblob + (size_t)layer * cb, bytes, b);
  }
  __device__ void load_wl(int layer, int slot) {
    const uint32_t b = bar(kTcBarWl + slot), bytes = (uint32_t)tc_wl_bytes(S);
    mbar_expect_tx(b, bytes);
    tma_bulk_load(smem_u32(sm.wl + (size_t)slot * bytes), wblob + (size_t)layer * cb + tc_w1_bytes(S), bytes, b);
  }
  __device__ void prime() {
    if (S.nb1 == S.Lc) { for (int l = 0; l < S.Lc; ++l) load_w1(l, l); }
    else for (uint32_t u = 0; u < (uint32_t)S.nb1 && u < total_uses; ++u) load_w1(layer_of(u), (int)u);
    if (S.nbl == S.Lc) { for (int l = 0; l < S.Lc; ++l) load_wl(l, l); }
    else for (uint32_t u = 0; u < (uint32_t)S.nbl && u < total_uses; ++u) load_wl(layer_of(u), (int)u);
  }
  __device__ void gemm1(int t, uint32_t w1_addr) {
    const uint32_t a0 = smem_u32(sm.a1[t]);
    const int ks = S.K1 / 16;
    for (int kk = 0; kk < ks; ++kk)
      umma_ss(tmem_base + t * kTcRegion, umma_desc(a0 + kk * 2 * (kTcRows * 16), kTcRows * 16, 128),
              umma_desc(w1_addr + kk * 2 * (S.Hp * 16), S.Hp * 16, 128), idesc1, kk > 0);
    umma_commit(bar(kTcBarG1 + t));
  }
  __device__ void gemm2(int t, uint32_t wl_addr) {
    const int ks = S.Hp / 16;
    const uint32_t lbo = (uint32_t)ks * S.N2p * 16;
    for (int s = 0; s < ks; ++s)
      umma_ts(tmem_base + t * kTcRegion + kTcUCol, tmem_base + t * kTcRegion + 8 * s,
              umma_desc(wl_addr + s * (S.N2p * 16), lbo, 128), idesc2, s > 0);
    umma_commit(bar(kTcBarG2 + t));
  }
  // one coupling (the next one of the sequence) for both tiles
  __device__ void coupling() {
    const uint32_t u = use, par = u & 1;
    const bool res1 = S.nb1 == S.Lc, resl = S.nbl == S.Lc;
    const int l = layer_of(u);
    const int s1 = res1 ? l : (int)(u % (uint32_t)S.nb1), sl = resl ? l : (int)(u % (uint32_t)S.nbl);
    const uint32_t w1_addr = smem_u32(sm.w1 + (size_t)s1 * tc_w1_bytes(S)), wl_addr = smem_u32(sm.wl + (size_t)sl * tc_wl_bytes(S));
    TC_TRACE_CTL(0);
    mbar_wait(bar(kTcBarW1 + s1), res1 ? 0u : ((u / (uint32_t)S.nb1) & 1));
    TC_TRACE_CTL(1);
    mbar_wait(bar(kTcBarA1 + 0), par);
    TC_TRACE_CTL(2);
    tc_fence_after();
    gemm1(0, w1_addr);
    TC_TRACE_CTL(3);
    mbar_wait(bar(kTcBarA1 + 1), par);
    TC_TRACE_CTL(4);
    tc_fence_after();
    gemm1(1, w1_addr);
    TC_TRACE_CTL(5);
    mbar_wait(bar(kTcBarWl + sl), resl ? 0u : ((u / (uint32_t)S.nbl) & 1));
    TC_TRACE_CTL(6);
    mbar_wait(bar(kTcBarHid + 0), par);
    TC_TRACE_CTL(7);
    tc_fence_after();
    gemm2(0, wl_addr);
    TC_TRACE_CTL(8);
    if (!res1 && u + S.nb1 < total_uses) {          // GEMM 1 of both tiles is done with this W1 buffer: refill it
      mbar_wait(bar(kTcBarG1 + 1), par);
      load_w1(layer_of(u + S.nb1), s1);
    }
    TC_TRACE_CTL(9);
    mbar_wait(bar(kTcBarHid + 1), par);
    TC_TRACE_CTL(10);
    tc_fence_after();
    gemm2(1, wl_addr);
    TC_TRACE_CTL(11);
    if (!resl && u + S.nbl < total_uses) {          // the tensor pipe is in order: nothing is lost by waiting here
      mbar_wait(bar(kTcBarG2 + 1), par);
      load_wl(layer_of(u + S.nbl), sl);
    }
    TC_TRACE_CTL(12);
    use = u + 1;
  }
};

// ---- epilogue warps ---------------------------------------------------------------------------------------------------
struct TcEpiSync {
  uint32_t g1[2], g2[2], a1[2], hid[2];
  uint32_t use;
#ifdef NFMC_TC_TRACE
  TcTrace tr;
#endif
  __device__ explicit TcEpiSync(const TcSmem& sm) : use(0) {
#ifdef NFMC_TC_TRACE
    tr.init(1);
#endif
    for (int t = 0; t < 2; ++t) {
      g1[t] = smem_u32(sm.bars + kTcBarG1 + t);
      g2[t] = smem_u32(sm.bars + kTcBarG2 + t);
      a1[t] = smem_u32(sm.bars + kTcBarA1 + t);
      hid[t] = smem_u32(sm.bars + kTcBarHid + t);
    }
  }
};

// chain row <-> registers.  Invalid slots (k >= d/2) hold the constant-one columns of GEMM 1 at k = d/2, d/2 + 1 and
// zero elsewhere; nothing below ever changes them (padded weight rows are zero, padded biases give alpha = 1).
__device__ __forceinline__ void tc_load_state(const float* __restrict__ src, int d, int da, int e0, bool fl, float (&lo)[kTcOwn],
                                              float (&hi)[kTcOwn]) {
  if (!fl && (d & 3) == 0 && (da & 1) == 0 && e0 + kTcOwn <= da) {
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
      const float pad = (k == da || k == da + 1) ? 1.f : 0.f;
      lo[q] = (k < da) ? __ldg(src + (fl ? d - 1 - k : k)) : pad;
      hi[q] = (k < da) ? __ldg(src + (fl ? d - 1 - (da + k) : da + k)) : pad;
    }
  }
}
__device__ __forceinline__ void tc_store_state(float* __restrict__ dst, int d, int da, int e0, bool fl, const float (&lo)[kTcOwn],
                                               const float (&hi)[kTcOwn]) {
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

// elementwise affine (forward {alpha, beta} or inverse {1/alpha, -beta/alpha}: the same fma)
__device__ __forceinline__ void tc_affine(const float* aff, int idx, bool inv, int d, int da, int e0, float (&lo)[kTcOwn], float (&hi)[kTcOwn]) {
  const float2* tab = reinterpret_cast<const float2*>(aff + idx * 4 * d + (inv ? 2 * d : 0));
#pragma unroll
  for (int q = 0; q < kTcOwn; ++q) {
    const int k = e0 + q;
    if (k < da) {
      const float2 pl = tab[k], ph = tab[da + k];
      lo[q] = fmaf(pl.x, lo[q], pl.y);
      hi[q] = fmaf(ph.x, hi[q], ph.y);
    }
  }
}

// A operand of GEMM 1: this thread's 16 source values as bf16 into k-groups 2g, 2g+1 of the [K1/8][128][8] image
__device__ __forceinline__ void tc_write_a1(unsigned char* a1, int r, int g, const float (&v)[kTcOwn]) {
#pragma unroll
  for (int h2 = 0; h2 < 2; ++h2) {
    uint4 w;
    const int q0 = h2 * 8;
    w.x = pack_bf16(v[q0 + 0], v[q0 + 1]);
    w.y = pack_bf16(v[q0 + 2], v[q0 + 3]);
    w.z = pack_bf16(v[q0 + 4], v[q0 + 5]);
    w.w = pack_bf16(v[q0 + 6], v[q0 + 7]);
    *reinterpret_cast<uint4*>(a1 + ((size_t)(2 * g + h2) * kTcRows + r) * 16) = w;
  }
}

// epilogue 1 of one tile: hid = tanh(Hpre) -> packed bf16 -> back into tensor memory as the A operand of GEMM 2.
// K-steps are dealt round-robin to the four column groups of a row.
__device__ __forceinline__ void tc_epi1(uint32_t trow, int Hp, int g) {
  const int nsteps = Hp >> 4;
  const uint32_t hi_off = (uint32_t)(Hp >> 1);
#pragma unroll 1
  for (int s = g; s < nsteps; s += kTcGroups) {
    uint32_t a[8], b[8], p[8];
    tmem_ld8(trow + 8 * s, a);
    tmem_ld8(trow + hi_off + 8 * s, b);
    tmem_wait_ld(a, b);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      p[i] = tanh_bf16x2(__uint_as_float(a[2 * i]), __uint_as_float(a[2 * i + 1]));
      p[4 + i] = tanh_bf16x2(__uint_as_float(b[2 * i]), __uint_as_float(b[2 * i + 1]));
    }
    tmem_st8(trow + 8 * s, p);
  }
  tmem_wait_st();
}

// epilogue 2 of one tile: (u_a', u_b') = U + bl' -> alpha = 2^u_a' + m, beta = u_b'; affine update of this thread's 16
// targets; log2-determinant accumulation (one lg2 per four scales; scales are >= m, so the product of four cannot
// underflow, and it cannot overflow while every scale is below 1e9 -- otherwise the slow branch takes them one by one).
template <bool INV>
__device__ __forceinline__ void tc_epi2(uint32_t tcol_u, const float* bl, int N2p, int g, float (&tgt)[kTcOwn], float& ld2) {
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) {
    const int c = 4 * g + c4;                       // 8 columns = 4 targets (u_a, u_b interleaved)
    if (c * 8 < N2p) {
      uint32_t v[8];
      tmem_ld8(tcol_u + c * 8, v);
      const float4* b4 = reinterpret_cast<const float4*>(bl + c * 8);
      const float4 bA = b4[0], bB = b4[1];
      tmem_wait_ld(v);
      const float a0 = fast_ex2(__uint_as_float(v[0]) + bA.x) + kMinScale, a1 = fast_ex2(__uint_as_float(v[2]) + bA.z) + kMinScale;
      const float a2 = fast_ex2(__uint_as_float(v[4]) + bB.x) + kMinScale, a3 = fast_ex2(__uint_as_float(v[6]) + bB.z) + kMinScale;
      const float ub0 = __uint_as_float(v[1]) + bA.y, ub1 = __uint_as_float(v[3]) + bA.w;
      const float ub2 = __uint_as_float(v[5]) + bB.y, ub3 = __uint_as_float(v[7]) + bB.w;
      const int q = c4 * 4;
      const float p01 = a0 * a1, p23 = a2 * a3;
      const bool tame = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)) < 1e9f;
      if (INV) {
        float r0, r1, r2, r3;
        if (tame) {
          const float r01 = fast_rcp(p01), r23 = fast_rcp(p23);
          r0 = r01 * a1; r1 = r01 * a0; r2 = r23 * a3; r3 = r23 * a2;
        } else {
          r0 = fast_rcp(a0); r1 = fast_rcp(a1); r2 = fast_rcp(a2); r3 = fast_rcp(a3);
        }
        tgt[q + 0] = (tgt[q + 0] - ub0) * r0;
        tgt[q + 1] = (tgt[q + 1] - ub1) * r1;
        tgt[q + 2] = (tgt[q + 2] - ub2) * r2;
        tgt[q + 3] = (tgt[q + 3] - ub3) * r3;
      } else {
        tgt[q + 0] = fmaf(a0, tgt[q + 0], ub0);
        tgt[q + 1] = fmaf(a1, tgt[q + 1], ub1);
        tgt[q + 2] = fmaf(a2, tgt[q + 2], ub2);
        tgt[q + 3] = fmaf(a3, tgt[q + 3], ub3);
      }
      if (tame) ld2 += fast_lg2(p01 * p23);
      else ld2 += (lg2_any(a0) + lg2_any(a1)) + (lg2_any(a2) + lg2_any(a3));
    }
  }
}

// one coupling for both tiles, epilogue side.  SRC: which half feeds the conditioner (0 = low, 1 = high); the other half
// is transformed.  aff_after: index of the elementwise affine that follows this coupling in pass order; write_next: the
// pass goes on with another coupling, whose source half (= this coupling's target half) goes out as the next A1 image.
template <bool INV, int SRC>
__device__ __forceinline__ void tc_coupling_epi(const TcSmem& sm, const TcShape& S, TcEpiSync& sy, uint32_t trow, int r, int g, int layer,
                                                int aff_after, bool write_next, float (&st)[2][2][kTcOwn], float (&ld2)[2]) {
  const uint32_t par = sy.use & 1;
  const int d = S.d, da = d / 2, e0 = g * kTcOwn;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    TC_TRACE_EPI(20 + t);
    mbar_wait(sy.g1[t], par);
    TC_TRACE_EPI(22 + t);
    tc_fence_after();
    tc_epi1(trow + t * kTcRegion, S.Hp, g);
    tc_fence_before();
    mbar_arrive(sy.hid[t]);
    TC_TRACE_EPI(24 + t);
  }
  const float* bl = sm.bl + (size_t)layer * S.N2p;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    mbar_wait(sy.g2[t], par);
    TC_TRACE_EPI(26 + t);
    tc_fence_after();
    tc_epi2<INV>(trow + t * kTcRegion + kTcUCol, bl, S.N2p, g, st[t][1 - SRC], ld2[t]);
    tc_affine(sm.aff, aff_after, INV, d, da, e0, st[t][0], st[t][1]);
    if (write_next) {
      tc_write_a1(sm.a1[t], r, g, st[t][1 - SRC]);
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(sy.a1[t]);
    }
    TC_TRACE_EPI(28 + t);
  }
  sy.use += 1;
}

// a whole pass (2 Lc + 1 layers) over both tiles.  st holds the input on entry and the output on return; ld2 receives
// the sum of log2(alpha) over the couplings (the caller adds the constant and flips the sign for the inverse).
template <bool INV>
__device__ __forceinline__ void tc_run_pass(const TcSmem& sm, const TcShape& S, TcEpiSync& sy, uint32_t trow, int r, int g,
                                            float (&st)[2][2][kTcOwn], float (&ld2)[2]) {
  const int Lc = S.Lc, d = S.d, da = d / 2, e0 = g * kTcOwn;
  const int l0 = INV ? Lc - 1 : 0;
  const bool src0_hi = (l0 & 1) == 0;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    tc_affine(sm.aff, INV ? Lc : 0, INV, d, da, e0, st[t][0], st[t][1]);
    if (src0_hi) tc_write_a1(sm.a1[t], r, g, st[t][1]);
    else tc_write_a1(sm.a1[t], r, g, st[t][0]);
    fence_async_smem();
    tc_fence_before();
    mbar_arrive(sy.a1[t]);
  }
#pragma unroll 1
  for (int i = 0; i < Lc; ++i) {
    const int l = INV ? Lc - 1 - i : i;
    const int aff_after = INV ? l : l + 1;
    const bool more = i + 1 < Lc;
    if ((l & 1) == 0) tc_coupling_epi<INV, 1>(sm, S, sy, trow, r, g, l, aff_after, more, st, ld2);
    else tc_coupling_epi<INV, 0>(sm, S, sy, trow, r, g, l, aff_after, more, st, ld2);
  }
}

inline int tc_validate(const nfmc_realnvp_tc* flow, TcShape& S, const char* who) {
  if (!flow || !flow->blob) return set_error(std::string(who) + ": flow is NULL");
  if (!tc_shape(flow->d, flow->n_coupling, flow->hidden, S))
    return set_error(std::string(who) + ": tensor-core path needs even d <= 128, hidden a multiple of 16 in [16, 256], n_coupling >= 1");
  const int64_t want = (int64_t)(tc_affine_bytes(S.d, S.Lc) + (size_t)S.Lc * tc_coupling_bytes(S));
  if (flow->blob_bytes != want) return set_error(std::string(who) + ": blob_bytes mismatch");
  return 0;
}

}  // namespace nfmc
