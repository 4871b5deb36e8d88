// tc_common.cuh -- building blocks of the tensor-core RealNVP kernels (tc_flow.cu, tc_jump.cu): PTX wrappers for
// tcgen05 / TMEM / TMA / mbarrier, the packed-flow shape, the shared-memory plan, the issue sequences of the service
// warps and the epilogue-warp side of a pipelined two-tile flow pass.  The design is described at the top of tc_flow.cu.
#pragma once
#include <cuda_bf16.h>
#include <cstdio>
#include "host_common.cuh"

namespace nfmc {

constexpr int kTcRows = 128;                 // chains per tile = UMMA_M
constexpr int kTcGroups = 4;                 // column groups per chain row
constexpr int kTcOwn = 16;                   // elements per half per thread
constexpr int kTcEpiWarps = 16;
constexpr int kTcEpiThreads = kTcEpiWarps * 32;
constexpr int kTcThreads = kTcEpiThreads + 128;  // + one warpgroup of service warps: MMA issuer, weight loader, tile loader
constexpr int kTcWarpMma = 16, kTcWarpWeights = 17, kTcWarpTiles = 18;
constexpr int kTcRegsEpi = 112, kTcRegsService = 32;   // setmaxnreg: 512*112 + 128*32 = 640*96, the launch allocation
constexpr int kTcRegion = 256;               // TMEM columns per tile
constexpr int kTcUCol = 128;                 // column of the U accumulator inside a tile's region
constexpr int kTcMaxSlots = 16;              // weight buffers per image kind (resident mode: one per coupling)
constexpr float kLog2e = 1.4426950408889634f;
#ifndef NFMC_TC_TANH_MUFU_HI
#define NFMC_TC_TANH_MUFU_HI 2
#endif
constexpr int kTcTanhMufuPairsHi = NFMC_TC_TANH_MUFU_HI;   // of a K-step's 8 pairs, 4 + this many go through MUFU.TANH, the rest through the FMA pipe

// Packed flow shape + shared-memory plan.
//   nb1 / nbl: number of shared-memory buffers for the W1 / Wl images; == Lc means every coupling's image stays resident
//              for the whole kernel, otherwise the buffers form a ring refilled by the weight-loader warp.
//   na1:       A1 image buffers (2 = one per tile, 1 = shared by the two tiles, with an extra wait on GEMM 1)
//   nx:        2 = chain tiles are staged through two tile buffers by TMA (loaded a whole pass ahead, stored
//              asynchronously); 0 = no room: tiles are transposed through tensor memory with ordinary loads / stores.
struct TcShape {
  int d, Lc, Hp, N2p, K1;
  int nb1, nbl, na1, nx;
};

__host__ __device__ __forceinline__ size_t tc_w1_bytes(const TcShape& S) { return (size_t)S.K1 * S.Hp * 2; }
__host__ __device__ __forceinline__ size_t tc_wl_bytes(const TcShape& S) { return (size_t)S.Hp * S.N2p * 2; }
__host__ __device__ __forceinline__ size_t tc_coupling_bytes(const TcShape& S) { return tc_w1_bytes(S) + tc_wl_bytes(S) + (size_t)S.N2p * 4; }
__host__ __device__ __forceinline__ size_t tc_affine_bytes(int d, int Lc) { return ((size_t)(Lc + 1) * 4 * d + 4) * 4; }
__host__ __device__ __forceinline__ size_t tc_a1_bytes(const TcShape& S) { return (size_t)(S.K1 < 64 ? 64 : S.K1) * kTcRows * 2; }
__host__ __device__ __forceinline__ size_t tc_tile_bytes(const TcShape& S) { return (size_t)kTcRows * S.d * 4; }
__host__ __device__ __forceinline__ size_t tc_aff4_bytes(const TcShape& S) { return (size_t)(S.Lc + 1) * 2 * 64 * 16; }

inline bool tc_shape(int d, int Lc, int hidden, TcShape& S) {
  if (d < 2 || d > 128 || (d & 1) || Lc < 1 || hidden < 16 || hidden > 256 || (hidden & 15)) return false;
  S.d = d; S.Lc = Lc; S.Hp = hidden;
  S.N2p = ((d - d / 2) * 2 + 15) & ~15;
  S.K1 = (d / 2 + 2 + 15) & ~15;
  S.nb1 = S.nbl = 1; S.na1 = 1; S.nx = 0;
  return true;
}

// shared-memory carve-up (in this order): A1 images (the row-reduction scratch [2 tiles][2][4][128] floats aliases the
// first one: it is used between passes only) | W1 slots | Wl slots | tile buffers | bl' [Lc][N2p] | affine tables float4
// [(Lc+1)][2][64] | extra (caller) | mbarriers | tmem slot
struct TcSmem {
  unsigned char* a1_base;     // A1 image of tile t: a1(t)
  uint32_t a1_stride;         // 0 when the two tiles share one image
  unsigned char* w1;
  unsigned char* wl;
  unsigned char* x_base;      // tile buffer t: x(t)
  uint32_t x_stride;
  __device__ __forceinline__ unsigned char* a1(int t) const { return a1_base + (size_t)t * a1_stride; }
  __device__ __forceinline__ unsigned char* x(int t) const { return x_base + (size_t)t * x_stride; }
  float* bl;
  float4* aff4;
  float* red;
  unsigned char* extra;
  uint64_t* bars;
  uint32_t* tmem_slot;
  float log_const;
};
constexpr int kTcBarG1 = 0, kTcBarG2 = 2, kTcBarA1 = 4, kTcBarHid = 6, kTcBarXFull = 8, kTcBarXReady = 10, kTcBarW1 = 12,
              kTcBarWl = 12 + kTcMaxSlots, kTcNumBars = 12 + 2 * kTcMaxSlots;

__host__ __device__ __forceinline__ size_t tc_smem_total(const TcShape& S, size_t extra) {
  return (size_t)S.na1 * tc_a1_bytes(S) + (size_t)S.nb1 * tc_w1_bytes(S) + (size_t)S.nbl * tc_wl_bytes(S) + (size_t)S.nx * tc_tile_bytes(S) +
         (size_t)S.Lc * S.N2p * 4 + tc_aff4_bytes(S) + ((extra + 15) & ~size_t(15)) + (size_t)kTcNumBars * 8 + 16;
}
// Choose the buffers greedily, in the order the measurements rank them (profiles/tc_r02*): TMA-staged tiles first
// (tile I/O was 45 % of the kernel with per-thread global accesses), then a second Wl buffer, a second A1 image, a second
// W1 buffer, and finally every coupling resident.  `staged_ok`: the caller's tensors allow TMA staging (16-byte aligned,
// d % 4 == 0).  NFMC_TC_PLAN="nb1,nbl,na1,nx" overrides (experiments; -1 = resident).
inline bool tc_plan_smem(TcShape& S, size_t extra, size_t& total, bool staged_ok) {
  const size_t cap = 227 * 1024;
  if (const char* e = getenv("NFMC_TC_PLAN")) {
    int a, b, c, dd;
    if (sscanf(e, "%d,%d,%d,%d", &a, &b, &c, &dd) == 4) {
      S.nb1 = a < 0 ? S.Lc : a; S.nbl = b < 0 ? S.Lc : b; S.na1 = c; S.nx = staged_ok ? dd : 0;
      total = tc_smem_total(S, extra);
      return total <= cap && S.nb1 >= 1 && S.nbl >= 1 && S.nb1 <= kTcMaxSlots && S.nbl <= kTcMaxSlots && (c == 1 || c == 2) && (S.nx == 0 || S.nx == 2);
    }
  }
  S.nb1 = S.nbl = 1; S.na1 = 1; S.nx = 0;
  total = tc_smem_total(S, extra);
  if (total > cap) return false;
  auto fits = [&](const TcShape& T) { return tc_smem_total(T, extra) <= cap; };
  TcShape T = S;
  if (staged_ok) { T = S; T.nx = 2; if (fits(T)) S = T; }
  if (S.Lc > 1) { T = S; T.nbl = 2; if (fits(T)) S = T; }
  T = S; T.na1 = 2; if (fits(T)) S = T;
  if (S.Lc > 1) { T = S; T.nb1 = 2; if (fits(T)) S = T; }
  if (S.Lc <= kTcMaxSlots) { T = S; T.nb1 = S.Lc; T.nbl = S.Lc; if (fits(T)) S = T; }
  if (S.nb1 > S.Lc) S.nb1 = S.Lc;
  if (S.nbl > S.Lc) S.nbl = S.Lc;
  total = tc_smem_total(S, extra);
  return true;
}

__device__ __forceinline__ TcSmem tc_carve(unsigned char* smem, const TcShape& S, size_t extra = 0) {
  TcSmem m;
  unsigned char* p = smem;
  m.a1_base = p;
  m.a1_stride = S.na1 == 2 ? (uint32_t)tc_a1_bytes(S) : 0u;
  m.red = reinterpret_cast<float*>(p);
  p += (size_t)S.na1 * tc_a1_bytes(S);
  m.w1 = p; p += (size_t)S.nb1 * tc_w1_bytes(S);
  m.wl = p; p += (size_t)S.nbl * tc_wl_bytes(S);
  m.x_base = p;
  m.x_stride = (uint32_t)tc_tile_bytes(S);
  p += (size_t)S.nx * tc_tile_bytes(S);
  m.bl = reinterpret_cast<float*>(p); p += (size_t)S.Lc * S.N2p * 4;
  m.aff4 = reinterpret_cast<float4*>(p); p += tc_aff4_bytes(S);
  m.extra = p; p += (extra + 15) & ~size_t(15);
  m.bars = reinterpret_cast<uint64_t*>(p);
  m.tmem_slot = reinterpret_cast<uint32_t*>(p + (size_t)kTcNumBars * 8);
  m.log_const = 0.f;
  return m;
}

// ---- optional timeline trace (build with -DNFMC_TC_TRACE; tools/tc_trace.py) ---------------------------------------
#ifdef NFMC_TC_TRACE
__device__ long long* g_tc_trace = nullptr;     // [2][2048] {event id, clock}: row 0 MMA lane, row 1 epilogue thread 0
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
// The suspend-time hint lets the hardware park a waiting warp (it wakes as soon as the phase completes) instead of having it
// spin through try_wait / branch / yield: in the r02c profile of the flow kernel those three were 27 % of all issued
// instructions, most of them from the service warps that share a scheduler with four epilogue warps each.
constexpr uint32_t kMbarSuspendNs = 200000u;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(kMbarSuspendNs)
        : "memory");
  }
}
// the same with a back-off between polls, for waits that are expected to be long (tile boundaries, the loader lanes): the
// warp scheduler serves the highest warp id first, so a polling warp takes issue slots from lower-numbered warps that still
// have work
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(kMbarSuspendNs)
        : "memory");
    if (done) break;
    __nanosleep(40);
  }
}
__device__ __forceinline__ void tma_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_bulk_store(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {   // at most N of this thread's bulk stores still reading smem
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool elect_one() {   // one lane of a converged warp
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory"); }
template <int R>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }

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
// of a K = 16 instruction may live anywhere (used by GEMM 2, see tc_flow.cu).  The start address sits in the low 14
// bits (units of 16 B), so stepping an operand through K is an integer add on the low word.
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
// D[tmem] (+)= A[smem] . B[smem].  Descriptors travel as (low, high) words: stepping through K only touches the low one.
template <bool ACC>
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "n"(ACC ? 1 : 0)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]: the A operand (128 lanes x 8 columns of packed bf16 pairs per K = 16) is read from
// tensor memory
template <bool ACC>
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "n"(ACC ? 1 : 0)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 8 consecutive 32-bit columns of this thread's TMEM lane (issue only; pair with tmem_wait_ld)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// fp32 pair -> packed bf16 pair with INTEGER instructions (add half an ulp to the magnitude, keep the high halves: round
// to nearest, ties away from zero).  cvt.rn.bf16x2.f32 would run on the XU pipe -- the same pipe as tanh / ex2 / lg2, which
// is the busiest unit of this kernel (ncu: profiles/tc_r02a) -- whereas IADD / PRMT go to the ALU.  a -> low half.
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  return __byte_perm(__float_as_uint(a) + 0x8000u, __float_as_uint(b) + 0x8000u, 0x7632);
}
// tanh of two values, packed as the bf16 pair the next GEMM consumes.  tanh.approx.bf16x2 looks like one operation per
// pair but compiles to pack (2 VIADD + PRMT), MUFU.TANH.BF16 on each half and a PRMT to re-pack (SASS of the r02c build):
// six instructions.  Taking tanh.approx.f32 of the unrounded values and packing the results costs five, and the
// activation sees the fp32 pre-activation instead of its bf16 rounding.
__device__ __forceinline__ float tanh_f32(float v) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(v));
  return y;
}
__device__ __forceinline__ uint32_t tanh_bf16x2(float a, float b) { return pack_bf16(tanh_f32(a), tanh_f32(b)); }
// The same pair on the FMA pipe: tanh(x) ~ xc * P(xc^2), xc = clamp(x, +-3.75), P of degree 8 (minimax in RELATIVE error:
// 9.5e-4 everywhere, half of bf16's rounding step), evaluated for both values at once with packed FFMA2.  Epilogue 1 is
// bound by the MUFU unit (16 results per clock and SM: 2 048 cycles for the 128 x 256 tanh of a tile at H = 256), so
// some pairs take this route and the two pipes work side by side.  Measured (d = 100, Lc = 4, H = 256, forward pass, ms): all
// MUFU 0.639, 2 of 8 pairs here 0.631 (log_prob 0.612 -> 0.592), 3 of 8 0.670, 4 of 8 0.719 -- the route costs 17 issue slots
// per pair against 5, and issue slots are the scarcer resource; 2 of 8 is the default.
__device__ __forceinline__ uint32_t tanh_poly_bf16x2(float a, float b) {
  const float c = 3.75f;
  const float2 x = make_float2(fminf(fmaxf(a, -c), c), fminf(fmaxf(b, -c), c));
  const float2 t = mul2(x, x);
  float2 p = splat2(1.816051865e-08f);
  p = fma2(p, t, splat2(-1.175949670e-06f));
  p = fma2(p, t, splat2(3.218734409e-05f));
  p = fma2(p, t, splat2(-4.870880060e-04f));
  p = fma2(p, t, splat2(4.496934319e-03f));
  p = fma2(p, t, splat2(-2.666925219e-02f));
  p = fma2(p, t, splat2(1.070194904e-01f));
  p = fma2(p, t, splat2(-3.221402460e-01f));
  p = fma2(p, t, splat2(9.991282172e-01f));
  const float2 y = mul2(x, p);
  return pack_bf16(y.x, y.y);
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
__device__ __forceinline__ uint32_t tc_prologue(TcSmem& sm, const unsigned char* blob, const TcShape& S) {
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int t = 0; t < 2; ++t) {
      mbar_init(smem_u32(sm.bars + kTcBarG1 + t), 1);
      mbar_init(smem_u32(sm.bars + kTcBarG2 + t), 1);
      mbar_init(smem_u32(sm.bars + kTcBarA1 + t), kTcEpiThreads);
      mbar_init(smem_u32(sm.bars + kTcBarHid + t), kTcEpiThreads);
      mbar_init(smem_u32(sm.bars + kTcBarXFull + t), 1);
      mbar_init(smem_u32(sm.bars + kTcBarXReady + t), kTcEpiThreads);
    }
    for (int s = 0; s < kTcMaxSlots; ++s) {
      mbar_init(smem_u32(sm.bars + kTcBarW1 + s), 1);
      mbar_init(smem_u32(sm.bars + kTcBarWl + s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(sm.tmem_slot), 512);
  // elementwise-affine tables as {alpha_lo, beta_lo, alpha_hi, beta_hi}[64] per (layer, direction), identity in the pad
  // slots, and the second-layer biases of every coupling: resident for the whole kernel
  const int d = S.d, da = d / 2;
  const float* gaff = reinterpret_cast<const float*>(blob);
  for (int i = tid; i < (S.Lc + 1) * 2 * 64; i += kTcThreads) {
    const int k = i & 63, dir = (i >> 6) & 1, idx = i >> 7;
    float4 v = make_float4(1.f, 0.f, 1.f, 0.f);
    if (k < da) {
      const float* tab = gaff + idx * 4 * d + dir * 2 * d;
      v = make_float4(__ldg(tab + 2 * k), __ldg(tab + 2 * k + 1), __ldg(tab + 2 * (da + k)), __ldg(tab + 2 * (da + k) + 1));
    }
    sm.aff4[i] = v;
  }
  sm.log_const = __ldg(gaff + (S.Lc + 1) * 4 * d);
  const unsigned char* wblob = blob + tc_affine_bytes(S.d, S.Lc);
  const size_t cb = tc_coupling_bytes(S), bl_off = tc_w1_bytes(S) + tc_wl_bytes(S);
  for (int i = tid; i < S.Lc * S.N2p; i += kTcThreads) {
    const int l = i / S.N2p, k = i % S.N2p;
    sm.bl[i] = __ldg(reinterpret_cast<const float*>(wblob + (size_t)l * cb + bl_off) + k);
  }
  // A1 images: zero, plus the two constant-one columns (k = d/2, d/2 + 1: the bias rows of W1) where they fall into
  // k-groups no epilogue thread owns (k >= 64)
  const int a1_words = (int)(tc_a1_bytes(S) / 4);
  for (int t = 0; t < S.na1; ++t) {
    uint32_t* w = reinterpret_cast<uint32_t*>(sm.a1(t));
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
  // The CTA owns all 512 columns of its SM's tensor memory (one CTA per SM: register file and TMEM both full), so the
  // allocation starts at column 0, lane 0.  Every TMEM address below is therefore a compile-time constant plus a
  // warp-uniform offset, which keeps the MMA issue loop on the uniform datapath.
  if (*sm.tmem_slot != 0u) __trap();
  return 0u;
}
__device__ __forceinline__ void tc_epilogue_dealloc(uint32_t tmem_base) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) tmem_dealloc(tmem_base, 512);
}

// ---- the sequence of couplings ("uses") a CTA executes --------------------------------------------------------------
// Every role counts uses to derive barrier parities.  seq: 0 = forward passes only (layers 0..Lc-1 per pass),
// 1 = inverse only (Lc-1..0), 2 = forward pass then inverse pass, alternating.
__device__ __forceinline__ int tc_layer_of(uint32_t u, uint32_t seq, uint32_t Lc) {
  if (seq == 0) return (int)(u % Lc);
  if (seq == 1) return (int)(Lc - 1 - u % Lc);
  const uint32_t i = u % (2 * Lc);
  return (int)(i < Lc ? i : 2 * Lc - 1 - i);
}

// ---- weight-loader warp: fills the W1 / Wl buffers by TMA ------------------------------------------------------------
// Executed by the whole warp in uniform control flow; `lead` (one elected lane) issues.
__device__ __forceinline__ void tc_weight_loader(const TcSmem& sm, const TcShape& S, const unsigned char* blob, uint32_t total_uses, uint32_t seq,
                                        bool lead) {
  const unsigned char* wblob = blob + tc_affine_bytes(S.d, S.Lc);
  const size_t cb = tc_coupling_bytes(S);
  const uint32_t b1 = (uint32_t)tc_w1_bytes(S), b2 = (uint32_t)tc_wl_bytes(S);
  auto load_w1 = [&](int layer, int slot) {
    if (lead) {
      const uint32_t bar = smem_u32(sm.bars + kTcBarW1 + slot);
      mbar_expect_tx(bar, b1);
      tma_bulk_load(smem_u32(sm.w1 + (size_t)slot * b1), wblob + (size_t)layer * cb, b1, bar);
    }
  };
  auto load_wl = [&](int layer, int slot) {
    if (lead) {
      const uint32_t bar = smem_u32(sm.bars + kTcBarWl + slot);
      mbar_expect_tx(bar, b2);
      tma_bulk_load(smem_u32(sm.wl + (size_t)slot * b2), wblob + (size_t)layer * cb + b1, b2, bar);
    }
  };
  const bool res1 = S.nb1 == S.Lc, resl = S.nbl == S.Lc;
  if (total_uses == 0) return;
  if (res1) { for (int l = 0; l < S.Lc; ++l) load_w1(l, l); }
  else for (uint32_t u = 0; u < (uint32_t)S.nb1 && u < total_uses; ++u) load_w1(tc_layer_of(u, seq, S.Lc), (int)u);
  if (resl) { for (int l = 0; l < S.Lc; ++l) load_wl(l, l); }
  else for (uint32_t u = 0; u < (uint32_t)S.nbl && u < total_uses; ++u) load_wl(tc_layer_of(u, seq, S.Lc), (int)u);
  if (res1 && resl) return;
  const uint32_t g1b = smem_u32(sm.bars + kTcBarG1 + 1), g2b = smem_u32(sm.bars + kTcBarG2 + 1);
  for (uint32_t u = 0; u < total_uses; ++u) {
    // a buffer is free once GEMM 1 (W1) / GEMM 2 (Wl) of the SECOND tile of use u has completed
    if (!res1 && u + S.nb1 < total_uses) {
      mbar_wait(g1b, u & 1);
      load_w1(tc_layer_of(u + S.nb1, seq, S.Lc), (int)(u % (uint32_t)S.nb1));
    }
    if (!resl && u + S.nbl < total_uses) {
      mbar_wait(g2b, u & 1);
      load_wl(tc_layer_of(u + S.nbl, seq, S.Lc), (int)(u % (uint32_t)S.nbl));
    }
  }
}

// ---- tile-prefetch lane: pulls the next pair's rows towards L2 while the current pair computes -------------------------
__device__ __forceinline__ void tc_tile_prefetch(const float* base, long long n, int d, long long my_pairs) {
  for (long long p = 1; p < my_pairs; ++p) {
    const long long row0 = ((long long)blockIdx.x + p * gridDim.x) * 2 * kTcRows;
    long long rows = n - row0;
    if (rows > 2 * kTcRows) rows = 2 * kTcRows;
    if (rows <= 0) break;
    const float* src = base + row0 * d;
    const size_t bytes = ((size_t)rows * d * sizeof(float)) & ~size_t(15);
    if (bytes >= 16 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) tma_prefetch_l2(src, (uint32_t)bytes);
  }
}

// ---- tile-loader lane (plan nx == 2): all tile traffic of the CTA as TMA bulk copies ----------------------------------
// Tile buffer t holds, in turn: the input of pair k (loaded here a whole pass ahead) -> at "boundary k" (start of pass k)
// every epilogue thread reads its pieces of that input and, for k > 0, overwrites them with its pieces of the OUTPUT of
// pair k - 1 (same locations, so no barrier is needed between the two) -> xready[t] completes -> this lane stores the
// buffer to global memory (asynchronously: the store drains while pass k computes), waits until the store has finished
// READING the buffer, and loads the input of pair k + 1 into it.  Boundary my_pairs is the last one: output only.
__device__ __forceinline__ void tc_x_loader(const TcSmem& sm, const TcShape& S, const float* in, float* out, long long n, long long my_pairs) {
  const size_t tb = tc_tile_bytes(S);
  auto tile_of = [&](long long k, int t) { return ((long long)blockIdx.x + k * gridDim.x) * 2 + t; };
  auto rows_of = [&](long long tile) { long long r = n - tile * kTcRows; return r > kTcRows ? (long long)kTcRows : (r < 0 ? 0ll : r); };
  auto load = [&](long long k, int t) {
    const long long tile = tile_of(k, t), rows = rows_of(tile);
    const uint32_t bar = smem_u32(sm.bars + kTcBarXFull + t);
    if (rows > 0) {
      const uint32_t bytes = (uint32_t)(rows * S.d * 4);
      mbar_expect_tx(bar, bytes);
      tma_bulk_load(smem_u32(sm.x(t)), in + tile * kTcRows * (long long)S.d, bytes, bar);
    } else {
      mbar_arrive(bar);
    }
  };
  if (my_pairs == 0) return;
  load(0, 0);
  load(0, 1);
  for (long long k = 0; k <= my_pairs; ++k) {
    for (int t = 0; t < 2; ++t) {
      mbar_wait(smem_u32(sm.bars + kTcBarXReady + t), (uint32_t)(k & 1));
      if (k > 0 && out) {
        const long long tile = tile_of(k - 1, t), rows = rows_of(tile);
        if (rows > 0) tma_bulk_store(out + tile * kTcRows * (long long)S.d, smem_u32(sm.x(t)), (uint32_t)(rows * S.d * 4));
        else asm volatile("cp.async.bulk.commit_group;" ::: "memory");     // keep one group per buffer and boundary
      }
    }
    if (k + 1 < my_pairs) {
      if (k > 0 && out) tma_store_wait_read<1>();    // buffer 0's store has been read out
      load(k + 1, 0);
      if (k > 0 && out) tma_store_wait_read<0>();
      load(k + 1, 1);
    }
  }
  tma_store_wait_all();      // the stores must have landed before the CTA's shared memory goes away
}

// ---- MMA-issuer warp -------------------------------------------------------------------------------------------------
// The whole warp runs the sequence in uniform control flow (so descriptors and tensor-memory addresses stay in uniform
// registers and stepping an operand through K is one uniform add); the elected lane issues tcgen05.mma / commit.
struct TcMma {
  const TcSmem& sm;
  const TcShape& S;
  uint32_t use, seq;
  uint32_t idesc1, idesc2;
  bool lead;
#ifdef NFMC_TC_TRACE
  TcTrace tr;
#endif
  __device__ TcMma(const TcSmem& sm_, const TcShape& S_, uint32_t seq_, bool lead_) : sm(sm_), S(S_), use(0), seq(seq_), lead(lead_) {
    idesc1 = umma_idesc(kTcRows, S.Hp);
    idesc2 = umma_idesc(kTcRows, S.N2p);
#ifdef NFMC_TC_TRACE
    tr.init(0);
    if (!lead) tr.p = nullptr;
#endif
  }
  __device__ uint32_t bar(int which) const { return smem_u32(sm.bars + which); }
  template <int T>
  __device__ __forceinline__ void gemm1(uint64_t bdesc) {
    constexpr uint32_t dcol = T * kTcRegion;
    const uint64_t ad = umma_desc(smem_u32(sm.a1(T)), kTcRows * 16, 128);
    uint32_t a_lo = (uint32_t)ad, b_lo = (uint32_t)bdesc;
    const uint32_t a_hi = (uint32_t)(ad >> 32), b_hi = (uint32_t)(bdesc >> 32);
    const uint32_t astep = (2 * (kTcRows * 16)) >> 4, bstep = (uint32_t)(2 * (S.Hp * 16)) >> 4;
    const int ks = S.K1 / 16;
    if (lead) umma_ss<false>(dcol, a_lo, a_hi, b_lo, b_hi, idesc1);
#pragma unroll 4
    for (int kk = 1; kk < ks; ++kk) {
      a_lo += astep; b_lo += bstep;
      if (lead) umma_ss<true>(dcol, a_lo, a_hi, b_lo, b_hi, idesc1);
    }
    if (lead) umma_commit(bar(kTcBarG1 + T));
  }
  template <int T>
  __device__ __forceinline__ void gemm2(uint64_t bdesc) {
    constexpr uint32_t dcol = T * kTcRegion + kTcUCol;
    uint32_t acol = T * kTcRegion, b_lo = (uint32_t)bdesc;
    const uint32_t b_hi = (uint32_t)(bdesc >> 32);
    const uint32_t bstep = (uint32_t)(S.N2p * 16) >> 4;
    const int ks = S.Hp / 16;
    if (lead) umma_ts<false>(dcol, acol, b_lo, b_hi, idesc2);
#pragma unroll 4
    for (int s = 1; s < ks; ++s) {
      acol += 8; b_lo += bstep;
      if (lead) umma_ts<true>(dcol, acol, b_lo, b_hi, idesc2);
    }
    if (lead) umma_commit(bar(kTcBarG2 + T));
  }
  // one coupling (the next one of the sequence) for both tiles
  __device__ void coupling() {
    const uint32_t u = use, par = u & 1;
    const bool res1 = S.nb1 == S.Lc, resl = S.nbl == S.Lc;
    const int l = tc_layer_of(u, seq, S.Lc);
    const int s1 = res1 ? l : (int)(u % (uint32_t)S.nb1), sl = resl ? l : (int)(u % (uint32_t)S.nbl);
    const uint64_t w1_desc = umma_desc(smem_u32(sm.w1 + (size_t)s1 * tc_w1_bytes(S)), S.Hp * 16, 128);
    const uint64_t wl_desc = umma_desc(smem_u32(sm.wl + (size_t)sl * tc_wl_bytes(S)), (uint32_t)(S.Hp / 16) * S.N2p * 16, 128);
    TC_TRACE_CTL(0);
    mbar_wait(bar(kTcBarW1 + s1), res1 ? 0u : ((u / (uint32_t)S.nb1) & 1));
    TC_TRACE_CTL(1);
    mbar_wait(bar(kTcBarA1 + 0), par);
    TC_TRACE_CTL(2);
    tc_fence_after();
    gemm1<0>(w1_desc);
    TC_TRACE_CTL(3);
    mbar_wait(bar(kTcBarA1 + 1), par);
    TC_TRACE_CTL(4);
    tc_fence_after();
    gemm1<1>(w1_desc);
    TC_TRACE_CTL(5);
    mbar_wait(bar(kTcBarWl + sl), resl ? 0u : ((u / (uint32_t)S.nbl) & 1));
    TC_TRACE_CTL(6);
    mbar_wait(bar(kTcBarHid + 0), par);
    TC_TRACE_CTL(7);
    tc_fence_after();
    gemm2<0>(wl_desc);
    TC_TRACE_CTL(8);
    mbar_wait(bar(kTcBarHid + 1), par);
    TC_TRACE_CTL(10);
    tc_fence_after();
    gemm2<1>(wl_desc);
    TC_TRACE_CTL(11);
    use = u + 1;
  }
};

// ---- epilogue warps ---------------------------------------------------------------------------------------------------
struct TcEpiSync {
  uint32_t bars;       // shared-space address of the barrier array
  uint32_t use;
#ifdef NFMC_TC_TRACE
  TcTrace tr;
#endif
  __device__ explicit TcEpiSync(const TcSmem& sm) : bars(smem_u32(sm.bars)), use(0) {
#ifdef NFMC_TC_TRACE
    tr.init(1);
#endif
  }
  __device__ __forceinline__ uint32_t bar(int which) const { return bars + which * 8; }
};

// slot k of a half that is not a chain coordinate: the constant-one columns of GEMM 1 sit at k = d/2 and d/2 + 1,
// everything else is zero
__device__ __forceinline__ float tc_pad_value(int k, int da) { return (k == da || k == da + 1) ? 1.f : 0.f; }

// ---- chain tiles: global memory <-> registers, transposed through tensor memory ---------------------------------------
// The compute layout gives thread (r, g) sixteen consecutive elements of each half of chain row r, so a warp-wide
// global access in that layout touches 32 different rows with 16 bytes each: 32 L1 lines per instruction and half-used
// sectors (measured: 16-22 k cycles per tile pair, ~45 % of the kernel).  Instead the tile goes through tensor memory,
// which can be addressed in two register layouts: the row-per-thread 32x32b shape the epilogues use, and the 16x256b
// fragment shape in which FOUR lanes hold 8 consecutive columns (32 contiguous bytes) of one row.  Global memory is
// accessed in the fragment layout (8 bytes per lane, whole sectors), tensor memory does the transposition:
//     load :  ld.global.v2 (fragment layout) -> tcgen05.st.16x256b -> barrier -> tcgen05.ld.32x32b (row layout)
//     store:  tcgen05.st.32x32b (row layout) -> barrier -> tcgen05.ld.16x256b -> st.global.v2 (fragment layout)
// The tile occupies columns [0, d) of its own 256-column region, in PHYSICAL element order (reversed when the flow has an
// odd number of permutations; the reversal is done on the global side, pairwise).  Warp (q, g) -- lane quarter q, column
// group g -- owns the 8-column chunks j = g, g + 4, g + 8, g + 12 of rows [32 q, 32 q + 32) in both directions.
constexpr int kTcMaxChunks = 4;
struct TcTileBuf {
  float2 v[kTcMaxChunks][2][2];   // [chunk][row half][row, row + 8]
};
__device__ __forceinline__ void tmem_st_frag(uint32_t taddr, float2 a, float2 b) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)),
               "r"(__float_as_uint(b.x)), "r"(__float_as_uint(b.y))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_frag(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
               "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
               "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
               : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, float v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(__float_as_uint(v)) : "memory");
}
// waits carry the loaded registers as operands so that no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&a)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]),
                 "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&a)[4], uint32_t (&b)[4], uint32_t (&c)[4], uint32_t (&d)[4]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(c[0]),
                 "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
               :
               : "memory");
}

// load, step 1: this warp's chunks of the tile, global -> registers (fragment layout).  tile_base: row 0 of the tile.
__device__ __forceinline__ void tc_tile_in_issue(const float* __restrict__ tile_base, long long rows_valid, int d, bool fl, int q, int g,
                                                 int lane, TcTileBuf& B) {
  const int rsub = lane >> 2, csub = 2 * (lane & 3);
#pragma unroll
  for (int c = 0; c < kTcMaxChunks; ++c) {
    const int col0 = 8 * (g + 4 * c) + csub;
    const bool cok = col0 < d;                       // d is even: the pair (col0, col0 + 1) is inside or outside as a whole
    const int gcol = fl ? d - 2 - col0 : col0;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int row = 32 * q + 16 * h + 8 * rr + rsub;
        float2 v = make_float2(0.f, 0.f);
        if (cok && row < rows_valid) {
          v = __ldg(reinterpret_cast<const float2*>(tile_base + (long long)row * d + gcol));
          if (fl) v = make_float2(v.y, v.x);
        }
        B.v[c][h][rr] = v;
      }
  }
}
// load, step 2: registers -> tensor memory (fragment layout).  tregion: first column of the tile's region.
__device__ __forceinline__ void tc_tile_in_commit(uint32_t tregion, int d, int q, int g, const TcTileBuf& B) {
  const int nch = (d + 7) >> 3;
#pragma unroll
  for (int c = 0; c < kTcMaxChunks; ++c) {
    const int j = g + 4 * c;
    if (j < nch) {
#pragma unroll
      for (int h = 0; h < 2; ++h) tmem_st_frag(tregion + ((uint32_t)(32 * q + 16 * h) << 16) + 8 * j, B.v[c][h][0], B.v[c][h][1]);
    }
  }
}
// load, step 3 (after tcgen05.wait::st, fence, barrier, fence): this thread's row pieces, tensor memory -> registers.
// Invalid slots (k >= d/2) get the constant-one columns of GEMM 1 at k = d/2, d/2 + 1 and zero elsewhere; nothing ever
// changes them (padded weight rows are zero, padded biases and affine slots are the identity).  trow: region + own lane.
__device__ __forceinline__ void tc_tile_in_rows(uint32_t trow, int da, int g, float (&lo)[kTcOwn], float (&hi)[kTcOwn]) {
  uint32_t a[16], b[16];
  tmem_ld16(trow + 16 * g, a);
  tmem_ld16(trow + da + 16 * g, b);
  tmem_wait_ld(a);
  tmem_wait_ld(b);
#pragma unroll
  for (int i = 0; i < kTcOwn; ++i) {
    const int k = 16 * g + i;
    const float pad = (k == da || k == da + 1) ? 1.f : 0.f;
    lo[i] = (k < da) ? __uint_as_float(a[i]) : pad;
    hi[i] = (k < da) ? __uint_as_float(b[i]) : pad;
  }
}
// store, step 1: this thread's valid row pieces, registers -> tensor memory
__device__ __forceinline__ void tc_tile_out_rows(uint32_t trow, int da, int g, const float (&lo)[kTcOwn], const float (&hi)[kTcOwn]) {
  const int nv = da - 16 * g;       // valid slots of this column group (warp-uniform)
  if (nv >= kTcOwn) {
    tmem_st16(trow + 16 * g, lo);
    tmem_st16(trow + da + 16 * g, hi);
  } else {
#pragma unroll
    for (int i = 0; i < kTcOwn; ++i)
      if (i < nv) {
        tmem_st1(trow + 16 * g + i, lo[i]);
        tmem_st1(trow + da + 16 * g + i, hi[i]);
      }
  }
}
// store, step 2 (after tcgen05.wait::st, fence, barrier, fence): this warp's chunks, tensor memory -> global memory
__device__ __forceinline__ void tc_tile_out_store(float* __restrict__ tile_base, long long rows_valid, int d, bool fl, uint32_t tregion,
                                                  int q, int g, int lane) {
  const int nch = (d + 7) >> 3;
  const int rsub = lane >> 2, csub = 2 * (lane & 3);
  uint32_t f[kTcMaxChunks][2][4];
#pragma unroll
  for (int c = 0; c < kTcMaxChunks; ++c) {
    const int j = g + 4 * c;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (j < nch) tmem_ld_frag(tregion + ((uint32_t)(32 * q + 16 * h) << 16) + 8 * j, f[c][h]);
      else f[c][h][0] = f[c][h][1] = f[c][h][2] = f[c][h][3] = 0u;
    }
  }
  tmem_wait_ld(f[0][0], f[0][1], f[1][0], f[1][1]);
  tmem_wait_ld(f[2][0], f[2][1], f[3][0], f[3][1]);
#pragma unroll
  for (int c = 0; c < kTcMaxChunks; ++c) {
    const int col0 = 8 * (g + 4 * c) + csub;
    if (col0 < d) {
      const int gcol = fl ? d - 2 - col0 : col0;
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int row = 32 * q + 16 * h + 8 * rr + rsub;
          if (row < rows_valid) {
            const float x = __uint_as_float(f[c][h][2 * rr]), y = __uint_as_float(f[c][h][2 * rr + 1]);
            *reinterpret_cast<float2*>(tile_base + (long long)row * d + gcol) = fl ? make_float2(y, x) : make_float2(x, y);
          }
        }
    }
  }
}

// staged tiles: this thread's pieces of chain row r in a tile buffer (row-major [128][d] fp32, d % 4 == 0) <-> registers,
// one half (HALF = 0: elements [0, d/2), 1: [d/2, d)) at a time.  Invalid slots as in tc_tile_in_rows.
template <int HALF>
__device__ __forceinline__ void tc_row_read_half(const float* row, int d, int da, int e0, bool fl, float (&v)[kTcOwn]) {
  if (!fl && e0 + kTcOwn <= da) {
    if (HALF == 0) {
      const float4* p4 = reinterpret_cast<const float4*>(row + e0);
#pragma unroll
      for (int i = 0; i < kTcOwn / 4; ++i) {
        const float4 w = p4[i];
        v[4 * i] = w.x; v[4 * i + 1] = w.y; v[4 * i + 2] = w.z; v[4 * i + 3] = w.w;
      }
    } else {
      const float2* p2 = reinterpret_cast<const float2*>(row + da + e0);
#pragma unroll
      for (int i = 0; i < kTcOwn / 2; ++i) {
        const float2 w = p2[i];
        v[2 * i] = w.x; v[2 * i + 1] = w.y;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < kTcOwn; ++i) {
      const int k = e0 + i, pos = HALF * da + k;
      const float pad = (k == da || k == da + 1) ? 1.f : 0.f;
      v[i] = (k < da) ? row[fl ? d - 1 - pos : pos] : pad;
    }
  }
}
template <int HALF>
__device__ __forceinline__ void tc_row_write_half(float* row, int d, int da, int e0, bool fl, const float (&v)[kTcOwn]) {
  if (!fl && e0 + kTcOwn <= da) {
    if (HALF == 0) {
      float4* p4 = reinterpret_cast<float4*>(row + e0);
#pragma unroll
      for (int i = 0; i < kTcOwn / 4; ++i) p4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
      float2* p2 = reinterpret_cast<float2*>(row + da + e0);
#pragma unroll
      for (int i = 0; i < kTcOwn / 2; ++i) p2[i] = make_float2(v[2 * i], v[2 * i + 1]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < kTcOwn; ++i) {
      const int k = e0 + i, pos = HALF * da + k;
      if (k < da) row[fl ? d - 1 - pos : pos] = v[i];
    }
  }
}
// One tile at a boundary: take this thread's pieces of the new input out of the tile buffer and (have_out) leave its
// pieces of the previous output there.  With unflipped output both occupy the same locations, so the exchange goes half
// by half through 16 temporaries; a flipped output lands on other threads' inputs, so everybody reads first.
__device__ __forceinline__ void tc_row_exchange(float* row, int d, int da, int e0, bool fl_in, bool fl_out, bool have_out,
                                                float (&lo)[kTcOwn], float (&hi)[kTcOwn]) {
  if (have_out && fl_out) {
    float nlo[kTcOwn], nhi[kTcOwn];
    tc_row_read_half<0>(row, d, da, e0, fl_in, nlo);
    tc_row_read_half<1>(row, d, da, e0, fl_in, nhi);
    tc_epi_barrier();
    tc_row_write_half<0>(row, d, da, e0, true, lo);
    tc_row_write_half<1>(row, d, da, e0, true, hi);
#pragma unroll
    for (int i = 0; i < kTcOwn; ++i) { lo[i] = nlo[i]; hi[i] = nhi[i]; }
  } else {
    float tmp[kTcOwn];
    tc_row_read_half<0>(row, d, da, e0, fl_in, tmp);
    if (have_out) tc_row_write_half<0>(row, d, da, e0, false, lo);
#pragma unroll
    for (int i = 0; i < kTcOwn; ++i) lo[i] = tmp[i];
    tc_row_read_half<1>(row, d, da, e0, fl_in, tmp);
    if (have_out) tc_row_write_half<1>(row, d, da, e0, false, hi);
#pragma unroll
    for (int i = 0; i < kTcOwn; ++i) hi[i] = tmp[i];
  }
}

// elementwise affine (forward {alpha, beta} or inverse {1/alpha, -beta/alpha}: the same fma; identity in the pad slots)
__device__ __forceinline__ void tc_affine(const float4* aff4, int idx, bool inv, int e0, float (&lo)[kTcOwn], float (&hi)[kTcOwn]) {
  const float4* tab = aff4 + (idx * 2 + (inv ? 1 : 0)) * 64 + e0;
#pragma unroll
  for (int q = 0; q < kTcOwn; ++q) {
    const float4 p = tab[q];
    lo[q] = fmaf(p.x, lo[q], p.y);
    hi[q] = fmaf(p.z, hi[q], p.w);
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
// K-steps are dealt round-robin to the four column groups of a row; the loads of step s + 4 are in flight while step s
// is converted.
__device__ __forceinline__ void tc_epi1_step(uint32_t dst, uint32_t (&a)[8], uint32_t (&b)[8]) {
  uint32_t p[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    p[i] = tanh_bf16x2(__uint_as_float(a[2 * i]), __uint_as_float(a[2 * i + 1]));
    if (i < kTcTanhMufuPairsHi) p[4 + i] = tanh_bf16x2(__uint_as_float(b[2 * i]), __uint_as_float(b[2 * i + 1]));
    else p[4 + i] = tanh_poly_bf16x2(__uint_as_float(b[2 * i]), __uint_as_float(b[2 * i + 1]));
  }
  tmem_st8(dst, p);
}
__device__ __forceinline__ void tc_epi1(uint32_t trow, int Hp, int g) {
  const int nsteps = Hp >> 4;
  const uint32_t hi_off = (uint32_t)(Hp >> 1);
  int s = g;
  if (s < nsteps) {
    uint32_t a[8], b[8];
    tmem_ld8(trow + 8 * s, a);
    tmem_ld8(trow + hi_off + 8 * s, b);
#pragma unroll 1
    while (true) {
      tmem_wait_ld(a, b);
      const int s2 = s + kTcGroups;
      if (s2 >= nsteps) break;
      uint32_t a2[8], b2[8];
      tmem_ld8(trow + 8 * s2, a2);
      tmem_ld8(trow + hi_off + 8 * s2, b2);
      tc_epi1_step(trow + 8 * s, a, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] = a2[i]; b[i] = b2[i]; }
      s = s2;
    }
    tc_epi1_step(trow + 8 * s, a, b);
  }
  tmem_wait_st();
}

// epilogue 2 of one tile: (u_a', u_b') = U + bl' -> alpha = 2^u_a' + m, beta = u_b'; affine update of this thread's 16
// targets; log2-determinant accumulation from pairwise products of the scales (one lg2 per two scales; the inverse
// direction needs the same products for its reciprocals: one rcp per two).  Scales are >= m = 1e-3, so a product cannot
// underflow; it overflows only if both scales exceed 1e19, where the log-determinant is reported as +inf.
template <bool INV>
__device__ __forceinline__ void tc_epi2_chunk(const uint32_t (&v)[8], const float4 bA, const float4 bB, float* tgt, float& ld2) {
  const float a0 = fast_ex2(__uint_as_float(v[0]) + bA.x) + kMinScale, a1 = fast_ex2(__uint_as_float(v[2]) + bA.z) + kMinScale;
  const float a2 = fast_ex2(__uint_as_float(v[4]) + bB.x) + kMinScale, a3 = fast_ex2(__uint_as_float(v[6]) + bB.z) + kMinScale;
  const float ub0 = __uint_as_float(v[1]) + bA.y, ub1 = __uint_as_float(v[3]) + bA.w;
  const float ub2 = __uint_as_float(v[5]) + bB.y, ub3 = __uint_as_float(v[7]) + bB.w;
  const float p01 = a0 * a1, p23 = a2 * a3;
  if (INV) {
    const float r01 = fast_rcp(p01), r23 = fast_rcp(p23);
    tgt[0] = (tgt[0] - ub0) * (r01 * a1);
    tgt[1] = (tgt[1] - ub1) * (r01 * a0);
    tgt[2] = (tgt[2] - ub2) * (r23 * a3);
    tgt[3] = (tgt[3] - ub3) * (r23 * a2);
  } else {
    tgt[0] = fmaf(a0, tgt[0], ub0);
    tgt[1] = fmaf(a1, tgt[1], ub1);
    tgt[2] = fmaf(a2, tgt[2], ub2);
    tgt[3] = fmaf(a3, tgt[3], ub3);
  }
  ld2 += lg2_any(p01) + lg2_any(p23);
}
template <bool INV>
__device__ __forceinline__ void tc_epi2(uint32_t tcol_u, const float* bl, int N2p, int g, float (&tgt)[kTcOwn], float& ld2) {
  // this thread's 16 targets = U columns [32 g, 32 g + 32): four 8-column chunks, loaded two at a time
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = 4 * g + 2 * h;                    // chunks c, c + 1; N2p is a multiple of 16: both or neither exist
    if (c * 8 < N2p) {
      uint32_t v0[8], v1[8];
      tmem_ld8(tcol_u + c * 8, v0);
      tmem_ld8(tcol_u + c * 8 + 8, v1);
      const float4* b4 = reinterpret_cast<const float4*>(bl + c * 8);
      const float4 bA = b4[0], bB = b4[1], bC = b4[2], bD = b4[3];
      tmem_wait_ld(v0, v1);
      tc_epi2_chunk<INV>(v0, bA, bB, &tgt[8 * h], ld2);
      tc_epi2_chunk<INV>(v1, bC, bD, &tgt[8 * h + 4], ld2);
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
  const int e0 = g * kTcOwn;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    mbar_wait(sy.bar(kTcBarG1 + t), par);
    TC_TRACE_EPI(22 + t);
    tc_fence_after();
    tc_epi1(trow + t * kTcRegion, S.Hp, g);
    tc_fence_before();
    mbar_arrive(sy.bar(kTcBarHid + t));
    TC_TRACE_EPI(24 + t);
  }
  const float* bl = sm.bl + (size_t)layer * S.N2p;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    mbar_wait(sy.bar(kTcBarG2 + t), par);
    TC_TRACE_EPI(26 + t);
    tc_fence_after();
    tc_epi2<INV>(trow + t * kTcRegion + kTcUCol, bl, S.N2p, g, st[t][1 - SRC], ld2[t]);
    tc_affine(sm.aff4, aff_after, INV, e0, st[t][0], st[t][1]);
    if (write_next) {
      // one A1 image shared by the two tiles: GEMM 1 of tile 0 (next coupling) must have consumed it before tile 1 writes
      if (t == 1 && S.na1 == 1) mbar_wait(sy.bar(kTcBarG1 + 0), par ^ 1);
      tc_write_a1(sm.a1(t), r, g, st[t][1 - SRC]);
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(sy.bar(kTcBarA1 + t));
    }
    TC_TRACE_EPI(28 + t);
  }
  sy.use += 1;
}

// start of a pass for one tile: first elementwise affine, first A1 image, hand-over to the MMA issuer
template <bool INV>
__device__ __forceinline__ void tc_pass_begin(const TcSmem& sm, const TcShape& S, TcEpiSync& sy, int t, int r, int g, float (&lo)[kTcOwn],
                                              float (&hi)[kTcOwn]) {
  const int l0 = INV ? S.Lc - 1 : 0;
  tc_affine(sm.aff4, INV ? S.Lc : 0, INV, g * kTcOwn, lo, hi);
  TC_TRACE_EPI(53 + t);
  if (t == 1 && S.na1 == 1) mbar_wait(sy.bar(kTcBarG1 + 0), sy.use & 1);   // shared A1 image: wait for tile 0's GEMM 1
  TC_TRACE_EPI(55 + t);
  if ((l0 & 1) == 0) tc_write_a1(sm.a1(t), r, g, hi);
  else tc_write_a1(sm.a1(t), r, g, lo);
  TC_TRACE_EPI(57 + t);
  fence_async_smem();
  tc_fence_before();
  TC_TRACE_EPI(59 + t);
  mbar_arrive(sy.bar(kTcBarA1 + t));
}
// the Lc couplings of a pass over both tiles.  st holds the input (after tc_pass_begin) on entry and the output on
// return; ld2 accumulates the sum of log2(alpha) (the caller adds the constant and flips the sign for the inverse).
template <bool INV>
__device__ __forceinline__ void tc_pass_couplings(const TcSmem& sm, const TcShape& S, TcEpiSync& sy, uint32_t trow, int r, int g,
                                                  float (&st)[2][2][kTcOwn], float (&ld2)[2]) {
  const int Lc = S.Lc;
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
