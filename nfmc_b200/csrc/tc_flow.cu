// tc_flow.cu -- RealNVP passes with the conditioner MLP on the 5th-generation tensor cores (tcgen05 + TMEM),
// warp-specialised and software-pipelined over two chain tiles.
//
// For 2-layer conditioners (hidden width zero-padded to Hp, a multiple of 16 <= 256; even d <= 128) the two GEMMs of
// a coupling layer
//       Hpre[128 x Hp]  = [S | 1 | 1][128 x K1] . [W1 | b1_hi | b1_lo]^T      (K1 = d/2 + 2 padded to 16)
//       U   [128 x N2p] = tanh(Hpre)[128 x Hp] . Wl'^T                          (K = Hp)
// run as tcgen05.mma (kind::f16, bf16 operands, fp32 accumulators in tensor memory) on tiles of 128 chains.
//
// One persistent CTA per SM, 17 warps:
//   * warps 0..15 (512 threads) are EPILOGUE warps.  Thread (r, g) = (tid % 128, tid / 128) works on chain row r of a
//     tile and owns elements [16 g, 16 g + 16) of each half of that chain, fp32, in registers -- for TWO tiles at once.
//   * warp 16 is the CONTROL warp: one lane issues every TMA weight copy and every tcgen05.mma.
// The two tiles ping-pong through the coupling: while the epilogue warps run tanh on tile 0, the tensor pipe computes
// tile 1's first GEMM; while they run tile 1's tanh it computes tile 0's second GEMM, and so on.  Both tiles use the
// same coupling weights, which are therefore staged once per tile PAIR (cp.async.bulk + mbarrier complete_tx, a ring
// of one or two buffers per weight image, or all couplings resident when they fit).
//
// Tensor-memory plan (512 columns, region of 256 per tile):
//   [0, Hp)           Hpre accumulator of GEMM 1 (fp32)
//   [8 s, 8 s + 8)    after epilogue 1: hidden activations of K-step s as packed bf16 pairs -- the A operand of GEMM 2
//                     is read by the tensor core STRAIGHT FROM TENSOR MEMORY (tcgen05.mma with a TMEM A operand), so
//                     the activations never touch shared memory.  K-step s holds hidden units [8 s, 8 s + 8) and
//                     [Hp/2 + 8 s, Hp/2 + 8 s + 8): exactly the Hpre columns the writing thread has just read, so no
//                     thread overwrites a column another thread still needs; the B descriptor's leading-dimension
//                     byte offset jumps over Hp/16 k-groups to fetch the matching rows of Wl.
//   [128, 128 + N2p)  U accumulator of GEMM 2 (fp32)
//
// Epilogue 1: tanh.approx.bf16x2 (one MUFU per two hidden units; b1 arrives through the GEMM as two bf16 constant-one
// columns, hi + lo, i.e. with ~16 bits).  Epilogue 2: alpha = 2^(u_a') + m, beta = u_b' with the 1/2, log(1-m) and
// log2(e) factors folded into Wl' / bl' at pack time; the log-determinant takes one lg2 per FOUR scales (product), the
// inverse direction one rcp per TWO.
//
// Same specification as the fp32 path (oracle/realnvp_ref.py); parity tolerance is the bf16 one of the north star
// (rtol 1e-2).  Replaces flow.bijection.forward / inverse / flow.log_prob for wide flows (neutra.py:60, jump.py:205,218,
// imh.py:214,221).
#include <cuda_bf16.h>
#include "host_common.cuh"
#include "tc_common.cuh"

namespace nfmc {

struct TcArgs {
  const unsigned char* blob;  // tc_common.cuh: affines | per coupling { W1 image, Wl image, bl' }
  TcShape S;
  int mode;                   // 0 forward, 1 inverse, 2 log_prob
  const float* in;
  float* out;
  float* aux;
  long long n;
};

// ---- the kernel ----------------------------------------------------------------------------------------------------
template <bool INV>
__global__ void __launch_bounds__(kTcThreads, 1) flow_tc_kernel(const TcArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  TcSmem sm = tc_carve(smem, A.S);
  const int tid = threadIdx.x, warp = tid >> 5;
  const TcShape& S = A.S;
  const int d = S.d, da = d / 2, Lc = S.Lc;

  const uint32_t tmem_base = tc_prologue(sm, A.blob, S);
  const long long tiles = (A.n + kTcRows - 1) / kTcRows;
  const long long pairs = (tiles + 1) / 2;
  long long my_pairs = 0;
  if ((long long)blockIdx.x < pairs) my_pairs = (pairs - 1 - blockIdx.x) / gridDim.x + 1;
  const uint32_t total_uses = (uint32_t)(my_pairs * Lc);

  if (warp == kTcEpiWarps) {
    // ================================ control warp ==================================================================
    if ((tid & 31) == 0) {
      TcControl ctl(sm, S, A.blob, tmem_base, total_uses, INV ? 1u : 0u);
      ctl.prime();
      for (long long p = 0; p < my_pairs; ++p) {
        // pull the NEXT pair's rows towards L2 while this pair computes
        const long long next_row0 = ((long long)blockIdx.x + (p + 1) * gridDim.x) * 2 * kTcRows;
        if (p + 1 < my_pairs) {
          long long rows = A.n - next_row0;
          if (rows > 2 * kTcRows) rows = 2 * kTcRows;
          const size_t bytes = ((size_t)rows * d * sizeof(float)) & ~size_t(15);
          const float* src = A.in + next_row0 * d;
          if (bytes >= 16 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) tma_prefetch_l2(src, (uint32_t)bytes);
        }
        for (int i = 0; i < Lc; ++i) ctl.coupling();
      }
    }
    __syncwarp();
  } else {
    // ================================ epilogue warps ================================================================
    const int r = tid & (kTcRows - 1), g = tid >> 7;
    const int e0 = g * kTcOwn;
    const uint32_t lane_off = (uint32_t)((r >> 5) * 32) << 16;
    TcEpiSync sync(sm);
    const bool flip = (Lc & 1) != 0;
    float st[2][2][kTcOwn];   // [tile][half][q]

    for (long long p = 0; p < my_pairs; ++p) {
      const long long row0 = ((long long)blockIdx.x + p * gridDim.x) * 2 * kTcRows + r;   // tile t: row0 + 128 t
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const long long rr = row0 + t * kTcRows;
        tc_load_state(A.in + (rr < A.n ? rr : A.n - 1) * (long long)d, d, da, e0, INV && flip, st[t][0], st[t][1]);
      }
      float ld2[2] = {0.f, 0.f};
      { TcEpiSync& sy = sync; TC_TRACE_EPI(40); }
      tc_run_pass<INV>(sm, S, sync, tmem_base + lane_off, r, g, st, ld2);

      // ---- results ------------------------------------------------------------------------------------------------
      { TcEpiSync& sy = sync; TC_TRACE_EPI(41); }
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const long long rr = row0 + t * kTcRows;
        float sq = 0.f;
        if (A.mode == 2) {
#pragma unroll
          for (int q = 0; q < kTcOwn; ++q)
            if (e0 + q < da) sq = fmaf(st[t][0][q], st[t][0][q], fmaf(st[t][1][q], st[t][1][q], sq));
        }
        sm.red[(t * 2 + 0) * kTcGroups * kTcRows + g * kTcRows + r] = ld2[t];
        sm.red[(t * 2 + 1) * kTcGroups * kTcRows + g * kTcRows + r] = sq;
        if (rr < A.n && A.out) tc_store_state(A.out + rr * (long long)d, d, da, e0, !INV && flip, st[t][0], st[t][1]);
      }
      tc_epi_barrier();
      // 512 threads, 256 rows: thread (r, g) with g < 2 finishes row r of tile g
      if (A.aux && g < 2 && row0 + g * kTcRows < A.n) {
        const float* rd = sm.red + (g * 2) * kTcGroups * kTcRows;
        float res = (rd[r] + rd[kTcRows + r] + rd[2 * kTcRows + r] + rd[3 * kTcRows + r]) * 0.6931471805599453f +
                    sm.aff[(Lc + 1) * 4 * d];
        if (INV) res = -res;
        if (A.mode == 2) {
          const float* rs = rd + kTcGroups * kTcRows;
          const float s = rs[r] + rs[kTcRows + r] + rs[2 * kTcRows + r] + rs[3 * kTcRows + r];
          res += -0.5f * s - 0.5f * (float)d * 1.8378770664093453f;
        }
        A.aux[row0 + g * kTcRows] = res;
      }
      tc_epi_barrier();
      { TcEpiSync& sy = sync; TC_TRACE_EPI(42); }
    }
  }
  tc_epilogue_dealloc(tmem_base);
}

}  // namespace nfmc

using namespace nfmc;

#ifdef NFMC_TC_TRACE
extern "C" __attribute__((visibility("default"))) int nfmc_tc_trace_set(long long* buf) {
  return (int)cudaMemcpyToSymbol(g_tc_trace, &buf, sizeof(buf));
}
#endif

extern "C" int64_t nfmc_realnvp_tc_blob_bytes(int32_t d, int32_t n_coupling, int32_t hidden) {
  TcShape S;
  if (!tc_shape(d, n_coupling, hidden, S)) return -1;
  return (int64_t)(tc_affine_bytes(d, n_coupling) + (size_t)n_coupling * tc_coupling_bytes(S));
}

extern "C" int nfmc_flow_tc_pass(const nfmc_realnvp_tc* flow, int32_t mode, const float* in, float* out, float* aux,
                                 int64_t n, void* stream) {
  if (!flow || !flow->blob || !in || n < 1) return set_error("flow_tc_pass: bad arguments");
  if (mode < 0 || mode > 2) return set_error("flow_tc_pass: mode must be 0 (forward), 1 (inverse) or 2 (log_prob)");
  TcArgs A;
  if (int e = tc_validate(flow, A.S, "flow_tc_pass")) return e;
  size_t smem = 0;
  if (!tc_plan_smem(A.S, 0, smem)) return set_error("flow_tc_pass: shared-memory plan exceeds 227 KB");
  A.blob = static_cast<const unsigned char*>(flow->blob);
  A.mode = mode; A.in = in; A.out = out; A.aux = aux; A.n = n;
  const long long pairs = ((n + kTcRows - 1) / kTcRows + 1) / 2;
  const int grid = (int)(pairs < sm_count() ? pairs : sm_count());
  if (mode == 1) {
    cudaFuncSetAttribute(flow_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    flow_tc_kernel<true><<<grid, kTcThreads, smem, (cudaStream_t)stream>>>(A);
  } else {
    cudaFuncSetAttribute(flow_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    flow_tc_kernel<false><<<grid, kTcThreads, smem, (cudaStream_t)stream>>>(A);
  }
  return check_cuda(cudaGetLastError(), "flow_tc_kernel launch");
}
