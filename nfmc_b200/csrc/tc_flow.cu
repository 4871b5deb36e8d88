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
template <bool INV, bool STAGED>
__global__ void __launch_bounds__(kTcThreads, 1) flow_tc_kernel(const __grid_constant__ TcArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  TcSmem sm = tc_carve(smem, A.S);
  const int tid = threadIdx.x, warp = tid >> 5;
  const TcShape& S = A.S;
  const int d = S.d, da = d / 2, Lc = S.Lc;

  tc_prologue(sm, A.blob, S);
  const long long tiles = (A.n + kTcRows - 1) / kTcRows;
  const long long pairs = (tiles + 1) / 2;
  long long my_pairs = 0;
  if ((long long)blockIdx.x < pairs) my_pairs = (pairs - 1 - blockIdx.x) / gridDim.x + 1;
  const uint32_t total_uses = (uint32_t)(my_pairs * Lc);

  if (warp >= kTcEpiWarps) {
    // ================================ service warps (uniform control flow, one elected lane issues) ================
    reg_dealloc<kTcRegsService>();
    const bool lead = elect_one();
    if (warp == kTcWarpMma) {
      TcMma mma(sm, S, INV ? 1u : 0u, lead);
      for (uint32_t u = 0; u < total_uses; ++u) mma.coupling();
    } else if (warp == kTcWarpWeights) {
      tc_weight_loader(sm, S, A.blob, total_uses, INV ? 1u : 0u, lead);
    } else if (warp == kTcWarpTiles) {
      if (lead) {
        if (STAGED) tc_x_loader(sm, S, A.in, A.out, A.n, my_pairs);
        else tc_tile_prefetch(A.in, A.n, d, my_pairs);
      }
    }
    __syncwarp();
  } else {
    // ================================ epilogue warps ================================================================
    reg_alloc<kTcRegsEpi>();
    const int r = tid & (kTcRows - 1), g = tid >> 7, q = (tid >> 5) & 3, lane = tid & 31;
    const int e0 = g * kTcOwn;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    TcEpiSync sync(sm);
    const bool flip = (Lc & 1) != 0;
    const bool fl_in = INV && flip, fl_out = !INV && flip;
    float st[2][2][kTcOwn];   // [tile][half][q]

    for (long long p = 0; p < my_pairs; ++p) {
      const long long tile0 = ((long long)blockIdx.x + p * gridDim.x) * 2;
      { TcEpiSync& sy = sync; TC_TRACE_EPI(39); }
      if (STAGED) {
        // ---- boundary p: take the input of this pair out of the tile buffers, leave the previous pair's output there -
        const bool have_out = p > 0 && A.out != nullptr;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          float* row = reinterpret_cast<float*>(sm.x(t)) + (size_t)r * d;
          mbar_wait(sync.bar(kTcBarXFull + t), (uint32_t)(p & 1));
          { TcEpiSync& sy = sync; TC_TRACE_EPI(43 + t); }
          tc_row_exchange(row, d, da, e0, fl_in, fl_out, have_out, st[t][0], st[t][1]);
          { TcEpiSync& sy = sync; TC_TRACE_EPI(49 + t); }
          if (have_out) fence_async_smem();
          { TcEpiSync& sy = sync; TC_TRACE_EPI(51 + t); }
          mbar_arrive(sync.bar(kTcBarXReady + t));
          { TcEpiSync& sy = sync; TC_TRACE_EPI(45 + t); }
          tc_pass_begin<INV>(sm, S, sync, t, r, g, st[t][0], st[t][1]);
          { TcEpiSync& sy = sync; TC_TRACE_EPI(47 + t); }
        }
      } else {
        // ---- input: global -> tensor memory (fragment layout) -> registers (row layout) ---------------------------
        {
          TcTileBuf B0, B1;
          tc_tile_in_issue(A.in + tile0 * kTcRows * (long long)d, A.n - tile0 * kTcRows, d, fl_in, q, g, lane, B0);
          tc_tile_in_issue(A.in + (tile0 + 1) * kTcRows * (long long)d, A.n - (tile0 + 1) * kTcRows, d, fl_in, q, g, lane, B1);
          tc_tile_in_commit(0, d, q, g, B0);
          tc_tile_in_commit(kTcRegion, d, q, g, B1);
        }
        tmem_wait_st();
        tc_fence_before();
        tc_epi_barrier();
        tc_fence_after();
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          tc_tile_in_rows(lane_off + t * kTcRegion, da, g, st[t][0], st[t][1]);
          tc_pass_begin<INV>(sm, S, sync, t, r, g, st[t][0], st[t][1]);
        }
      }
      float ld2[2] = {0.f, 0.f};
      { TcEpiSync& sy = sync; TC_TRACE_EPI(40); }
      tc_pass_couplings<INV>(sm, S, sync, lane_off, r, g, st, ld2);

      // ---- results ------------------------------------------------------------------------------------------------
      { TcEpiSync& sy = sync; TC_TRACE_EPI(41); }
      const bool frag_out = A.out != nullptr && !STAGED;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        float sq = 0.f;
        if (A.mode == 2) {
#pragma unroll
          for (int i = 0; i < kTcOwn; ++i)
            if (e0 + i < da) sq = fmaf(st[t][0][i], st[t][0][i], fmaf(st[t][1][i], st[t][1][i], sq));
        }
        sm.red[(t * 2 + 0) * kTcGroups * kTcRows + g * kTcRows + r] = ld2[t];
        sm.red[(t * 2 + 1) * kTcGroups * kTcRows + g * kTcRows + r] = sq;
        if (frag_out) tc_tile_out_rows(lane_off + t * kTcRegion, da, g, st[t][0], st[t][1]);
      }
      if (frag_out) {
        tmem_wait_st();
        tc_fence_before();
      }
      tc_epi_barrier();
      if (frag_out) {
        tc_fence_after();
#pragma unroll
        for (int t = 0; t < 2; ++t)
          tc_tile_out_store(A.out + (tile0 + t) * kTcRows * (long long)d, A.n - (tile0 + t) * kTcRows, d, fl_out, t * kTcRegion, q, g, lane);
      }
      // 512 threads, 256 rows: thread (r, g) with g < 2 finishes row r of tile g
      if (A.aux && g < 2 && (tile0 + g) * kTcRows + r < A.n) {
        const float* rd = sm.red + (g * 2) * kTcGroups * kTcRows;
        float res = (rd[r] + rd[kTcRows + r] + rd[2 * kTcRows + r] + rd[3 * kTcRows + r]) * 0.6931471805599453f + sm.log_const;
        if (INV) res = -res;
        if (A.mode == 2) {
          const float* rs = rd + kTcGroups * kTcRows;
          const float s = rs[r] + rs[kTcRows + r] + rs[2 * kTcRows + r] + rs[3 * kTcRows + r];
          res += -0.5f * s - 0.5f * (float)d * 1.8378770664093453f;
        }
        A.aux[(tile0 + g) * kTcRows + r] = res;
      }
      tc_epi_barrier();      // the scratch (it aliases the A1 image) is rewritten by the next pass
      { TcEpiSync& sy = sync; TC_TRACE_EPI(42); }
    }
    if (STAGED && my_pairs > 0) {
      // ---- last boundary: the output of the last pair goes into the tile buffers ---------------------------------------
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (A.out) {
          float* row = reinterpret_cast<float*>(sm.x(t)) + (size_t)r * d;
          tc_row_write_half<0>(row, d, da, e0, fl_out, st[t][0]);
          tc_row_write_half<1>(row, d, da, e0, fl_out, st[t][1]);
          fence_async_smem();
        }
        mbar_arrive(sync.bar(kTcBarXReady + t));
      }
    }
  }
  tc_epilogue_dealloc(0u);
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
  const bool staged_ok = (flow->d & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (!tc_plan_smem(A.S, 0, smem, staged_ok)) return set_error("flow_tc_pass: shared-memory plan exceeds 227 KB");
  if ((reinterpret_cast<uintptr_t>(in) & 7) || (reinterpret_cast<uintptr_t>(out) & 7)) return set_error("flow_tc_pass: in / out must be 8-byte aligned");
  A.blob = static_cast<const unsigned char*>(flow->blob);
  A.mode = mode; A.in = in; A.out = out; A.aux = aux; A.n = n;
  const long long pairs = ((n + kTcRows - 1) / kTcRows + 1) / 2;
  const int grid = (int)(pairs < sm_count() ? pairs : sm_count());
  auto launch = [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kernel<<<grid, kTcThreads, smem, (cudaStream_t)stream>>>(A);
  };
  if (A.S.nx == 2) { if (mode == 1) launch(flow_tc_kernel<true, true>); else launch(flow_tc_kernel<false, true>); }
  else { if (mode == 1) launch(flow_tc_kernel<true, false>); else launch(flow_tc_kernel<false, false>); }
  return check_cuda(cudaGetLastError(), "flow_tc_kernel launch");
}
