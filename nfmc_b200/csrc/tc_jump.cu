// tc_jump.cu -- one NF jump / one IMH iteration for a wide flow as ONE kernel on the tensor cores.
//
// Reference: JumpNFMC.sample's jump (nfmc/jump.py:203-243) and FixedIMH / AdaptiveIMH.sample (nfmc/imh.py:122-150,
// 214-249):  x', log q(x') = flow.sample;  log q(x) = flow.log_prob(x) (or its cache);  U(x), U(x');
// log alpha = -U(x') + U(x) + log q(x) - log q(x');  accept iff log u < log alpha;  overwrite, counters, moments.
//
// Per pair of 128-chain tiles a CTA (same warp roles, tensor-memory plan and weight pipeline as tc_flow.cu) does:
//   1. x tiles arrive in the shared-memory tile buffers by TMA;  every thread takes its pieces of chain row r
//      (the buffers KEEP x);  U(x) by a four-way row reduction
//   2. phase A (unless a valid log q cache is given): forward pass x -> z on the tensor cores, log q(x)
//   3. base draw z' from Philox (stream 1, keyed by global chain: the same numbers the CUDA-core path draws) or injected
//   4. phase B: inverse pass z' -> x' on the tensor cores, log q(x')
//   5. U(x'), log alpha, the accept test; accepted rows overwrite their pieces of the tile buffer
//   6. running moments of the post-jump state by column sums over the tile buffers, then the buffers go back to global
//      memory by TMA (and, if samples are stored, a second time into the sample sink)
// Chain state therefore makes exactly one HBM round trip per jump: 8 d + 8 bytes per chain (+ 4 d if stored).
#include <cuda_bf16.h>
#include "host_common.cuh"
#include "chain_kernel.cuh"
#include "tc_common.cuh"

namespace nfmc {

struct TcJumpArgs {
  const unsigned char* blob;
  TcShape S;
  int gs;                  // lanes per chain of the CUDA-core layout for this d (fixes the Philox counter of each element)
  int pot_kind;
  PotParams pot;
  float* x;                // [n, d], updated in place
  float* logq_cache;       // [n] or nullptr
  int phase_a;             // 1: compute log q(x) by a forward pass; 0: read it from logq_cache
  int adjusted;
  RngArgs rng;
  long long chain0;
  long long n;
  StatsArgs stats;
  float* sink_row;         // [n, d] destination of the post-jump state, or nullptr
};

// per-CTA scratch behind the common carve-up
struct TcJumpSmem {
  float* u_x;       // [2][128] U(x)
  float* lq_x;      // [2][128] log q(x)
  int* acc;         // [2][128] accept flags
  double* sx;       // [d] running sum of the post-jump state
  double* sx2;      // [d]
  unsigned long long* cnt;   // [4]
};
__host__ __device__ inline size_t tc_jump_extra_bytes(int d) {
  return (size_t)3 * 2 * kTcRows * 4 + (size_t)2 * d * 8 + 4 * 8;
}
__device__ __forceinline__ TcJumpSmem tc_jump_carve(unsigned char* p, int d) {
  TcJumpSmem j;
  j.sx = reinterpret_cast<double*>(p);
  j.sx2 = j.sx + d;
  j.cnt = reinterpret_cast<unsigned long long*>(j.sx2 + d);
  j.u_x = reinterpret_cast<float*>(j.cnt + 4);
  j.lq_x = j.u_x + 2 * kTcRows;
  j.acc = reinterpret_cast<int*>(j.lq_x + 2 * kTcRows);
  return j;
}

// ---- potentials in the (row, column group) layout --------------------------------------------------------------------
// Same closed forms as potentials.cuh (pot_prepare).  Every built-in potential is a sum of per-slot terms plus a few
// chain-level terms of x[0], x[1]; a thread adds up the terms of its 16 slots (Rosenbrock pairs (x_k, x_{k + d/2}) sit in
// one slot), the four column groups of a row are combined through shared memory, and the thread that owns x[0], x[1]
// (group 0) finishes the value.
__device__ __forceinline__ float tc_pot_partial(int kind, const PotParams& P, int da, int e0, const float (&lo)[kTcOwn], const float (&hi)[kTcOwn]) {
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
    for (int i = 0; i < kTcOwn; ++i) {
      if (e0 + i < da) {
        const float t = lo[i] - 1.f, r = hi[i] - lo[i] * lo[i];
        s += t * t + P.s0 * r * r;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < kTcOwn; ++i)
      if (e0 + i < da) s = fmaf(lo[i], lo[i], fmaf(hi[i], hi[i], s));
  }
  return s;
}
__device__ __forceinline__ float tc_pot_finish(int kind, const PotParams& P, int d, float S, float x0, float x1) {
  switch (kind) {
    case NFMC_POT_ISO_GAUSSIAN: return 0.5f * P.s0 * S;
    case NFMC_POT_DIAG_GAUSSIAN: return 0.5f * S;
    case NFMC_POT_ROSENBROCK: return S;
    case NFMC_POT_FUNNEL: {
      const float ex = __expf(-x0), rest = S - x0 * x0;
      return x0 * x0 * P.s1 + 0.5f * (float)(d - 1) * x0 + 0.5f * ex * rest;      // P.s1 = 1/(2 s^2)
    }
    default: {  // NFMC_POT_MIXTURE4
      const float a = P.s0, base = -0.5f * (S + 2.f * a * a);
      const float e0 = base + a * (x0 + x1), e1 = base + a * (x0 - x1), e2 = base + a * (-x0 + x1), e3 = base + a * (-x0 - x1);
      const float m = fmaxf(fmaxf(e0, e1), fmaxf(e2, e3));
      return -(m + __logf(__expf(e0 - m) + __expf(e1 - m) + __expf(e2 - m) + __expf(e3 - m)));
    }
  }
}

// ---- base draw in the (row, column group) layout ---------------------------------------------------------------------------
// The CUDA-core kernels draw, for chain c and flow step s, pair p = e + 1 of lane j = k % gs (slot e = k / gs) from
// Philox4x32-10 with counter (32 (p / 2) + j, stream 1 | step_hi << 8, step_lo, c) and turn it into the PHYSICAL slot
// (lo[k], hi[k]) by Box-Muller (common.cuh: draw_step_noise).  This thread owns k in [16 g, 16 g + 16): for each lane
// index j it walks its slots in order of e, so consecutive pairs share a Philox block.
template <int GS>
__device__ __forceinline__ void tc_draw_base(const RngKey& key, int g, int da, float (&lo)[kTcOwn], float (&hi)[kTcOwn]) {
  constexpr int EL = kTcOwn / GS;                  // slots per lane index in this thread's range
  const int e_first = EL * g;
#pragma unroll
  for (int j = 0; j < GS; ++j) {
    uint4 w = make_uint4(0, 0, 0, 0);
    int have = -1;
#pragma unroll
    for (int el = 0; el < EL; ++el) {
      const int i = el * GS + j, k = kTcOwn * g + i;
      const int p = e_first + el + 1, quad = p >> 1;
      if (quad != have) { w = rng_quad(key, quad, j); have = quad; }           // warp-uniform
      float z0, z1;
      if (p & 1) box_muller(w.z, w.w, z0, z1);
      else box_muller(w.x, w.y, z0, z1);
      lo[i] = (k < da) ? z0 : tc_pad_value(k, da);
      hi[i] = (k < da) ? z1 : tc_pad_value(k, da);
    }
  }
}

// ---- tile-loader lane ---------------------------------------------------------------------------------------------------------
// pair k: load both x tiles -> (epilogue: both passes, merge, moments) -> xready[t] -> store the tile back over x (and into
// the sample sink) -> wait until the stores have read the buffer -> next pair
__device__ __forceinline__ void tc_jump_loader(const TcSmem& sm, const TcShape& S, float* x, float* sink_row, long long n, int my_pairs) {
  auto tile_of = [&](int k, int t) { return ((long long)blockIdx.x + (long long)k * gridDim.x) * 2 + t; };
  auto rows_of = [&](long long tile) { long long r = n - tile * kTcRows; return r > kTcRows ? (long long)kTcRows : (r < 0 ? 0ll : r); };
  for (int k = 0; k < my_pairs; ++k) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const long long tile = tile_of(k, t), rows = rows_of(tile);
      const uint32_t bar = smem_u32(sm.bars + kTcBarXFull + t);
      if (rows > 0) {
        const uint32_t bytes = (uint32_t)(rows * S.d * 4);
        mbar_expect_tx(bar, bytes);
        tma_bulk_load(smem_u32(sm.x(t)), x + tile * kTcRows * (long long)S.d, bytes, bar);
      } else {
        mbar_arrive(bar);
      }
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      mbar_wait(smem_u32(sm.bars + kTcBarXReady + t), (uint32_t)(k & 1));
      const long long tile = tile_of(k, t), rows = rows_of(tile);
      if (rows > 0) {
        const uint32_t bytes = (uint32_t)(rows * S.d * 4);
        tma_bulk_store(x + tile * kTcRows * (long long)S.d, smem_u32(sm.x(t)), bytes);
        if (sink_row) tma_bulk_store(sink_row + tile * kTcRows * (long long)S.d, smem_u32(sm.x(t)), bytes);
      }
    }
    tma_store_wait_read<0>();
  }
  tma_store_wait_all();
}

// ---- the kernel ------------------------------------------------------------------------------------------------------------------
template <int GS>
__global__ void __launch_bounds__(kTcThreads, 1) jump_tc_kernel(const __grid_constant__ TcJumpArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  const TcShape& S = A.S;
  const int d = S.d, da = d / 2, Lc = S.Lc;
  TcSmem sm = tc_carve(smem, S, tc_jump_extra_bytes(d));
  const TcJumpSmem js = tc_jump_carve(sm.extra, d);
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < 2 * d; i += kTcThreads) js.sx[i] = 0.0;
  if (tid < 4) js.cnt[tid] = 0ull;
  tc_prologue(sm, A.blob, S);
  const long long tiles = (A.n + kTcRows - 1) / kTcRows;
  const long long pairs = (tiles + 1) / 2;
  int my_pairs = 0;
  if ((long long)blockIdx.x < pairs) my_pairs = (int)((pairs - 1 - blockIdx.x) / gridDim.x + 1);
  const uint32_t seq = A.phase_a ? 2u : 1u;
  const uint32_t total_uses = (uint32_t)my_pairs * (uint32_t)Lc * (A.phase_a ? 2u : 1u);

  if (warp >= kTcEpiWarps) {
    reg_dealloc<kTcRegsService>();
    const bool lead = elect_one();
    if (warp == kTcWarpMma) {
      TcMma mma(sm, S, seq, lead);
      for (uint32_t u = 0; u < total_uses; ++u) mma.coupling();
    } else if (warp == kTcWarpWeights) {
      tc_weight_loader(sm, S, A.blob, total_uses, seq, lead);
    } else if (warp == kTcWarpTiles) {
      if (lead) tc_jump_loader(sm, S, A.x, A.sink_row, A.n, my_pairs);
    }
    __syncwarp();
  } else {
    reg_alloc<kTcRegsEpi>();
    const int r = tid & (kTcRows - 1), g = tid >> 7, q = (tid >> 5) & 3;
    const int e0 = g * kTcOwn;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    TcEpiSync sync(sm);
    const bool flip = (Lc & 1) != 0;
    const float log_norm = -0.5f * (float)d * 1.8378770664093453f;   // -d/2 log(2 pi)
    float* red = sm.red;                        // [2 tiles][2 values][4 groups][128 rows]
    unsigned int n_acc = 0, n_bad = 0, n_rows = 0;
    float st[2][2][kTcOwn];

    for (int p = 0; p < my_pairs; ++p) {
      const long long tile0 = ((long long)blockIdx.x + (long long)p * gridDim.x) * 2;
      // ---- 1. x out of the tile buffers; U(x) ---------------------------------------------------------------------------
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const float* row = reinterpret_cast<const float*>(sm.x(t)) + (size_t)r * d;
        mbar_wait(sync.bar(kTcBarXFull + t), (uint32_t)(p & 1));
        tc_row_read_half<0>(row, d, da, e0, false, st[t][0]);
        tc_row_read_half<1>(row, d, da, e0, false, st[t][1]);
        red[(t * 2) * kTcGroups * kTcRows + g * kTcRows + r] = tc_pot_partial(A.pot_kind, A.pot, da, e0, st[t][0], st[t][1]);
      }
      tc_epi_barrier();
      if (g == 0) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const float* rd = red + (t * 2) * kTcGroups * kTcRows;
          const float Ssum = rd[r] + rd[kTcRows + r] + rd[2 * kTcRows + r] + rd[3 * kTcRows + r];
          js.u_x[t * kTcRows + r] = tc_pot_finish(A.pot_kind, A.pot, d, Ssum, st[t][0][0], da >= 2 ? st[t][0][1] : st[t][1][0]);
          if (!A.phase_a && A.adjusted) {
            const long long row = (tile0 + t) * kTcRows + r;
            js.lq_x[t * kTcRows + r] = row < A.n ? __ldg(A.logq_cache + row) : 0.f;      // imh.py:214
          }
        }
      }
      tc_epi_barrier();
      // ---- 2. phase A: log q(x) = log N(T(x)) + log|det dT/dx|  (jump.py:218, imh.py:133) ---------------------------------------
      if (A.phase_a) {
        float ld2[2] = {0.f, 0.f};
#pragma unroll
        for (int t = 0; t < 2; ++t) tc_pass_begin<false>(sm, S, sync, t, r, g, st[t][0], st[t][1]);
        tc_pass_couplings<false>(sm, S, sync, lane_off, r, g, st, ld2);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          float sq = 0.f;
#pragma unroll
          for (int i = 0; i < kTcOwn; ++i)
            if (e0 + i < da) sq = fmaf(st[t][0][i], st[t][0][i], fmaf(st[t][1][i], st[t][1][i], sq));
          red[(t * 2 + 0) * kTcGroups * kTcRows + g * kTcRows + r] = ld2[t];
          red[(t * 2 + 1) * kTcGroups * kTcRows + g * kTcRows + r] = sq;
        }
        tc_epi_barrier();
        if (g < 2) {
          const float* rd = red + (g * 2) * kTcGroups * kTcRows;
          const float* rs = rd + kTcGroups * kTcRows;
          const float ld = (rd[r] + rd[kTcRows + r] + rd[2 * kTcRows + r] + rd[3 * kTcRows + r]) * 0.6931471805599453f + sm.log_const;
          const float sq = rs[r] + rs[kTcRows + r] + rs[2 * kTcRows + r] + rs[3 * kTcRows + r];
          js.lq_x[g * kTcRows + r] = ld - 0.5f * sq + log_norm;
        }
        tc_epi_barrier();
      }
      // ---- 3. base draw z' (physical order), its squared norm -----------------------------------------------------------------
      float sqz[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const long long row_raw = (tile0 + t) * kTcRows + r;
        const long long row = row_raw < A.n ? row_raw : A.n - 1;
        if (A.rng.normals) {
          // injected: the LOGICAL z the reference's flow.sample would have drawn (flipped into physical order if Lc is odd)
          const float* zr = A.rng.normals + row * (long long)d;
#pragma unroll
          for (int i = 0; i < kTcOwn; ++i) {
            const int k = e0 + i;
            st[t][0][i] = (k < da) ? __ldg(zr + (flip ? d - 1 - k : k)) : tc_pad_value(k, da);
            st[t][1][i] = (k < da) ? __ldg(zr + (flip ? d - 1 - (da + k) : da + k)) : tc_pad_value(k, da);
          }
        } else {
          const RngKey key = make_rng_key(A.rng.seed, 1u, A.rng.step0, (uint64_t)(A.chain0 + row));
          tc_draw_base<GS>(key, g, da, st[t][0], st[t][1]);
        }
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < kTcOwn; ++i)
          if (e0 + i < da) sq = fmaf(st[t][0][i], st[t][0][i], fmaf(st[t][1][i], st[t][1][i], sq));
        sqz[t] = sq;
      }
      // ---- 4. phase B: x' = T^-1(z'), log|det dx'/dz'| (jump.py:205, imh.py:221) -------------------------------------------------
      float ld2[2] = {0.f, 0.f};
#pragma unroll
      for (int t = 0; t < 2; ++t) tc_pass_begin<true>(sm, S, sync, t, r, g, st[t][0], st[t][1]);
      tc_pass_couplings<true>(sm, S, sync, lane_off, r, g, st, ld2);
      // ---- 5. U(x'), log alpha, accept ------------------------------------------------------------------------------------------------
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        red[(t * 2 + 0) * kTcGroups * kTcRows + g * kTcRows + r] = ld2[t] * 0.6931471805599453f - 0.5f * sqz[t];
        red[(t * 2 + 1) * kTcGroups * kTcRows + g * kTcRows + r] = tc_pot_partial(A.pot_kind, A.pot, da, e0, st[t][0], st[t][1]);
      }
      tc_epi_barrier();
      if (g == 0) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const long long row = (tile0 + t) * kTcRows + r;
          const bool active = row < A.n;
          const float* rd = red + (t * 2) * kTcGroups * kTcRows;
          const float* ru = rd + kTcGroups * kTcRows;
          // log q(x') = log N(z') - log|det dx'/dz'|, and the inverse pass accumulates -log|det|
          const float lq_p = (rd[r] + rd[kTcRows + r] + rd[2 * kTcRows + r] + rd[3 * kTcRows + r]) + sm.log_const + log_norm;
          const float u_p = tc_pot_finish(A.pot_kind, A.pot, d, ru[r] + ru[kTcRows + r] + ru[2 * kTcRows + r] + ru[3 * kTcRows + r],
                                          st[t][0][0], da >= 2 ? st[t][0][1] : st[t][1][0]);
          bool accept = true;
          if (A.adjusted) {
            const float log_alpha = (-u_p) - (-js.u_x[t * kTcRows + r]) + js.lq_x[t * kTcRows + r] - lq_p;   // util.py:392
            float u;
            if (A.rng.uniforms) u = __ldg(A.rng.uniforms + (active ? row : A.n - 1));
            else {
              const RngKey key = make_rng_key(A.rng.seed, 1u, A.rng.step0, (uint64_t)(A.chain0 + (active ? row : A.n - 1)));
              u = uniform_from_bits(rng_quad(key, 0, 0).x);
            }
            accept = logf(u) < log_alpha;                                                    // jump.py:225
            if (active && !(fabsf(log_alpha) <= 3.0e38f)) ++n_bad;
          }
          js.acc[t * kTcRows + r] = accept ? 1 : 0;
          if (active) {
            ++n_rows;
            if (accept) {
              ++n_acc;
              if (A.adjusted && A.logq_cache) A.logq_cache[row] = lq_p;                      // imh.py:233
            }
          }
        }
      }
      tc_epi_barrier();
      // ---- 6. accepted rows overwrite their pieces of the tile buffers; moments of the post-jump state; hand the buffers back ----
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (js.acc[t * kTcRows + r]) {
          float* row = reinterpret_cast<float*>(sm.x(t)) + (size_t)r * d;
          tc_row_write_half<0>(row, d, da, e0, false, st[t][0]);
          tc_row_write_half<1>(row, d, da, e0, false, st[t][1]);
        }
      }
      fence_async_smem();
      tc_epi_barrier();
      if (A.stats.sum_x) {
        // column sums over the (up to) 256 rows of the pair: thread -> (column c, row group); fp32 over <= 64 rows, then fp64
        const int ngrp = kTcEpiThreads / d;               // >= 4 for d <= 128
        const int c = tid % d, grp = tid / d;
        if (grp < ngrp) {
          long long valid = A.n - tile0 * kTcRows;
          if (valid > 2 * kTcRows) valid = 2 * kTcRows;
          const int per = (2 * kTcRows + ngrp - 1) / ngrp;
          const int r0 = grp * per, r1 = min((int)valid, r0 + per);
          float s1 = 0.f, s2 = 0.f;
          for (int rr = r0; rr < r1; ++rr) {
            const float v = reinterpret_cast<const float*>(sm.x(rr >> 7))[(size_t)(rr & 127) * d + c];
            s1 += v;
            s2 = fmaf(v, v, s2);
          }
          if (r1 > r0) {
            atomicAdd(js.sx + c, (double)s1);
            atomicAdd(js.sx2 + c, (double)s2);
          }
        }
      }
#pragma unroll
      for (int t = 0; t < 2; ++t) mbar_arrive(sync.bar(kTcBarXReady + t));
    }
    // ---- counters ---------------------------------------------------------------------------------------------------------------------
    n_acc = __reduce_add_sync(0xffffffffu, n_acc);
    n_bad = __reduce_add_sync(0xffffffffu, n_bad);
    n_rows = __reduce_add_sync(0xffffffffu, n_rows);
    if ((tid & 31) == 0) {
      if (n_acc) atomicAdd(js.cnt + 0, (unsigned long long)n_acc);
      if (n_rows) atomicAdd(js.cnt + 1, (unsigned long long)n_rows);
      if (n_bad) atomicAdd(js.cnt + 2, (unsigned long long)n_bad);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (A.stats.sum_x && A.stats.sum_x2)
    for (int i = tid; i < d; i += kTcThreads) {
      atomicAdd(A.stats.sum_x + i, js.sx[i]);
      atomicAdd(A.stats.sum_x2 + i, js.sx2[i]);
    }
  if (A.stats.counts && tid < 4 && js.cnt[tid]) atomicAdd(A.stats.counts + tid, js.cnt[tid]);
  if (warp == 0) tmem_dealloc(0u, 512);
}

}  // namespace nfmc

using namespace nfmc;

// Fused jump on the tensor cores.  Returns 0 on launch, 1 on error, -1 if this shape / these tensors are not eligible (the
// caller then composes the jump from nfmc_flow_tc_pass launches, flow_api.cu).
int nfmc_jump_step_tc_fused(const nfmc_potential* pot, const nfmc_realnvp_tc* flow, float* x, float* logq_cache, int recompute_logq,
                            int64_t n, int adjusted, const nfmc_rng* rng, int64_t chain0, const nfmc_stats* stats,
                            const nfmc_sink* sink, void* stream) {
  TcJumpArgs A;
  if (int e = tc_validate(flow, A.S, "jump_step_tc")) return e;
  const int d = flow->d;
  float* sink_row = nullptr;
  if (sink && sink->samples) {
    const int64_t th = sink->thinning > 0 ? sink->thinning : 1;
    if (sink->seen0 % th == 0) sink_row = sink->samples;
  }
  const bool staged_ok = (d & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(sink_row) & 15) == 0;
  size_t smem = 0;
  if (!staged_ok || !tc_plan_smem(A.S, tc_jump_extra_bytes(d), smem, true) || A.S.nx != 2) return -1;
  Layout L;
  if (!layout_for_dim(d, L) || (L.gs != 1 && L.gs != 2 && L.gs != 4)) return -1;
  A.blob = static_cast<const unsigned char*>(flow->blob);
  A.gs = L.gs;
  A.pot_kind = pot->kind;
  A.pot = pot_params(pot);
  A.x = x;
  A.logq_cache = logq_cache;
  A.phase_a = (adjusted && (!logq_cache || recompute_logq)) ? 1 : 0;
  A.adjusted = adjusted;
  A.rng = RngArgs{rng->seed, rng->step0, rng->normals, rng->uniforms};
  A.chain0 = chain0;
  A.n = n;
  A.stats = StatsArgs{stats ? stats->sum_x : nullptr, stats ? stats->sum_x2 : nullptr, stats ? stats->counts : nullptr};
  A.sink_row = sink_row;
  const long long pairs = ((n + kTcRows - 1) / kTcRows + 1) / 2;
  const int grid = (int)(pairs < sm_count() ? pairs : sm_count());
  auto launch = [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kernel<<<grid, kTcThreads, smem, (cudaStream_t)stream>>>(A);
  };
  if (L.gs == 1) launch(jump_tc_kernel<1>);
  else if (L.gs == 2) launch(jump_tc_kernel<2>);
  else launch(jump_tc_kernel<4>);
  return check_cuda(cudaGetLastError(), "jump_tc_kernel launch");
}
