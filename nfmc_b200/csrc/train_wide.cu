// train_wide.cu -- RealNVP training for wide / deep conditioners (any number of linear layers M >= 1, any hidden width H):
// loss and the gradient with respect to every flow parameter, straight in MODULE ORDER (theta in, grad_theta out -- no
// blob pack / unpack), for the objectives the reference trains flows with:
//   * maximum likelihood (Flow.fit: jump.py:139-151,201; imh.py:171-175):  mode NLL = x -> z pass + backward sweep
//   * reverse KL (Flow.variational_fit: imh.py:67-72; neutra.py:84-91):    mode PASS (z -> x) + potential kernel + mode SWEEP
// The register-resident kernel (train_kernels.cu) covers M = 2, H <= 8; this one covers everything else.
//
// Row-tile design.  A CTA of 256 threads owns R rows at a time; the tile's state V[R][d], its cotangent G[R][d], the
// conditioner's source copy, hidden activations, outputs and hidden cotangents all live in shared memory, the weights
// stream from L2 (they are shared by every CTA).  The three contractions of a linear layer
//     forward  out[r][n] = b[n] + sum_k in[r][k] W[n][k]          dgrad  din[r][k] = sum_n dout[r][n] W[n][k]
//     wgrad    dW[n][k] += sum_r dout[r][n] in[r][k]
// run as register-blocked fp32 loops: a thread owns one output column for 8 rows (forward / dgrad: one weight load feeds 8
// FMAs, the tile operand arrives by 128-bit shared-memory broadcasts) or a 4 x 1 block of dW (wgrad: reduced over the
// tile's rows in registers, then one red.global per weight and tile).  fp32 throughout: training gradients are checked
// against autograd of the oracle at rtol 1e-4.
//
// As in train_kernels.cu no activation outlives its coupling: the backward sweep re-derives a layer's input from its
// output (the flow is invertible) and re-evaluates the conditioner on the (unchanged) source half.
//
// theta layout (state_dict order of nfmc_b200.flow.RealNVP / oracle.realnvp_ref):
//   affine_0.value[d][2] | Lc x { linear_0 {W[out][in], b[out]} ... linear_{M-1} | actnorm_l.value[d][2] } |
//   affine_T.value[d][2] | actnorm_T.value[d][2];   value[i] = (u_a, u_b): alpha = exp(log(1-m) + u_a/2) + m, beta = u_b/2.
// Reverse permutations are index arithmetic on the shared-memory tile (logical index i lives in column i or d-1-i).
#include <algorithm>
#include "host_common.cuh"
#include "common.cuh"

namespace nfmc {

constexpr int kWT = 256;   // threads per CTA
constexpr int kWRT = 8;    // rows per thread in the forward / dgrad loops

struct WideDims {
  int d, da, db, Lc, M, H;
  __host__ __device__ int lin_in(int m) const { return m == 0 ? da : H; }
  __host__ __device__ int lin_out(int m) const { return m == M - 1 ? 2 * db : H; }
  __host__ __device__ long long cpl_w() const {
    long long s = 0;
    for (int m = 0; m < M; ++m) s += (long long)lin_out(m) * lin_in(m) + lin_out(m);
    return s;
  }
  __host__ __device__ long long cpl_theta() const { return cpl_w() + 2ll * d; }
  __host__ __device__ long long n_theta() const { return 2ll * d + (long long)Lc * cpl_theta() + 4ll * d; }
  __host__ __device__ long long off_cpl(int l) const { return 2ll * d + (long long)l * cpl_theta(); }
  __host__ __device__ long long off_act(int l) const { return off_cpl(l) + cpl_w(); }
  __host__ __device__ long long off_aff_tail() const { return 2ll * d + (long long)Lc * cpl_theta(); }
  __host__ __device__ long long off_act_tail() const { return off_aff_tail() + 2ll * d; }
};

struct WideArgs {
  WideDims D;
  const float* theta;
  float* gtheta;          // NLL / SWEEP: gradient accumulator (module order)
  const float* x;         // NLL / PASS: input rows [*, d];   SWEEP: the pass OUTPUT rows y [n, d]
  const long long* rows;  // optional row indirection into x (NLL only)
  const float* gy;        // SWEEP: cotangent of the pass output [n, d]
  float* y;               // PASS: output rows [n, d]
  float* ld;              // PASS: log|det| of the pass direction [n]
  float* logp;            // PASS (optional): log q at the data end of the pass [n] -- x -> z: log N(z) + log|det dz/dx|; z -> x: log N(z) - log|det dx/dz|
  double* loss;           // NLL: += sum_i -log q(x_i)
  float* gx;              // optional (NLL / SWEEP): cotangent that reaches the pass input [n, d]
  long long n;
  int mode;               // 0 NLL, 1 PASS, 2 SWEEP
  int inv;                // direction of the pass: 0 = x -> z, 1 = z -> x
  const float* theta_t;   // SWEEP (optional): a copy of theta with every linear's weight transposed -- the conditioner's forward
                          // GEMMs read it (coalesced), the dgrad GEMMs keep `theta` (whose order is the coalesced one for them)
  int wt;                 // PASS only: every linear's weight is stored TRANSPOSED in theta ([in][out] instead of the modules'
                          // [out][in]; same offsets): the forward GEMMs' lanes run over output units, so only this layout is
                          // read coalesced (module order costs 32 cache lines per weight load: 9 % of the fp32 peak)
};

enum { kWideNll = 0, kWidePass = 1, kWideSweep = 2 };

__host__ __device__ inline int r4(int v) { return (v + 3) & ~3; }

// shared-memory plan (floats)
struct WidePlan {
  int ldv, lds, ldh, ldu;
  int oV, oG, oS, oAct, oU, oD0, oD1, oLd, total;
};
__host__ __device__ inline WidePlan wide_plan(const WideDims& D, int R, bool need_grad) {
  WidePlan P;
  P.ldv = r4(D.d); P.lds = r4(D.da); P.ldh = r4(D.H); P.ldu = r4(2 * D.db);
  int o = 0;
  P.oV = o; o += R * P.ldv;
  P.oG = o; o += need_grad ? R * P.ldv : 0;
  P.oS = o; o += R * P.lds;
  P.oAct = o; o += (D.M > 1 ? (D.M - 1) : 0) * R * P.ldh;
  P.oU = o; o += R * P.ldu;
  P.oD0 = o; o += (need_grad && D.M > 1) ? R * P.ldh : 0;
  P.oD1 = o; o += (need_grad && D.M > 2) ? R * P.ldh : 0;
  P.oLd = o; o += 2 * R;
  P.total = o;
  return P;
}

__device__ __forceinline__ float wide_alpha(float ua) { return expf(kLogOneMinusM + 0.5f * ua) + kMinScale; }

// out[r][j] = f( bias[j] + sum_q in[r][q] * W[j * sj + q * sq] ),  r < R, j < NJ, q < NQ.  `in` rows are zero padded to a
// multiple of 4 floats.  forward: j = output unit, q = input unit (sj = K, sq = 1); dgrad: j = input unit, q = output unit
// (sj = 1, sq = K, no bias).  MUL: out[r][j] = acc * (1 - mul[r][j]^2)  (tanh' of the activation that produced `in`'s input).
template <int R, bool TANH, bool MUL>
__device__ __forceinline__ void wide_gemm(const float* __restrict__ W, int sj, int sq, const float* __restrict__ bias, int NJ, int NQ,
                                          const float* in, int ldi, float* out, int ldo, const float* mul, int ldm) {
  const int NJp = (NJ + 31) & ~31;
  const int items = NJp * (R / kWRT);
  for (int it = threadIdx.x; it < items; it += kWT) {
    const int j = it % NJp, rb = it / NJp;
    if (j >= NJ) continue;
    float acc[kWRT];
    const float b = bias ? __ldg(bias + j) : 0.f;
#pragma unroll
    for (int r = 0; r < kWRT; ++r) acc[r] = b;
    const float* w = W + (size_t)j * sj;
    const float* ip = in + (size_t)rb * kWRT * ldi;
    int q = 0;
    for (; q + 4 <= NQ; q += 4) {
      const float w0 = __ldg(w + (size_t)q * sq), w1 = __ldg(w + (size_t)(q + 1) * sq);
      const float w2 = __ldg(w + (size_t)(q + 2) * sq), w3 = __ldg(w + (size_t)(q + 3) * sq);
#pragma unroll
      for (int r = 0; r < kWRT; ++r) {
        const float4 v = *reinterpret_cast<const float4*>(ip + r * ldi + q);
        acc[r] = fmaf(v.x, w0, fmaf(v.y, w1, fmaf(v.z, w2, fmaf(v.w, w3, acc[r]))));
      }
    }
    for (; q < NQ; ++q) {
      const float w0 = __ldg(w + (size_t)q * sq);
#pragma unroll
      for (int r = 0; r < kWRT; ++r) acc[r] = fmaf(ip[r * ldi + q], w0, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < kWRT; ++r) {
      const int row = rb * kWRT + r;
      float v = TANH ? tanhf(acc[r]) : acc[r];
      if (MUL) { const float a = mul[row * ldm + j]; v *= (1.f - a * a); }
      out[row * ldo + j] = v;
    }
  }
}

// Forward GEMM on TRANSPOSED weights WT[q][j] (WideArgs::wt) when NJ % 4 == 0 and WT is 16-byte aligned: a thread owns 4 rows x 4
// consecutive output units, so one 128-bit weight load (lanes on consecutive 16-byte words: fully coalesced) and one 128-bit
// broadcast load of the input tile feed 16 FMAs each -- 8 memory instructions per 64 FMAs against 12 per 32 in wide_gemm, which is
// bound by the load/store unit.  Same accumulation order as wide_gemm (bit-identical results).
constexpr int kWRB = 4;
template <int R, bool TANH>
__device__ __forceinline__ void wide_gemm_t4(const float* __restrict__ WT, const float* __restrict__ bias, int NJ, int NQ,
                                             const float* in, int ldi, float* out, int ldo) {
  const int NJ4 = NJ >> 2;
  const int items = NJ4 * (R / kWRB);
  for (int it = threadIdx.x; it < items; it += kWT) {
    const int j = (it % NJ4) << 2, rb = it / NJ4;
    float acc[kWRB][4];
    {
      const float b0 = bias ? __ldg(bias + j) : 0.f, b1 = bias ? __ldg(bias + j + 1) : 0.f;
      const float b2 = bias ? __ldg(bias + j + 2) : 0.f, b3 = bias ? __ldg(bias + j + 3) : 0.f;
#pragma unroll
      for (int r = 0; r < kWRB; ++r) { acc[r][0] = b0; acc[r][1] = b1; acc[r][2] = b2; acc[r][3] = b3; }
    }
    const float* ip = in + (size_t)rb * kWRB * ldi;
    const float* w = WT + j;
    int q = 0;
    for (; q + 4 <= NQ; q += 4) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + (size_t)q * NJ));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + (size_t)(q + 1) * NJ));
      const float4 w2 = __ldg(reinterpret_cast<const float4*>(w + (size_t)(q + 2) * NJ));
      const float4 w3 = __ldg(reinterpret_cast<const float4*>(w + (size_t)(q + 3) * NJ));
#pragma unroll
      for (int r = 0; r < kWRB; ++r) {
        const float4 v = *reinterpret_cast<const float4*>(ip + r * ldi + q);
        acc[r][0] = fmaf(v.x, w0.x, fmaf(v.y, w1.x, fmaf(v.z, w2.x, fmaf(v.w, w3.x, acc[r][0]))));
        acc[r][1] = fmaf(v.x, w0.y, fmaf(v.y, w1.y, fmaf(v.z, w2.y, fmaf(v.w, w3.y, acc[r][1]))));
        acc[r][2] = fmaf(v.x, w0.z, fmaf(v.y, w1.z, fmaf(v.z, w2.z, fmaf(v.w, w3.z, acc[r][2]))));
        acc[r][3] = fmaf(v.x, w0.w, fmaf(v.y, w1.w, fmaf(v.z, w2.w, fmaf(v.w, w3.w, acc[r][3]))));
      }
    }
    for (; q < NQ; ++q) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + (size_t)q * NJ));
#pragma unroll
      for (int r = 0; r < kWRB; ++r) {
        const float v = ip[r * ldi + q];
        acc[r][0] = fmaf(v, w0.x, acc[r][0]); acc[r][1] = fmaf(v, w0.y, acc[r][1]);
        acc[r][2] = fmaf(v, w0.z, acc[r][2]); acc[r][3] = fmaf(v, w0.w, acc[r][3]);
      }
    }
#pragma unroll
    for (int r = 0; r < kWRB; ++r) {
      float4 o;
      o.x = TANH ? tanhf(acc[r][0]) : acc[r][0]; o.y = TANH ? tanhf(acc[r][1]) : acc[r][1];
      o.z = TANH ? tanhf(acc[r][2]) : acc[r][2]; o.w = TANH ? tanhf(acc[r][3]) : acc[r][3];
      *reinterpret_cast<float4*>(out + (rb * kWRB + r) * ldo + j) = o;
    }
  }
}

// dW[n][k] += sum_r dout[r][n] in[r][k];  db[n] += sum_r dout[r][n].   dout rows are zero padded to a multiple of 4.
template <int R>
__device__ __forceinline__ void wide_wgrad(float* __restrict__ gW, float* __restrict__ gb, int N, int K, const float* dout, int ldo,
                                           const float* in, int ldi) {
  const int Kp = (K + 31) & ~31;
  const int N4 = (N + 3) >> 2;
  const int items = Kp * N4;
  for (int it = threadIdx.x; it < items; it += kWT) {
    const int k = it % Kp, n0 = (it / Kp) * 4;
    if (k >= K) continue;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
    for (int r = 0; r < R; ++r) {
      const float v = in[r * ldi + k];
      const float4 dv = *reinterpret_cast<const float4*>(dout + r * ldo + n0);
      a0 = fmaf(dv.x, v, a0); a1 = fmaf(dv.y, v, a1); a2 = fmaf(dv.z, v, a2); a3 = fmaf(dv.w, v, a3);
    }
    float* g = gW + (size_t)n0 * K + k;
    if (a0 != 0.f) atomicAdd(g, a0);
    if (n0 + 1 < N && a1 != 0.f) atomicAdd(g + K, a1);
    if (n0 + 2 < N && a2 != 0.f) atomicAdd(g + 2 * (size_t)K, a2);
    if (n0 + 3 < N && a3 != 0.f) atomicAdd(g + 3 * (size_t)K, a3);
  }
  for (int n = threadIdx.x; n < N; n += kWT) {
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < R; ++r) s += dout[r * ldo + n];
    if (s != 0.f) atomicAdd(gb + n, s);
  }
}

template <int R>
struct WideTile {
  const WideDims& D;
  const WidePlan& P;
  float* sm;
  bool rev;          // logical index i lives in column d-1-i
  bool wt;           // the forward GEMMs read linear weights stored [in][out] (WideArgs::wt / WideArgs::theta_t)
  const float* fwd;  // SWEEP with WideArgs::theta_t: parameter vector the conditioner's forward GEMMs read instead of `theta`
  __device__ WideTile(const WideDims& D_, const WidePlan& P_, float* sm_) : D(D_), P(P_), sm(sm_), rev(false), wt(false), fwd(nullptr) {}
  __device__ __forceinline__ int col(int i) const { return rev ? D.d - 1 - i : i; }
  __device__ __forceinline__ float* V() const { return sm + P.oV; }
  __device__ __forceinline__ float* G() const { return sm + P.oG; }
  __device__ __forceinline__ float* S() const { return sm + P.oS; }
  __device__ __forceinline__ float* Act(int m) const { return sm + P.oAct + m * R * P.ldh; }
  __device__ __forceinline__ float* U() const { return sm + P.oU; }
  __device__ __forceinline__ float* Ld() const { return sm + P.oLd; }        // [R] log-det accumulators
  __device__ __forceinline__ float* Active() const { return sm + P.oLd + R; }  // [R] 1 = the row exists

  // ---- elementwise affine with parameters value[i] = (u_a, u_b) at theta + off ------------------------------------
  // pass direction: inv = 0: v <- alpha v + beta, ld += log alpha;  inv = 1: v <- (v - beta) / alpha, ld -= log alpha
  __device__ void affine_pass(const float* theta, long long off, bool inv) {
    const int d = D.d;
    float* v = V();
    for (int i = threadIdx.x; i < d; i += kWT) {
      const float al = wide_alpha(__ldg(theta + off + 2 * i)), be = 0.5f * __ldg(theta + off + 2 * i + 1);
      const float ra = 1.f / al;
      const int c = col(i);
#pragma unroll 4
      for (int r = 0; r < R; ++r) {
        const float x = v[r * P.ldv + c];
        v[r * P.ldv + c] = inv ? (x - be) * ra : fmaf(al, x, be);
      }
    }
    // the log-determinant of an elementwise affine is row independent: one warp sums it, every row receives it
    if (threadIdx.x < 32) {
      float s = 0.f;
      for (int i = threadIdx.x; i < d; i += 32) s += logf(wide_alpha(__ldg(theta + off + 2 * i)));
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (threadIdx.x < R) Ld()[threadIdx.x] += inv ? -s : s;
      if (R > 32 && threadIdx.x + 32 < R) Ld()[threadIdx.x + 32] += inv ? -s : s;
    }
    __syncthreads();
  }
  // backward through the same layer: (V, G) = pass output and its cotangent -> pass input and its cotangent; emits d theta
  __device__ void affine_back(const float* theta, long long off, bool inv, float* gtheta) {
    const int d = D.d;
    float* v = V();
    float* g = G();
    const float* actv = Active();
    for (int i = threadIdx.x; i < d; i += kWT) {
      const float al = wide_alpha(__ldg(theta + off + 2 * i)), be = 0.5f * __ldg(theta + off + 2 * i + 1);
      const float ra = 1.f / al;
      const int c = col(i);
      float dal = 0.f, dbe = 0.f;
#pragma unroll 4
      for (int r = 0; r < R; ++r) {
        const float y = v[r * P.ldv + c], gy = g[r * P.ldv + c], a = actv[r];
        if (!inv) {
          const float xin = (y - be) * ra;
          dal += gy * xin - a * ra;
          dbe += gy;
          v[r * P.ldv + c] = xin;
          g[r * P.ldv + c] = gy * al;
        } else {
          dal += (a - gy * y) * ra;
          dbe -= gy * ra;
          v[r * P.ldv + c] = fmaf(al, y, be);
          g[r * P.ldv + c] = gy * ra;
        }
      }
      if (gtheta) {
        atomicAdd(gtheta + off + 2 * i, dal * (al - kMinScale) * 0.5f);
        atomicAdd(gtheta + off + 2 * i + 1, 0.5f * dbe);
      }
    }
    __syncthreads();
  }

  // ---- conditioner MLP on the source half (logical [0, da)): S <- source, Act[m] <- tanh layers, U <- outputs ------
  __device__ void conditioner(const float* theta, long long off) {
    const int da = D.da;
    float* s = S();
    const float* v = V();
    for (int it = threadIdx.x; it < R * P.lds; it += kWT) {
      const int r = it / P.lds, k = it % P.lds;
      s[it] = k < da ? v[r * P.ldv + col(k)] : 0.f;
    }
    __syncthreads();
    const float* in = s;
    int ldi = P.lds;
    long long o = off;
    for (int m = 0; m < D.M; ++m) {
      const int K = D.lin_in(m), N = D.lin_out(m);
      const float* W = (fwd ? fwd : theta) + o;      // fwd: the transposed copy, forward GEMMs only (biases are the same in both)
      const float* b = W + (size_t)N * K;
      if (m < D.M - 1) {
        float* out = Act(m);
        // pad columns of the activation rows must be zero: they are read 4 at a time by the next layer
        if ((D.H & 3) != 0)
          for (int it = threadIdx.x; it < R * (P.ldh - D.H); it += kWT) out[(it / (P.ldh - D.H)) * P.ldh + D.H + it % (P.ldh - D.H)] = 0.f;
        if (wt && (N & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0) wide_gemm_t4<R, true>(W, b, N, K, in, ldi, out, P.ldh);
        else wide_gemm<R, true, false>(W, wt ? 1 : K, wt ? N : 1, b, N, K, in, ldi, out, P.ldh, nullptr, 0);
        in = out; ldi = P.ldh;
      } else {
        if (wt && (N & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0) wide_gemm_t4<R, false>(W, b, N, K, in, ldi, U(), P.ldu);
        else wide_gemm<R, false, false>(W, wt ? 1 : K, wt ? N : 1, b, N, K, in, ldi, U(), P.ldu, nullptr, 0);
      }
      o += (long long)N * K + N;
      __syncthreads();
    }
  }
  // coupling, pass direction: target t (logical da + t) <- alpha_t b + beta_t  /  (b - beta_t) / alpha_t
  __device__ void coupling_pass(const float* theta, long long off, bool inv) {
    conditioner(theta, off);
    float* v = V();
    const float* u = U();
    float* ld = Ld();
    for (int it = threadIdx.x; it < R * 32; it += kWT) {      // one warp-row per tile row: lanes over the targets
      const int r = it >> 5, lane = it & 31;
      float s = 0.f;
      for (int t = lane; t < D.db; t += 32) {
        const float al = wide_alpha(u[r * P.ldu + 2 * t]), be = 0.5f * u[r * P.ldu + 2 * t + 1];
        const int c = col(D.da + t);
        const float b = v[r * P.ldv + c];
        v[r * P.ldv + c] = inv ? (b - be) / al : fmaf(al, b, be);
        s += logf(al);
      }
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) ld[r] += inv ? -s : s;
    }
    __syncthreads();
  }
  // coupling, backward: re-evaluate the conditioner on the (unchanged) source half, undo the target map, turn U into
  // dL/dU, then walk the MLP back (wgrad + dgrad); the source's cotangent receives the input-VJP.
  __device__ void coupling_back(const float* theta, long long off, bool inv, float* gtheta) {
    conditioner(theta, off);
    float* v = V();
    float* g = G();
    float* u = U();
    const float* actv = Active();
    for (int it = threadIdx.x; it < R * (P.ldu >> 1); it += kWT) {
      const int r = it / (P.ldu >> 1), t = it % (P.ldu >> 1);
      float d_ua = 0.f, d_ub = 0.f;
      if (t < D.db) {
        const float al = wide_alpha(u[r * P.ldu + 2 * t]), be = 0.5f * u[r * P.ldu + 2 * t + 1];
        const float ra = 1.f / al, a = actv[r];
        const int c = col(D.da + t);
        const float y = v[r * P.ldv + c], gy = g[r * P.ldv + c];
        float dal, dbe;
        if (!inv) {
          const float xin = (y - be) * ra;
          dal = gy * xin - a * ra; dbe = gy;
          v[r * P.ldv + c] = xin; g[r * P.ldv + c] = gy * al;
        } else {
          dal = (a - gy * y) * ra; dbe = -gy * ra;
          v[r * P.ldv + c] = fmaf(al, y, be); g[r * P.ldv + c] = gy * ra;
        }
        d_ua = dal * (al - kMinScale) * 0.5f;
        d_ub = 0.5f * dbe;
      }
      u[r * P.ldu + 2 * t] = d_ua;          // pad pairs get zeros
      u[r * P.ldu + 2 * t + 1] = d_ub;
    }
    __syncthreads();
    // offsets of the linear layers
    long long offs[16];
    {
      long long o = off;
      for (int m = 0; m < D.M; ++m) { offs[m] = o; o += (long long)D.lin_out(m) * D.lin_in(m) + D.lin_out(m); }
    }
    const float* dout = u;
    int ldo = P.ldu;
    float* dbuf[2] = {sm + P.oD0, sm + P.oD1};
    int which = 0;
    for (int m = D.M - 1; m >= 0; --m) {
      const int K = D.lin_in(m), N = D.lin_out(m);
      const float* in = m == 0 ? S() : Act(m - 1);
      const int ldi = m == 0 ? P.lds : P.ldh;
      if (gtheta) wide_wgrad<R>(gtheta + offs[m], gtheta + offs[m] + (size_t)N * K, N, K, dout, ldo, in, ldi);
      const float* W = theta + offs[m];
      if (m > 0) {
        float* din = dbuf[which];
        if ((D.H & 3) != 0)
          for (int it = threadIdx.x; it < R * (P.ldh - D.H); it += kWT) din[(it / (P.ldh - D.H)) * P.ldh + D.H + it % (P.ldh - D.H)] = 0.f;
        // din[r][k] = (sum_n dout[r][n] W[n][k]) * (1 - act[m-1][r][k]^2)
        wide_gemm<R, false, true>(W, 1, K, nullptr, K, N, dout, ldo, din, P.ldh, Act(m - 1), P.ldh);
        __syncthreads();
        dout = din; ldo = P.ldh; which ^= 1;
      } else {
        // source cotangent: reuse S as the output buffer is not possible (wgrad of layer 0 reads it concurrently): the
        // result goes through the free hidden-cotangent buffer or, for M == 1, straight into G
        __syncthreads();
        const int NJp = (K + 31) & ~31;
        for (int it = threadIdx.x; it < NJp * (R / kWRT); it += kWT) {
          const int k = it % NJp, rb = it / NJp;
          if (k >= K) continue;
          float acc[kWRT];
#pragma unroll
          for (int r = 0; r < kWRT; ++r) acc[r] = 0.f;
          const float* ip = dout + (size_t)rb * kWRT * ldo;
          int n = 0;
          for (; n + 4 <= N; n += 4) {
            const float w0 = __ldg(W + (size_t)n * K + k), w1 = __ldg(W + (size_t)(n + 1) * K + k);
            const float w2 = __ldg(W + (size_t)(n + 2) * K + k), w3 = __ldg(W + (size_t)(n + 3) * K + k);
#pragma unroll
            for (int r = 0; r < kWRT; ++r) {
              const float4 dv = *reinterpret_cast<const float4*>(ip + r * ldo + n);
              acc[r] = fmaf(dv.x, w0, fmaf(dv.y, w1, fmaf(dv.z, w2, fmaf(dv.w, w3, acc[r]))));
            }
          }
          for (; n < N; ++n) {
            const float w0 = __ldg(W + (size_t)n * K + k);
#pragma unroll
            for (int r = 0; r < kWRT; ++r) acc[r] = fmaf(ip[r * ldo + n], w0, acc[r]);
          }
          const int c = col(k);
#pragma unroll
          for (int r = 0; r < kWRT; ++r) g[(rb * kWRT + r) * P.ldv + c] += acc[r];
        }
        __syncthreads();
      }
    }
  }
};

template <int R, int MINB>
__global__ void __launch_bounds__(kWT, MINB) flow_train_wide_kernel(const WideArgs A) {
  extern __shared__ __align__(16) float wsm[];
  const WideDims& D = A.D;
  const bool need_grad = A.mode != kWidePass;
  const WidePlan P = wide_plan(D, R, need_grad);
  const int d = D.d;
  const long long tiles = (A.n + R - 1) / R;
  double loss = 0.0;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    WideTile<R> T(D, P, wsm);
    T.wt = (A.mode == kWidePass && A.wt != 0) || (A.mode == kWideSweep && A.theta_t != nullptr);
    T.fwd = A.mode == kWideSweep ? A.theta_t : nullptr;
    const long long row0 = tile * R;
    // ---- load (global rows are always in LOGICAL order; a tile that starts at the latent end of a flow with an odd number
    //      of reversals is seated reversed) ---------------------------------------------------------------------------------
    const bool pass_inv = A.inv != 0;
    const bool latent_in = (A.mode == kWidePass && pass_inv) || (A.mode == kWideSweep && !pass_inv);
    const bool rev_in = latent_in && (D.Lc & 1) != 0;
    for (int it = threadIdx.x; it < R * P.ldv; it += kWT) {
      const int r = it / P.ldv, c = it % P.ldv;
      const long long row = row0 + r;
      float v = 0.f, gv = 0.f;
      if (row < A.n && c < d) {
        const long long src = (A.mode == kWideNll && A.rows) ? A.rows[row] : row;
        const int i = rev_in ? d - 1 - c : c;
        v = __ldg(A.x + src * d + i);
        if (A.mode == kWideSweep) gv = __ldg(A.gy + row * d + i);
      }
      T.V()[it] = v;
      if (need_grad) T.G()[it] = gv;
    }
    for (int r = threadIdx.x; r < R; r += kWT) {
      T.Ld()[r] = 0.f;
      T.Active()[r] = (row0 + r < A.n) ? 1.f : 0.f;
    }
    __syncthreads();
    if (A.mode == kWidePass && pass_inv && A.logp) {
      // z -> x with log q(x) wanted (Flow.sample(return_log_prob=True)): log N(z; 0, I) now, while the tile still holds z;
      // log|det dx/dz| is subtracted when the pass has finished
      for (int it = threadIdx.x; it < R * 32; it += kWT) {
        const int r = it >> 5, lane = it & 31;
        float s = 0.f;
        for (int c = lane; c < d; c += 32) { const float z = T.V()[r * P.ldv + c]; s = fmaf(z, z, s); }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && row0 + r < A.n) A.logp[row0 + r] = -0.5f * s - 0.5f * (float)d * 1.8378770664093453f;
      }
      __syncthreads();
    }
    // ---- the pass (NLL: x -> z;  PASS: either direction) ---------------------------------------------------------------
    if (A.mode != kWideSweep) {
      if (!pass_inv) {
        T.rev = false;
        T.affine_pass(A.theta, 0, false);
        for (int l = 0; l < D.Lc; ++l) {
          T.rev = !T.rev;
          T.coupling_pass(A.theta, D.off_cpl(l), false);
          T.affine_pass(A.theta, D.off_act(l), false);
        }
        T.affine_pass(A.theta, D.off_aff_tail(), false);
        T.affine_pass(A.theta, D.off_act_tail(), false);
      } else {
        T.rev = (D.Lc & 1) != 0;
        T.affine_pass(A.theta, D.off_act_tail(), true);
        T.affine_pass(A.theta, D.off_aff_tail(), true);
        for (int l = D.Lc - 1; l >= 0; --l) {
          T.affine_pass(A.theta, D.off_act(l), true);
          T.coupling_pass(A.theta, D.off_cpl(l), true);
          T.rev = !T.rev;
        }
        T.affine_pass(A.theta, 0, true);
      }
    }
    if (A.mode == kWidePass) {
      // NOTE on columns: the latent z of a flow with an odd number of reversals sits reversed in the tile; stores and
      // loads always use LOGICAL indices, so nothing is flipped in global memory.
      const bool rev_out = pass_inv ? false : ((D.Lc & 1) != 0);
      if (A.y)
        for (int it = threadIdx.x; it < R * d; it += kWT) {
          const int r = it / d, i = it % d;
          if (row0 + r < A.n) A.y[(row0 + r) * d + i] = T.V()[r * P.ldv + (rev_out ? d - 1 - i : i)];
        }
      if (A.ld)
        for (int r = threadIdx.x; r < R; r += kWT)
          if (row0 + r < A.n) A.ld[row0 + r] = T.Ld()[r];
      if (A.logp && pass_inv)
        for (int r = threadIdx.x; r < R; r += kWT)
          if (row0 + r < A.n) A.logp[row0 + r] -= T.Ld()[r];
      if (A.logp && !pass_inv)
        for (int it = threadIdx.x; it < R * 32; it += kWT) {
          const int r = it >> 5, lane = it & 31;
          float s = 0.f;
          for (int c = lane; c < d; c += 32) { const float z = T.V()[r * P.ldv + c]; s = fmaf(z, z, s); }
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          if (lane == 0 && row0 + r < A.n) A.logp[row0 + r] = T.Ld()[r] - 0.5f * s - 0.5f * (float)d * 1.8378770664093453f;
        }
      __syncthreads();
      continue;
    }
    bool sweep_inv;
    if (A.mode == kWideNll) {
      // loss_i = |z|^2 / 2 + d/2 log 2 pi - ld;  cotangent dL/dz = z (inactive rows: 0)
      for (int it = threadIdx.x; it < R * 32; it += kWT) {
        const int r = it >> 5, lane = it & 31;
        float s = 0.f;
        for (int c = lane; c < d; c += 32) { const float z = T.V()[r * P.ldv + c]; s = fmaf(z, z, s); }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && row0 + r < A.n) loss += (double)(0.5f * s + 0.5f * (float)d * 1.8378770664093453f - T.Ld()[r]);
      }
      for (int it = threadIdx.x; it < R * P.ldv; it += kWT) {
        const int r = it / P.ldv;
        T.G()[it] = T.Active()[r] * T.V()[it];
      }
      __syncthreads();
      sweep_inv = false;
    } else {
      sweep_inv = pass_inv;
      T.rev = rev_in;
    }
    // ---- backward sweep: undo the pass layer by layer -----------------------------------------------------------------
    if (!sweep_inv) {
      T.affine_back(A.theta, D.off_act_tail(), false, A.gtheta);
      T.affine_back(A.theta, D.off_aff_tail(), false, A.gtheta);
      for (int l = D.Lc - 1; l >= 0; --l) {
        T.affine_back(A.theta, D.off_act(l), false, A.gtheta);
        T.coupling_back(A.theta, D.off_cpl(l), false, A.gtheta);
        T.rev = !T.rev;
      }
      T.affine_back(A.theta, 0, false, A.gtheta);
    } else {
      T.affine_back(A.theta, 0, true, A.gtheta);
      for (int l = 0; l < D.Lc; ++l) {
        T.rev = !T.rev;
        T.coupling_back(A.theta, D.off_cpl(l), true, A.gtheta);
        T.affine_back(A.theta, D.off_act(l), true, A.gtheta);
      }
      T.affine_back(A.theta, D.off_aff_tail(), true, A.gtheta);
      T.affine_back(A.theta, D.off_act_tail(), true, A.gtheta);
    }
    if (A.gx) {
      for (int it = threadIdx.x; it < R * d; it += kWT) {
        const int r = it / d, i = it % d;
        if (row0 + r < A.n) A.gx[(row0 + r) * d + i] = T.G()[r * P.ldv + T.col(i)];
      }
    }
    __syncthreads();
  }
  if (A.mode == kWideNll && A.loss) {
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if ((threadIdx.x & 31) == 0 && loss != 0.0) atomicAdd(A.loss, loss);
  }
}

namespace {

int wide_check(int d, int Lc, int M, int H) {
  if (d < 2 || d > NFMC_MAX_DIM || Lc < 0) return set_error("wide flow training: bad d / n_coupling");
  if (M < 1 || M > 16 || H < 1 || H > 4096) return set_error("wide flow training: n_linear must be in [1, 16], hidden in [1, 4096]");
  return 0;
}

// MINB = resident CTAs per SM the kernel is compiled for: 2 (128 registers, no spills) for the passes of the sampling path --
// full batches, where a second CTA hides the weight loads' latency (ncu with one CTA: 2 warps per scheduler, long_scoreboard
// 1.7 per issue; deep-flow pass 7.2 -> 5.0 ms) --, 1 (182 registers) for the training modes, whose minibatches are fewer
// tiles than SMs and run 7-12 % slower under the register cap.
template <int R, int MINB>
int wide_launch_rm(const WideArgs& A, size_t smem, cudaStream_t s) {
  if (int e = check_cuda(cudaFuncSetAttribute(flow_train_wide_kernel<R, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "wide smem attribute")) return e;
  const long long tiles = (A.n + R - 1) / R;
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, flow_train_wide_kernel<R, MINB>, kWT, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  if (per_sm > MINB) per_sm = MINB;
  const int grid = (int)std::min<long long>(tiles, (long long)per_sm * sm_count());
  flow_train_wide_kernel<R, MINB><<<grid, kWT, smem, s>>>(A);
  return check_cuda(cudaGetLastError(), "flow_train_wide_kernel launch");
}
template <int R>
int wide_launch_r(const WideArgs& A, size_t smem, cudaStream_t s) {
  // the sampling-side sweep (NeuTra's latent gradient: cotangent to the pass input only, full batches) shares the two-CTA build
  const bool sampling = A.mode == kWidePass || (A.mode == kWideSweep && !A.gtheta);
  return sampling ? wide_launch_rm<R, 2>(A, smem, s) : wide_launch_rm<R, 1>(A, smem, s);
}

int wide_launch(WideArgs& A, cudaStream_t s) {
  // rows per tile: 32 when the batch fills the machine at that size (fewer weight reads and atomics per row), else 16 / 8
  // (more CTAs for a latency-bound minibatch); always what fits in 227 KB of shared memory
  const bool need_grad = A.mode != kWidePass;
  const size_t cap = 227 * 1024;
  int R = (A.n >= 32ll * sm_count()) ? 32 : (A.n >= 16ll * sm_count() / 2 ? 16 : 8);
  // NeuTra's sweep (cotangent to the input only): 16-row tiles, so that two CTAs fit an SM next to the gradient buffers (+6 %)
  if (A.mode == kWideSweep && !A.gtheta && R == 32) R = 16;
  if (const char* e = getenv("NFMC_WIDE_ROWS")) { const int v = atoi(e); if (v == 8 || v == 16 || v == 32) R = v; }
  while (R > 8 && (size_t)wide_plan(A.D, R, need_grad).total * sizeof(float) > cap) R >>= 1;
  // sampling-side launches are compiled for two CTAs per SM: prefer a tile two of which fit (d = 1000, H = 64: 8-row tiles
  // 3.8 ms per pass against 4.9 ms with one 16-row CTA per SM)
  if (A.mode == kWidePass || (A.mode == kWideSweep && !A.gtheta))
    while (R > 8 && (size_t)wide_plan(A.D, R, need_grad).total * sizeof(float) > cap / 2) R >>= 1;
  const size_t smem = (size_t)wide_plan(A.D, R, need_grad).total * sizeof(float);
  if (smem > cap) return set_error("wide flow training: the row tile does not fit in shared memory (d x hidden too large)");
  if (R == 32) return wide_launch_r<32>(A, smem, s);
  if (R == 16) return wide_launch_r<16>(A, smem, s);
  return wide_launch_r<8>(A, smem, s);
}

__global__ void adamw_wide_kernel(float* __restrict__ theta, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                                  long long n, float gscale, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = grad[i] * gscale;
    float p = theta[i] * (1.f - lr * wd);
    const float mi = fmaf(b1, m[i], (1.f - b1) * g);
    const float vi = fmaf(b2, v[i], (1.f - b2) * g * g);
    m[i] = mi; v[i] = vi;
    p -= (lr / bc1) * (mi / (sqrtf(vi) / bc2_sqrt + eps));
    theta[i] = p;
  }
}

}  // namespace
}  // namespace nfmc

using namespace nfmc;

static WideDims wide_dims(int d, int Lc, int M, int H) {
  WideDims D;
  D.d = d; D.da = d / 2; D.db = d - d / 2; D.Lc = Lc; D.M = M; D.H = H;
  return D;
}

extern "C" int64_t nfmc_flow_wide_param_count(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden) {
  if (wide_check(d, n_coupling, n_linear, hidden)) return -1;
  return wide_dims(d, n_coupling, n_linear, hidden).n_theta();
}

extern "C" int nfmc_flow_wide_nll_grad(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden, const float* theta,
                                       const float* x, const int64_t* rows, int64_t n, float* grad_theta, double* loss,
                                       float* grad_x, int32_t accumulate, void* stream) {
  if (int e = wide_check(d, n_coupling, n_linear, hidden)) return e;
  if (!theta || !x || !grad_theta || n < 1) return set_error("flow_wide_nll_grad: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  WideArgs A{};
  A.D = wide_dims(d, n_coupling, n_linear, hidden);
  if (!accumulate) {
    if (int e = check_cuda(cudaMemsetAsync(grad_theta, 0, (size_t)A.D.n_theta() * sizeof(float), s), "zero grad")) return e;
    if (loss) if (int e = check_cuda(cudaMemsetAsync(loss, 0, sizeof(double), s), "zero loss")) return e;
  }
  A.theta = theta; A.gtheta = grad_theta; A.x = x; A.rows = reinterpret_cast<const long long*>(rows); A.n = n;
  A.loss = loss; A.gx = grad_x; A.mode = kWideNll; A.inv = 0;
  return wide_launch(A, s);
}

extern "C" int nfmc_flow_wide_pass(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden, const float* theta,
                                   int32_t inverse, const float* in, float* out, float* log_det, int64_t n, void* stream) {
  if (int e = wide_check(d, n_coupling, n_linear, hidden)) return e;
  if (!theta || !in || !out || n < 1) return set_error("flow_wide_pass: bad arguments");
  WideArgs A{};
  A.D = wide_dims(d, n_coupling, n_linear, hidden);
  A.theta = theta; A.x = in; A.y = out; A.ld = log_det; A.n = n; A.mode = kWidePass; A.inv = (inverse & 1) ? 1 : 0;
  A.wt = (inverse & 2) ? 1 : 0;
  return wide_launch(A, (cudaStream_t)stream);
}

extern "C" int nfmc_flow_wide_log_prob(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden, const float* theta,
                                       int32_t transposed, const float* x, float* log_q, int64_t n, void* stream) {
  if (int e = wide_check(d, n_coupling, n_linear, hidden)) return e;
  if (!theta || !x || !log_q || n < 1) return set_error("flow_wide_log_prob: bad arguments");
  WideArgs A{};
  A.D = wide_dims(d, n_coupling, n_linear, hidden);
  A.theta = theta; A.x = x; A.logp = log_q; A.n = n; A.mode = kWidePass; A.inv = 0; A.wt = transposed ? 1 : 0;
  return wide_launch(A, (cudaStream_t)stream);
}

extern "C" int nfmc_flow_wide_sample(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden, const float* theta,
                                     int32_t transposed, const nfmc_rng* rng, int64_t chain0, float* x, float* log_q, int64_t n,
                                     void* stream) {
  if (int e = wide_check(d, n_coupling, n_linear, hidden)) return e;
  if (!theta || !rng || !x || n < 1) return set_error("flow_wide_sample: bad arguments");
  const float* z = rng->normals;
  if (!z) {                    // base draw: Philox stream 1, the numbers nfmc_flow_sample and the jump kernels draw
    nfmc_rng r2{rng->seed, rng->step0, nullptr, nullptr};
    if (int e = nfmc_rng_fill(&r2, 1, chain0, d, n, 1, x, nullptr, stream)) return e;
    z = x;                     // in place: a CTA reads its rows into shared memory before it writes them
  }
  WideArgs A{};
  A.D = wide_dims(d, n_coupling, n_linear, hidden);
  A.theta = theta; A.x = z; A.y = x; A.logp = log_q; A.n = n; A.mode = kWidePass; A.inv = 1; A.wt = transposed ? 1 : 0;
  return wide_launch(A, (cudaStream_t)stream);
}

extern "C" int nfmc_flow_wide_sweep(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden, const float* theta,
                                    int32_t inverse, const float* y, const float* grad_y, int64_t n, float* grad_theta,
                                    float* grad_in, int32_t accumulate, void* stream) {
  if (int e = wide_check(d, n_coupling, n_linear, hidden)) return e;
  if (!theta || !y || !grad_y || n < 1 || (!grad_theta && !grad_in)) return set_error("flow_wide_sweep: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  WideArgs A{};
  A.D = wide_dims(d, n_coupling, n_linear, hidden);
  if (grad_theta && !accumulate)
    if (int e = check_cuda(cudaMemsetAsync(grad_theta, 0, (size_t)A.D.n_theta() * sizeof(float), s), "zero grad")) return e;
  A.theta = theta; A.gtheta = grad_theta; A.x = y; A.gy = grad_y; A.gx = grad_in; A.n = n; A.mode = kWideSweep; A.inv = inverse ? 1 : 0;
  return wide_launch(A, s);
}

extern "C" int nfmc_flow_wide_pullback(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden, const float* theta,
                                       const float* theta_t, int32_t inverse, const float* y, const float* grad_y, int64_t n,
                                       float* grad_in, void* stream) {
  if (int e = wide_check(d, n_coupling, n_linear, hidden)) return e;
  if (!theta || !y || !grad_y || !grad_in || n < 1) return set_error("flow_wide_pullback: bad arguments");
  WideArgs A{};
  A.D = wide_dims(d, n_coupling, n_linear, hidden);
  A.theta = theta; A.theta_t = theta_t; A.x = y; A.gy = grad_y; A.gx = grad_in; A.n = n; A.mode = kWideSweep; A.inv = inverse ? 1 : 0;
  return wide_launch(A, (cudaStream_t)stream);
}

extern "C" int nfmc_adamw_step_scaled(float* theta, const float* grad, float grad_scale, float* exp_avg, float* exp_avg_sq,
                                      int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                                      void* stream) {
  if (!theta || !grad || !exp_avg || !exp_avg_sq || n < 1 || step < 1) return set_error("adamw_step_scaled: bad arguments");
  const float bc1 = 1.f - (float)pow((double)beta1, (double)step);
  const float bc2 = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  const int grid = (int)std::min<long long>((n + 255) / 256, 4 * sm_count());
  adamw_wide_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(theta, grad, exp_avg, exp_avg_sq, n, grad_scale, lr, beta1, beta2, eps,
                                                             weight_decay, bc1, bc2);
  return check_cuda(cudaGetLastError(), "adamw_wide_kernel launch");
}

// One epoch of minibatch maximum likelihood on one GPU for a wide / deep flow: per batch of perm[n]: loss + gradient in
// module order, AdamW with the 1/batch factor folded in.  losses[b] = summed loss of batch b.
extern "C" int nfmc_flow_wide_fit_epoch(int32_t d, int32_t n_coupling, int32_t n_linear, int32_t hidden, float* theta,
                                        float* exp_avg, float* exp_avg_sq, float* grad_theta, double* losses, const float* x,
                                        const int64_t* perm, int64_t n, int64_t batch_size, float lr, float beta1, float beta2,
                                        float eps, float weight_decay, int32_t step0, void* stream) {
  if (int e = wide_check(d, n_coupling, n_linear, hidden)) return e;
  if (!perm || n < 1 || batch_size < 1) return set_error("flow_wide_fit_epoch: bad perm / n / batch_size");
  const int64_t P = wide_dims(d, n_coupling, n_linear, hidden).n_theta();
  int32_t step = step0;
  int64_t b = 0;
  for (int64_t i = 0; i < n; i += batch_size, ++b) {
    const int64_t m = std::min<int64_t>(batch_size, n - i);
    if (int e = nfmc_flow_wide_nll_grad(d, n_coupling, n_linear, hidden, theta, x, perm + i, m, grad_theta, losses ? losses + b : nullptr,
                                        nullptr, 0, stream)) return e;
    if (int e = nfmc_adamw_step_scaled(theta, grad_theta, 1.f / (float)m, exp_avg, exp_avg_sq, P, lr, beta1, beta2, eps, weight_decay,
                                       ++step, stream)) return e;
  }
  return 0;
}
