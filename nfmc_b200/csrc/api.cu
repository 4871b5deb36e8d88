// api.cu -- error plumbing, device queries and the host-buffer end-to-end entry point of the C ABI.
#include <cstring>
#include <string>
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "host_common.cuh"
#include "flow.cuh"

namespace nfmc {

static thread_local std::string g_last_error;

int set_error(const std::string& msg) {
  g_last_error = msg;
  return 1;
}
int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  return set_error(std::string(what) + ": " + cudaGetErrorString(e));
}
int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace nfmc

using namespace nfmc;

extern "C" const char* nfmc_last_error(void) { return g_last_error.c_str(); }
extern "C" int nfmc_abi_version(void) { return NFMC_ABI_VERSION; }

static inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

extern "C" int64_t nfmc_jump_workspace_bytes(int32_t d, int64_t n, int64_t blob_floats) {
  return (int64_t)(align256((size_t)n * d * sizeof(float)) + align256((size_t)blob_floats * sizeof(float)) +
                   align256((size_t)2 * d * sizeof(float)) + align256((size_t)2 * d * sizeof(double)) + align256(8 * sizeof(unsigned long long)) +
                   align256((size_t)n * sizeof(float)));
}

// Slab pipeline shared by the two whole-run entry points.  The chain batch is cut into slabs that are spread over four
// streams.  With host buffers, slab i's H2D copy, its kernels and its D2H copy overlap the neighbours', so the PCIe
// transfers (2 x 4*n*d bytes) hide behind the compute.  With device-resident chains the point is the kernels themselves:
// CTAs of the latency-bound jump kernel of one slab run beside CTAs of the issue-bound local kernel of another and fill
// issue slots those leave idle (measured at d = 100, 2^20 chains: 20.6 -> 19.4 ms per outer iteration).  Chains keep
// their global index (chain0 + row), so the result is identical to one big launch per kernel.
namespace {
constexpr int kPipeStreams = 4;
struct Pipe {
  cudaStream_t streams[kPipeStreams] = {};
  cudaEvent_t fork = nullptr, join[kPipeStreams] = {};
  bool ready = false;
};
constexpr int kMaxDevices = 64;
thread_local Pipe g_pipes[kMaxDevices];          // one set of side streams per device (and host thread), created once
int ensure_pipe(Pipe*& out) {
  int dev = 0;
  if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
  if (dev < 0 || dev >= kMaxDevices) return set_error("jump_sample: device ordinal out of range");
  Pipe& p = g_pipes[dev];
  if (!p.ready) {
    for (int i = 0; i < kPipeStreams; ++i) {
      if (int e = check_cuda(cudaStreamCreateWithFlags(&p.streams[i], cudaStreamNonBlocking), "stream create")) return e;
      if (int e = check_cuda(cudaEventCreateWithFlags(&p.join[i], cudaEventDisableTiming), "event create")) return e;
    }
    if (int e = check_cuda(cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming), "event create")) return e;
    p.ready = true;
  }
  out = &p;
  return 0;
}

struct RunSpec {
  int inner_kind;        // 0 = Langevin, 1 = HMC, 2 = random-walk Metropolis
  int n_outer, n_inner;
  float step_size;
  int n_leapfrog;
  const float* inv_mass_diag;   // device or nullptr
  int local_adjusted, jump_adjusted;
  uint64_t seed, local_step0, jump_step0;
  int64_t chain0;
};

// slab sizes: about 1.25 x [12 * SMs CTAs' worth of chains] (12 = lcm of the kernels' 4 and 3 CTAs per SM), at least 32768
// chains, at most ~24 slabs -- large enough that a launch spans several waves, small enough that ~15 slabs interleave.
// With host copies the first and the last slabs are a quarter / a half of a regular one, so that the copy nothing can
// overlap with (H2D of the first slab, D2H of the last) is short.
int plan_slabs(int d, int64_t n, bool ramp_wanted, std::vector<int64_t>& sizes) {
  Layout L;
  if (!layout_for_dim(d, L)) return set_error("jump_sample: unsupported event size");
  const int64_t unit = 12ll * sm_count() * (kThreads / L.gs);
  int64_t slab = unit * ((32768 + unit - 1) / unit);
  // measured (2^20 chains, d = 100, unit = 56832 chains; device-resident / host-buffer chain-steps/s): 1 unit 5.68e9 /
  // 4.95e9, 1.25 units 5.73e9 / 5.46e9, 1.5 units 5.73e9 / 5.33e9, 2 units 5.55e9 / 5.27e9, 2.5 units 5.76e9 / 5.10e9
  // and, after the jump became two kernels (device / host): 0.67 units 5.81e9 / 5.63e9, 1.25 units 5.85e9 / 5.50e9,
  // 2 units 5.92e9 / 5.38e9 -- large slabs suit device-resident chains (fewer launch tails), small ones the host path
  // (shorter copies at both ends, finer interleaving of copy and compute)
  slab = ramp_wanted ? std::max<int64_t>(32768, (slab * 2 / 3 + 1023) / 1024 * 1024) : slab * 2;
  while ((n + slab - 1) / slab > 32) slab += unit;
  if (const char* ev = getenv("NFMC_SLAB_CHAINS")) { const long long v = atoll(ev); if (v >= 1024) slab = v; }
  const int64_t q = std::max<int64_t>(slab / 4 / 1024 * 1024, 1024), h = std::max<int64_t>(slab / 2 / 1024 * 1024, 1024);
  int64_t left = n;
  const bool ramp = ramp_wanted && n >= 4 * slab;
  if (ramp) { sizes.push_back(q); sizes.push_back(h); left -= q + h; }
  const int64_t tail = ramp ? q + h : 0;
  while (left - tail > 0) { const int64_t c = std::min<int64_t>(slab, left - tail); sizes.push_back(c); left -= c; }
  if (ramp) { sizes.push_back(h); sizes.push_back(q); }
  return 0;
}

// enqueue the whole run; `s` is forked into the pipe streams and joined again (no host synchronisation here).  Whatever
// happens in between -- including an error half-way through the slabs -- the side streams are joined back into `s` before
// returning, so the caller's later work on `s` never races with slabs still in flight.
int run_pipeline(const nfmc_potential& pot, const nfmc_realnvp& flow, float* x_dev, float* x_host, int64_t n, const RunSpec& R,
                 const nfmc_stats* st_local, const nfmc_stats* st_jump, float* logq_scratch, cudaStream_t s) {
  Pipe* pipe = nullptr;
  if (int e = ensure_pipe(pipe)) return e;
  const int d = pot.d;
  std::vector<int64_t> sizes;
  if (int e = plan_slabs(d, n, x_host != nullptr, sizes)) return e;
  if (int e = check_cuda(cudaEventRecord(pipe->fork, s), "fork record")) return e;
  for (int i = 0; i < kPipeStreams; ++i)
    if (int e = check_cuda(cudaStreamWaitEvent(pipe->streams[i], pipe->fork, 0), "fork wait")) return e;
  int err = 0;
  int64_t first = 0;
  for (size_t k = 0; k < sizes.size() && !err; first += sizes[k], ++k) {
    const int64_t cnt = sizes[k];
    cudaStream_t ss = pipe->streams[k % kPipeStreams];
    float* xs = x_dev + first * d;
    if (x_host) err = check_cuda(cudaMemcpyAsync(xs, x_host + first * d, (size_t)cnt * d * sizeof(float), cudaMemcpyHostToDevice, ss), "H2D x");
    for (int it = 0; it < R.n_outer && !err; ++it) {
      nfmc_rng r_local{R.seed, R.local_step0 + (uint64_t)it * (uint64_t)R.n_inner, nullptr, nullptr};
      nfmc_rng r_jump{R.seed, R.jump_step0 + (uint64_t)it, nullptr, nullptr};
      if (R.inner_kind == 0)
        err = nfmc_mala_steps(&pot, xs, cnt, R.n_inner, R.step_size, R.inv_mass_diag, R.local_adjusted, &r_local, R.chain0 + first, st_local, nullptr, ss);
      else if (R.inner_kind == 1)
        err = nfmc_hmc_steps(&pot, xs, cnt, R.n_inner, R.step_size, R.n_leapfrog, R.inv_mass_diag, R.local_adjusted, &r_local, R.chain0 + first, st_local, nullptr, ss);
      else
        err = nfmc_mh_steps(&pot, xs, cnt, R.n_inner, R.inv_mass_diag, R.local_adjusted, &r_local, R.chain0 + first, st_local, nullptr, ss);
      if (!err)
        err = nfmc_jump_step2(&pot, &flow, xs, logq_scratch ? logq_scratch + first : nullptr, cnt, R.jump_adjusted, &r_jump,
                              R.chain0 + first, st_jump, nullptr, ss);
    }
    if (x_host && !err)
      err = check_cuda(cudaMemcpyAsync(x_host + first * d, xs, (size_t)cnt * d * sizeof(float), cudaMemcpyDeviceToHost, ss), "D2H x");
  }
  for (int i = 0; i < kPipeStreams; ++i) {        // always join, also on error
    const int e1 = check_cuda(cudaEventRecord(pipe->join[i], pipe->streams[i]), "join record");
    const int e2 = e1 ? e1 : check_cuda(cudaStreamWaitEvent(s, pipe->join[i], 0), "join wait");
    if (!err) err = e2;
  }
  return err;
}
}  // namespace

extern "C" int64_t nfmc_jump_sample_slabs(int32_t d, int64_t n, int32_t host_buffers) {
  std::vector<int64_t> sizes;
  if (n < 1 || plan_slabs(d, n, host_buffers != 0, sizes)) return -1;
  return (int64_t)sizes.size();
}

extern "C" int nfmc_jump_sample_device(const nfmc_potential* pot, const nfmc_realnvp* flow, float* x, int64_t n,
                                       int32_t inner_kind, int32_t n_outer, int32_t n_inner, float step_size,
                                       int32_t n_leapfrog, const float* inv_mass_diag, int32_t local_adjusted,
                                       int32_t jump_adjusted, uint64_t seed, uint64_t local_step0, uint64_t jump_step0,
                                       int64_t chain0, const nfmc_stats* local_stats, const nfmc_stats* jump_stats,
                                       float* logq_scratch, void* stream) {
  if (int e = validate_pot(pot)) return e;
  if (!flow || !flow->blob || !x || n < 1) return set_error("jump_sample_device: NULL / empty argument");
  if (pot->d != flow->d) return set_error("jump_sample_device: potential and flow event sizes differ");
  if (inner_kind < 0 || inner_kind > 2 || n_outer < 0 || n_inner < 0) return set_error("jump_sample_device: bad inner_kind / iteration counts");
  if (jump_adjusted && !logq_scratch) return set_error("jump_sample_device: logq_scratch (n floats) is required for adjusted jumps");
  const RunSpec R{inner_kind, n_outer, n_inner, step_size, n_leapfrog, inv_mass_diag, local_adjusted, jump_adjusted,
                  seed, local_step0, jump_step0, chain0};
  return run_pipeline(*pot, *flow, x, nullptr, n, R, local_stats, jump_stats, logq_scratch, (cudaStream_t)stream);
}

extern "C" int nfmc_jump_sample_host(const nfmc_potential* pot_h, const float* pot_params_host, int64_t pot_params_floats,
                                     const nfmc_realnvp* flow_h, const float* blob_host, float* x_host, int64_t n,
                                     int32_t inner_kind, int32_t n_outer, int32_t n_inner, float step_size, int32_t n_leapfrog,
                                     uint64_t seed, int64_t chain0, double* sum_x_host, double* sum_x2_host,
                                     unsigned long long* counts_host, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!pot_h || !flow_h || !blob_host || !x_host || !workspace) return set_error("jump_sample_host: NULL argument");
  const int d = pot_h->d;
  if (workspace_bytes < nfmc_jump_workspace_bytes(d, n, flow_h->blob_floats)) return set_error("jump_sample_host: workspace too small");
  if (pot_params_floats > 2 * (int64_t)d) return set_error("jump_sample_host: too many potential parameters");
  if (inner_kind < 0 || inner_kind > 1) return set_error("jump_sample_host: inner_kind must be 0 (MALA) or 1 (HMC)");
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char* w = static_cast<unsigned char*>(workspace);
  float* x_dev = reinterpret_cast<float*>(w); w += align256((size_t)n * d * sizeof(float));
  float* blob_dev = reinterpret_cast<float*>(w); w += align256((size_t)flow_h->blob_floats * sizeof(float));
  float* pp_dev = reinterpret_cast<float*>(w); w += align256((size_t)2 * d * sizeof(float));
  double* mom_dev = reinterpret_cast<double*>(w); w += align256((size_t)2 * d * sizeof(double));
  unsigned long long* cnt_dev = reinterpret_cast<unsigned long long*>(w); w += align256(8 * sizeof(unsigned long long));
  float* logq_dev = reinterpret_cast<float*>(w);

  if (int e = check_cuda(cudaMemcpyAsync(blob_dev, blob_host, (size_t)flow_h->blob_floats * sizeof(float), cudaMemcpyHostToDevice, s), "H2D blob")) return e;
  if (pot_params_host && pot_params_floats > 0)
    if (int e = check_cuda(cudaMemcpyAsync(pp_dev, pot_params_host, (size_t)pot_params_floats * sizeof(float), cudaMemcpyHostToDevice, s), "H2D pot")) return e;
  cudaMemsetAsync(mom_dev, 0, (size_t)2 * d * sizeof(double), s);
  cudaMemsetAsync(cnt_dev, 0, 8 * sizeof(unsigned long long), s);

  nfmc_potential pot = *pot_h;
  pot.params = (pot_params_host && pot_params_floats > 0) ? pp_dev : nullptr;
  nfmc_realnvp flow = *flow_h;
  flow.blob = blob_dev;
  const nfmc_stats st_local{mom_dev, mom_dev + d, cnt_dev};
  const nfmc_stats st_jump{mom_dev, mom_dev + d, cnt_dev + 4};
  const RunSpec R{inner_kind, n_outer, n_inner, step_size, n_leapfrog, nullptr, 1, 1, seed, 0, 0, chain0};
  if (int e = run_pipeline(pot, flow, x_dev, x_host, n, R, &st_local, &st_jump, logq_dev, s)) return e;
  if (sum_x_host) cudaMemcpyAsync(sum_x_host, mom_dev, (size_t)d * sizeof(double), cudaMemcpyDeviceToHost, s);
  if (sum_x2_host) cudaMemcpyAsync(sum_x2_host, mom_dev + d, (size_t)d * sizeof(double), cudaMemcpyDeviceToHost, s);
  if (counts_host) cudaMemcpyAsync(counts_host, cnt_dev, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s);
  return check_cuda(cudaStreamSynchronize(s), "jump_sample_host sync");
}
