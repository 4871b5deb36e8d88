// api.cu -- error plumbing, device queries and the host-buffer end-to-end entry point of the C ABI.
#include <cstring>
#include <string>
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
                   align256((size_t)2 * d * sizeof(float)) + align256((size_t)2 * d * sizeof(double)) + align256(8 * sizeof(unsigned long long)));
}

// Host-buffer path: x0 in, final state + pooled statistics out; every copy is inside the call.
extern "C" int nfmc_jump_sample_host(const nfmc_potential* pot_h, const float* pot_params_host, int64_t pot_params_floats,
                                     const nfmc_realnvp* flow_h, const float* blob_host, float* x_host, int64_t n,
                                     int32_t inner_kind, int32_t n_outer, int32_t n_inner, float step_size, int32_t n_leapfrog,
                                     uint64_t seed, int64_t chain0, double* sum_x_host, double* sum_x2_host,
                                     unsigned long long* counts_host, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!pot_h || !flow_h || !blob_host || !x_host || !workspace) return set_error("jump_sample_host: NULL argument");
  const int d = pot_h->d;
  if (workspace_bytes < nfmc_jump_workspace_bytes(d, n, flow_h->blob_floats)) return set_error("jump_sample_host: workspace too small");
  if (pot_params_floats > 2 * (int64_t)d) return set_error("jump_sample_host: too many potential parameters");
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char* w = static_cast<unsigned char*>(workspace);
  float* x_dev = reinterpret_cast<float*>(w); w += align256((size_t)n * d * sizeof(float));
  float* blob_dev = reinterpret_cast<float*>(w); w += align256((size_t)flow_h->blob_floats * sizeof(float));
  float* pp_dev = reinterpret_cast<float*>(w); w += align256((size_t)2 * d * sizeof(float));
  double* mom_dev = reinterpret_cast<double*>(w); w += align256((size_t)2 * d * sizeof(double));
  unsigned long long* cnt_dev = reinterpret_cast<unsigned long long*>(w);

  if (int e = check_cuda(cudaMemcpyAsync(x_dev, x_host, (size_t)n * d * sizeof(float), cudaMemcpyHostToDevice, s), "H2D x")) return e;
  if (int e = check_cuda(cudaMemcpyAsync(blob_dev, blob_host, (size_t)flow_h->blob_floats * sizeof(float), cudaMemcpyHostToDevice, s), "H2D blob")) return e;
  if (pot_params_host && pot_params_floats > 0)
    if (int e = check_cuda(cudaMemcpyAsync(pp_dev, pot_params_host, (size_t)pot_params_floats * sizeof(float), cudaMemcpyHostToDevice, s), "H2D pot")) return e;
  cudaMemsetAsync(mom_dev, 0, (size_t)2 * d * sizeof(double), s);
  cudaMemsetAsync(cnt_dev, 0, 8 * sizeof(unsigned long long), s);

  nfmc_potential pot = *pot_h;
  pot.params = (pot_params_host && pot_params_floats > 0) ? pp_dev : nullptr;
  nfmc_realnvp flow = *flow_h;
  flow.blob = blob_dev;
  nfmc_stats st_local{mom_dev, mom_dev + d, cnt_dev};
  nfmc_stats st_jump{mom_dev, mom_dev + d, cnt_dev + 4};
  for (int it = 0; it < n_outer; ++it) {
    nfmc_rng r_local{seed, (uint64_t)it * (uint64_t)n_inner, nullptr, nullptr};
    nfmc_rng r_jump{seed, (uint64_t)it, nullptr, nullptr};
    int e = inner_kind == 0
                ? nfmc_mala_steps(&pot, x_dev, n, n_inner, step_size, nullptr, 1, &r_local, chain0, &st_local, nullptr, stream)
                : nfmc_hmc_steps(&pot, x_dev, n, n_inner, step_size, n_leapfrog, nullptr, 1, &r_local, chain0, &st_local, nullptr, stream);
    if (e) return e;
    if ((e = nfmc_jump_step(&pot, &flow, x_dev, n, 1, &r_jump, chain0, &st_jump, nullptr, stream))) return e;
  }
  if (int e = check_cuda(cudaMemcpyAsync(x_host, x_dev, (size_t)n * d * sizeof(float), cudaMemcpyDeviceToHost, s), "D2H x")) return e;
  if (sum_x_host) cudaMemcpyAsync(sum_x_host, mom_dev, (size_t)d * sizeof(double), cudaMemcpyDeviceToHost, s);
  if (sum_x2_host) cudaMemcpyAsync(sum_x2_host, mom_dev + d, (size_t)d * sizeof(double), cudaMemcpyDeviceToHost, s);
  if (counts_host) cudaMemcpyAsync(counts_host, cnt_dev, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s);
  return check_cuda(cudaStreamSynchronize(s), "jump_sample_host sync");
}
