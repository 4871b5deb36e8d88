// host_common.cuh -- host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <cstdlib>
#include "../../include/nfmc_b200.h"
#include "potentials.cuh"

namespace nfmc {

int set_error(const std::string& msg);  // records the message, returns 1
int check_cuda(cudaError_t e, const char* what);
int sm_count();

struct Layout {
  int gs;      // lanes per chain
  int E;       // slots per half (template value)
  bool exact;  // ceil(db / gs) == E: every slot below the last is valid in both halves for every lane
};
// smallest power-of-two group such that ceil(db / gs) <= 16, then the smallest instantiated E that fits
inline bool layout_for_dim(int d, Layout& L) {
  if (d < 1 || d > NFMC_MAX_DIM) return false;
  const int db = d - d / 2;
  int gs = 1;
  static const int max_slots = [] { const char* e = getenv("NFMC_LAYOUT_MAX_SLOTS"); int v = e ? atoi(e) : 16; return (v >= 4 && v <= 16) ? v : 16; }();
  while ((db + gs - 1) / gs > max_slots) gs <<= 1;
  if (gs > 32) return false;
  const int e = (db + gs - 1) / gs;
  L.gs = gs;
  L.E = e <= 4 ? 4 : e <= 7 ? 7 : e <= 13 ? 13 : 16;
  L.exact = (e == L.E);
  return true;
}

// persistent grid sized from what the kernel can actually keep resident (its registers and shared memory decide, not a
// guess): min(tiles, resident CTAs per SM x SMs).  A grid larger than that leaves a partial second wave behind.
template <typename K>
inline int occupancy_grid(K kernel, size_t smem, int64_t n, int gs) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int64_t chains_per_cta = kThreads / gs;
  const int64_t tiles = (n + chains_per_cta - 1) / chains_per_cta;
  const int64_t cap = (int64_t)sm_count() * per_sm;
  return (int)(tiles < cap ? tiles : cap);
}

// most lanes per chain: the smallest instantiated E whose group still fits a warp.  For latency-bound launches (a
// training minibatch is a few hundred rows) the serial work per lane is what matters, not lane efficiency.
inline bool layout_wide(int d, Layout& L) {
  if (d < 1 || d > NFMC_MAX_DIM) return false;
  const int db = d - d / 2;
  const int Es[4] = {4, 7, 13, 16};
  for (int i = 0; i < 4; ++i) {
    int gs = 1;
    while ((db + gs - 1) / gs > Es[i]) gs <<= 1;
    if (gs <= 32) {
      L.gs = gs; L.E = Es[i]; L.exact = ((db + gs - 1) / gs == Es[i]);
      return true;
    }
  }
  return false;
}

inline PotParams pot_params(const nfmc_potential* p) {
  PotParams P;
  P.params = p->params;
  P.s0 = p->scalar[0]; P.s1 = p->scalar[1]; P.s2 = p->scalar[2]; P.s3 = p->scalar[3];
  if (p->kind == NFMC_POT_FUNNEL) P.s1 = 1.0f / (2.0f * p->scalar[0] * p->scalar[0]);
  return P;
}
inline int validate_pot(const nfmc_potential* p) {
  if (!p) return set_error("potential is NULL");
  if (p->d < 1 || p->d > NFMC_MAX_DIM) return set_error("potential: d out of range [1, 1024]");
  if (p->kind == NFMC_POT_DIAG_GAUSSIAN && !p->params) return set_error("diag gaussian needs params");
  if (p->kind == NFMC_POT_ROSENBROCK && (p->d % 2)) return set_error("rosenbrock needs even d");
  if ((p->kind == NFMC_POT_FUNNEL || p->kind == NFMC_POT_MIXTURE4) && p->d < 2) return set_error("potential needs d >= 2");
  if (p->kind < 0 || p->kind > NFMC_POT_MIXTURE4) return set_error("unknown potential kind");
  return 0;
}

// persistent grid: enough CTAs to fill the machine, never more than there are tiles
inline int grid_for(int64_t n, int gs, int ctas_per_sm) {
  const int64_t chains_per_cta = kThreads / gs;
  const int64_t tiles = (n + chains_per_cta - 1) / chains_per_cta;
  const int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  return (int)(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
}

}  // namespace nfmc
