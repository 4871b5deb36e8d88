#!/bin/bash
# compute-sanitizer over a reduced GPU subset: one case (or a few) per kernel family, memcheck then racecheck.
# Usage (on a GPU box):  bash tools/sanitize.sh [outdir]        summaries land in <outdir>/sanitizer_{memcheck,racecheck}.txt
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
SEL='golden_mala or golden_hmc or golden_jump or golden_fixed_imh or golden_neutra_hmc or golden_neutra_mh or golden_adaptive_imh or golden_ess or golden_tess or golden_dlmc or golden_warmup or golden_mh or golden_ula or golden_uhmc or test_philox_mode_equals_injected_mode or (test_tc_forward_inverse_logprob and 64) or (test_tc_fused_jump_equals_composed_launches and 257) or (test_nll_gradient_matches_autograd and 6-2-4) or (test_reverse_kl_gradient_matches_autograd and 6-2-4) or (test_wide_nll_gradient_matches_autograd and 6-2-2-16) or (test_wide_reverse_kl_gradient_matches_autograd and 6-2-2-16) or test_adamw_matches_torch or (test_realnvp_forward_inverse_logprob and 0) or test_potential_odd_dims'
for tool in memcheck racecheck; do
  log="$OUT/sanitizer_${tool}.log"
  timeout 1500 compute-sanitizer --tool $tool --target-processes all --error-exitcode 0 --print-limit 20 \
      python -m pytest tests -q -m gpu -x -k "$SEL" -p no:cacheprovider > "$log" 2>&1
  echo "exit $?" >> "$log"
  { echo "== compute-sanitizer --tool $tool =="; grep -E "passed|failed|error" "$log" | tail -3;
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Race reported|Invalid|hazard" "$log" | sort | uniq -c | sort -rn | head -20; } > "$OUT/sanitizer_${tool}.txt"
  cat "$OUT/sanitizer_${tool}.txt"
done
