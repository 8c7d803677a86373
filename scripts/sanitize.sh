#!/bin/bash
# compute-sanitizer over the parity tests (run under gpurun, 1 GPU; ~20 GPU-minutes): memcheck, racecheck and synccheck of
# this library's kernels (mangled names containing _ZN2lg, i.e. namespace lg; torch's own kernels are not instrumented) while the reference
# fixtures are replayed through both step routes, one fused BASELINE-size step runs against the oracle, and the ticket,
# direct-prefix and chained-launch variants of the pre-physics pass run their parity tests.
#   gpurun --timeout 1500 -- 'bash scripts/sanitize.sh r02'      -> gpurun_out/sanitizer_<tool>_r02.log
# Afterwards: cp gpurun_out/sanitizer_*_r02.log profiles/
cd "$(dirname "$0")/.."
R=${1:-r02}
mkdir -p gpurun_out
TESTS="tests/test_cuda_golden.py tests/test_cuda_vs_oracle.py::test_full_size_steps_match_oracle tests/test_cuda_round2.py::test_ticket_path_full_step_matches_oracle tests/test_cuda_round2.py::test_chained_pre_physics_equals_stream_order tests/test_cuda_round2.py::test_direct_prefix_counts_any_nonzero_flag_byte"
for tool in memcheck racecheck synccheck; do   # initcheck needs > 20 min on these tests: run it by hand when wanted
  extra=""
  [ $tool = memcheck ] && extra="--leak-check no --padding 32"
  [ $tool = racecheck ] && extra="--racecheck-report all"
  [ $tool = initcheck ] && extra=""
  timeout 1200 compute-sanitizer --tool $tool $extra --kernel-name kns=_ZN2lg --error-exitcode 66 --print-limit 20 \
      python -m pytest $TESTS -x -q -m gpu -p no:cacheprovider > gpurun_out/sanitizer_${tool}_$R.log 2>&1
  echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_$R.log | tail -1) | $(tail -1 gpurun_out/sanitizer_${tool}_$R.log)"
done
