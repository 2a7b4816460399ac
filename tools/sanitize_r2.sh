# compute-sanitizer over the kernels added in round 2 (single GPU): staged quantizer (bulk-copy ring, mbarriers, named
# barriers), NaN-propagating reductions / careful paths, act_mul statistics mode, symm barrier is multi-GPU only.
set -x
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 86 python -m pytest -x -q -m gpu \
  "tests/test_gpu_quant.py::test_staged_quantizer_bit_exact" "tests/test_gpu_quant.py::test_act_quant_matches_exact_rational_golden" \
  "tests/test_gpu_quant.py::test_nonfinite_and_denormal_rows_follow_the_policy" \
  "tests/test_gpu_module.py::test_linear_with_nonfinite_tokens_follows_the_policy" \
  "tests/test_gpu_module.py::test_serialisation_roundtrip_keeps_the_bits" \
  "tests/test_gpu_rowparallel.py::test_parallel_gated_mlp_single_rank_equals_chained_modules" \
  > gpurun_out/sanitize_r2_memcheck.log 2>&1; echo "memcheck exit $?"; grep -v "Host Frame" gpurun_out/sanitize_r2_memcheck.log | tail -6
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 86 python -m pytest -x -q -m gpu \
  "tests/test_gpu_quant.py::test_staged_quantizer_bit_exact" -k "dtype0 or dtype2" \
  > gpurun_out/sanitize_r2_racecheck.log 2>&1; echo "racecheck exit $?"; grep -v "Host Frame" gpurun_out/sanitize_r2_racecheck.log | tail -6
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 86 python -m pytest -x -q -m gpu \
  "tests/test_gpu_quant.py::test_staged_quantizer_bit_exact" -k "dtype0 and shape3" \
  > gpurun_out/sanitize_r2_synccheck.log 2>&1; echo "synccheck exit $?"; grep -v "Host Frame" gpurun_out/sanitize_r2_synccheck.log | tail -6
