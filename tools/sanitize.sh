# compute-sanitizer memcheck over a small cross-section of the GPU tests (slow: keep the selection small)
compute-sanitizer --tool memcheck --error-exitcode 86 python -m pytest -x -q -m gpu \
  "tests/test_gpu_fused.py::test_rmsnorm_quant[shape0-dt0]" "tests/test_gpu_fused.py::test_layernorm_quant[shape0-dt0]" \
  "tests/test_gpu_fused.py::test_act_mul_quant_on_column_slices_of_one_gemm_output" \
  "tests/test_gpu_rowparallel.py::test_scatter_gemm_writes_each_column_block_to_its_destination" \
  "tests/test_gpu_rowparallel.py::test_reduce_dequant" \
  "tests/test_gpu_rowparallel.py::test_k_split_on_one_gpu_equals_unsplit_linear[shape2-4]" \
  "tests/test_gpu_gemm.py::test_persistent_scheduler_many_tiles" \
  "tests/test_gpu_gemm.py::test_default_heuristics_bit_exact" \
  "tests/test_gpu_quant.py" -k "not at_scale and not large" > gpurun_out/sanitize_full.log 2>&1; grep -v "Host Frame" gpurun_out/sanitize_full.log | tail -40
echo "sanitizer exit: $?"
