timeout 170 compute-sanitizer --tool memcheck --error-exitcode 86 python -m pytest -x -q -m gpu \
  "tests/test_gpu_gemm.py::test_fused_epilogue_bit_exact_vs_oracle" -k "out2 and shape3" \
  > gpurun_out/sanitize2_full.log 2>&1; echo "exit $?"; grep -v "Host Frame" gpurun_out/sanitize2_full.log | tail -8
