python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -5
compute-sanitizer --tool racecheck --error-exitcode 86 python -m pytest -x -q -m gpu \
  "tests/test_gpu_fused.py::test_rmsnorm_quant[shape1-dt0]" "tests/test_gpu_fused.py::test_layernorm_quant[shape0-dt0]" \
  "tests/test_gpu_rowparallel.py::test_reduce_dequant" "tests/test_gpu_rowparallel.py::test_row_absmax_and_quantize_with_external_max" \
  > gpurun_out/racecheck_full.log 2>&1; echo "racecheck exit $?"; grep -v "Host Frame" gpurun_out/racecheck_full.log | tail -12
