timeout 120 python tools/prof_gemm.py 2048 4096 4096 16 20 0
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q -k "int32" 2>&1 | tail -5
for s in "2048 4096 4096" "2048 11008 4096" "2048 4096 11008" "4096 4096 4096" "8192 8192 8192" "2048 8192 28672" "2048 28672 8192"; do
  for cfg in -1 1 16 17; do timeout 120 python tools/prof_gemm.py $s $cfg 10 0; done
done
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -5
