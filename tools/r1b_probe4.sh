timeout 120 python tools/prof_gemm.py 2048 4096 4096 0 20 0
timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -5
for s in "2048 4096 4096" "2048 11008 4096" "2048 4096 11008" "4096 3072 768" "2048 4096 128"; do
  for cfg in -1 0 1 8 11; do
    timeout 120 python tools/prof_gemm.py $s $cfg 20 0
    PQ_TMA_STORE=0 timeout 120 python tools/prof_gemm.py $s $cfg 20 0
  done
done
timeout 120 python tools/prof_gemm.py 8192 8192 8192 -1 10 0
PQ_TMA_STORE=0 timeout 120 python tools/prof_gemm.py 8192 8192 8192 -1 10 0
timeout 120 python tools/timeline.py 2048 4096 4096 0 0 | head -6
timeout 120 python tools/timeline.py 2048 4096 4096 11 0 | head -6
