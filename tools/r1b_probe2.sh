python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -3
for s in "2048 4096 4096" "2048 11008 4096" "2048 4096 11008"; do
  for cfg in -1 1 0 8 11; do python tools/prof_gemm.py $s $cfg 20 0; done
done
python tools/prof_gemm.py 2048 8192 8192 -1 20 0; python tools/prof_gemm.py 2048 8192 8192 9 20 0; python tools/prof_gemm.py 2048 8192 8192 1 20 0
python tools/prof_gemm.py 2048 3584 8192 10 20 0; python tools/prof_gemm.py 2048 3584 8192 13 20 0; python tools/prof_gemm.py 2048 3584 8192 1 20 0; python tools/prof_gemm.py 2048 3584 8192 0 20 0
for cfg in 0 1; do
  python tools/prof_gemm.py 2048 4096 128 $cfg 20 0
  PQ_EPI=1 python tools/prof_gemm.py 2048 4096 128 $cfg 20 0
  PQ_STAGED=1 python tools/prof_gemm.py 2048 4096 128 -1 20 0
  PQ_EPI=1 python tools/prof_gemm.py 2048 4096 4096 $cfg 20 0
done
python tools/timeline.py 2048 4096 128 0 0 | head -8
python tools/timeline.py 2048 4096 128 1 0 | head -8
python tools/timeline.py 2048 4096 4096 0 0 | head -8
python tools/timeline.py 2048 4096 4096 1 0 | head -8
python tools/timeline.py 2048 4096 4096 8 0 | head -8
