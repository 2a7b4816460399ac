python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -3
for s in "8192 8192 8192" "4096 4096 8192" "2048 4096 4096"; do
  python tools/clock_probe.py $s -1 2.5
  PQ_EPI=2 python tools/clock_probe.py $s -1 2.5
done
for s in "128 1024 16384" "128 4096 11008" "64 4096 1024" "64 4096 2048" "128 4096 4096" "2048 4096 4096"; do python tools/prof_gemm.py $s -1 20 -1; done
