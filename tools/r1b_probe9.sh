for s in "128 1024 16384" "128 2048 16384" "128 4096 16384" "128 1024 8192" "128 2048 8192" "128 4096 8192" "128 1024 4096" "128 4096 4096" "256 1024 16384" "256 2048 8192" "256 4096 4096" "512 1024 16384" "96 4096 11008" "128 4096 11008" "128 11008 4096"; do
  python tools/prof_gemm.py $s -1 20 0
  python tools/prof_gemm.py $s -1 20 1
  python tools/prof_gemm.py $s 2 20 1
  python tools/prof_gemm.py $s 0 20 1
done
for s in "64 1024 1024" "64 4096 1024" "64 4096 2048" "16 4096 1024" "16 4096 2048" "32 4096 2048" "64 16384 2048" "48 4096 4096"; do
  python tools/prof_gemm.py $s -1 20 0
  python tools/prof_gemm.py $s 3 20 0
done
