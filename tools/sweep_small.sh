python -m pytest tests/test_gpu_gemm.py -m gpu -x -q -k "small_m" 2>&1 | tail -4
for c in 7; do
python tools/prof_gemm.py 1 4096 4096 $c 20 0
python tools/prof_gemm.py 16 4096 4096 $c 20 0
python tools/prof_gemm.py 16 11008 4096 $c 20 0
python tools/prof_gemm.py 16 4096 11008 $c 20 0
python tools/prof_gemm.py 32 4096 4096 $c 20 0
python tools/prof_gemm.py 64 4096 4096 $c 20 0


python tools/prof_gemm.py 16 28672 8192 $c 20 0
python tools/prof_gemm.py 16 8192 28672 $c 20 0
done
python tools/prof_gemm.py 8 768 3072 7 20 0
python tools/prof_gemm.py 8 3072 768 7 20 0
python tools/prof_gemm.py 96 4096 4096 -1 20 0
python tools/prof_gemm.py 128 11008 4096 -1 20 0
