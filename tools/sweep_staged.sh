python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -3
for st in 0 1; do export PQ_STAGED=$st
python tools/prof_gemm.py 4096 768 768 -1 20 0
python tools/prof_gemm.py 4096 3072 768 -1 20 0
python tools/prof_gemm.py 4096 768 3072 -1 20 0
python tools/prof_gemm.py 2048 4096 4096 -1 20 0
python tools/prof_gemm.py 2048 11008 4096 -1 20 0
python tools/prof_gemm.py 2048 4096 11008 -1 20 0
python tools/prof_gemm.py 1024 4096 4096 -1 20 0
python tools/prof_gemm.py 4096 4096 1024 -1 20 0
python tools/prof_gemm.py 8192 8192 1024 -1 20 0
python tools/prof_gemm.py 8192 8192 8192 -1 10 0
done
