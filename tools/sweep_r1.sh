python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -4
for sk in 0 1; do for c in 1 0; do
python tools/prof_gemm.py 2048 4096 4096 $c 20 $sk
python tools/prof_gemm.py 2048 11008 4096 $c 20 $sk
python tools/prof_gemm.py 2048 4096 11008 $c 20 $sk
done; done
for sk in 0 1; do for c in 0 3; do
python tools/prof_gemm.py 16 4096 4096 $c 20 $sk
python tools/prof_gemm.py 16 11008 4096 $c 20 $sk
python tools/prof_gemm.py 16 4096 11008 $c 20 $sk
python tools/prof_gemm.py 128 4096 4096 $c 20 $sk
done; done
for sk in 0 1; do
python tools/prof_gemm.py 512 4096 4096 1 20 $sk
python tools/prof_gemm.py 1024 4096 4096 1 20 $sk
python tools/prof_gemm.py 512 4096 4096 0 20 $sk
python tools/prof_gemm.py 256 4096 4096 0 20 $sk
python tools/prof_gemm.py 4096 3072 768 1 20 $sk
python tools/prof_gemm.py 4096 768 3072 1 20 $sk
python tools/prof_gemm.py 4096 768 768 1 20 $sk
python tools/prof_gemm.py 4096 768 768 0 20 $sk
done
