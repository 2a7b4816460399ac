for M in 128 256 512 1024; do for c in 0 1 2 3 4; do
python tools/prof_gemm.py $M 4096 4096 $c 20 0
done; done
for M in 256 512 1024; do for c in 0 1 2 4; do
python tools/prof_gemm.py $M 11008 4096 $c 20 0
python tools/prof_gemm.py $M 4096 11008 $c 20 0
done; done
for c in 0 1 2 4; do python tools/prof_gemm.py 4096 768 768 $c 20 0; python tools/prof_gemm.py 4096 3072 768 $c 20 0; python tools/prof_gemm.py 4096 768 3072 $c 20 0; done
