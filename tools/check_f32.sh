python -m pytest tests/test_gpu_gemm.py tests/test_gpu_module.py tests/test_gpu_quant.py -m gpu -x -q 2>&1 | tail -3
for s in "2048 4096 4096" "2048 11008 4096" "4096 3072 768"; do
  PQ_OUT=f32 python tools/prof_gemm.py $s -1 20 0
  PQ_OUT=f32 PQ_TMA_STORE=0 python tools/prof_gemm.py $s -1 20 0
done
python tools/prof_gemm.py 2048 4096 4096 -1 20 0
