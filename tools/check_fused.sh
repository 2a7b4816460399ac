python -m pytest tests/test_gpu_fused.py -m gpu -x -q 2>&1 | tail -3
python tools/prof_fused.py 2>&1 | grep rmsnorm
