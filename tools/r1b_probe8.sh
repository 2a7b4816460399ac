python tools/prof_quant.py 2048 4096 bf16 0 0 128 4 64 8 256 2 128 3 64 6 32 8
python tools/prof_quant.py 2048 11008 bf16 0 0 512 4 256 6 512 3 256 8 1024 2
python tools/prof_quant.py 4096 768 bf16 0 0 32 3 32 4 64 2
python tools/prof_quant.py 4096 3072 bf16 0 0 64 6 128 3 64 8 32 8
python tools/prof_quant.py 2048 8192 bf16 0 0 128 8 256 4 512 2
