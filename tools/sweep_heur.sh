for M in 96 128 256 512 1024 2048 4096; do
python tools/prof_gemm.py $M 4096 4096 -1 20 0
python tools/prof_gemm.py $M 11008 4096 -1 20 0
python tools/prof_gemm.py $M 4096 11008 -1 20 0
done
python tools/prof_gemm.py 4096 768 768 -1 20 0; python tools/prof_gemm.py 4096 3072 768 -1 20 0; python tools/prof_gemm.py 4096 768 3072 -1 20 0
python tools/prof_gemm.py 8192 8192 8192 -1 10 0
python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -2
python tools/prefetch_check.py 2>&1 | head -2
python bench.py --steps 300 --warmup 10 > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; tail -3 gpurun_out/bench_r1f.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_r1f.json') if l.startswith('{')][-1]); print({k:d[k] for k in ['value','ms_per_step','gpu_launches','clocks']}); r=d['roofline']; print(r['achieved'], r['frac'], r['avg_launch_ms'], r['act_quant']['achieved']); print(d['e2e']); print(d['decode_16tok'])"
