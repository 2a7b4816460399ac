for M in 128 256 512 1024 2048 4096; do
python tools/prof_gemm.py $M 4096 4096 -1 20 0
python tools/prof_gemm.py $M 11008 4096 -1 20 0
python tools/prof_gemm.py $M 4096 11008 -1 20 0
done
for s in "4096 768 768" "4096 3072 768" "4096 768 3072" "2048 3584 8192" "2048 8192 28672" "2048 28672 8192" "8192 8192 8192"; do python tools/prof_gemm.py $s -1 20 0; done
python tools/prof_gemm.py 512 4096 4096 2 20 0; python tools/prof_gemm.py 512 4096 4096 0 20 0; python tools/prof_gemm.py 1024 4096 4096 0 20 0; python tools/prof_gemm.py 1024 4096 4096 2 20 0; python tools/prof_gemm.py 1024 4096 4096 11 20 0
python tools/prof_gemm.py 1024 11008 4096 0 20 0; python tools/prof_gemm.py 1024 11008 4096 1 20 0;python tools/prof_gemm.py 1024 11008 4096 12 20 0; python tools/prof_gemm.py 1024 11008 4096 13 20 0
python bench.py --steps 300 --warmup 10 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; tail -3 gpurun_out/bench_r1b.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_r1b.json') if l.startswith('{')][-1]); print({k:d[k] for k in ['value','ms_per_step','gpu_launches','clocks']}); r=d['roofline']; print(r['achieved'], r['frac'], r['avg_launch_ms'], r['act_quant']['achieved']); print(d['e2e']); print(d['decode_16tok'])"
