python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 100 --warmup 10 > gpurun_out/bench_r1c_100.json 2> gpurun_out/bench_r1c_100.err; tail -2 gpurun_out/bench_r1c_100.err
python bench.py > gpurun_out/bench_r1c_default.json 2> gpurun_out/bench_r1c_default.err; tail -2 gpurun_out/bench_r1c_default.err
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
