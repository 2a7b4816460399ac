# ncu evidence for the round-1 (session 2) kernels; outputs under gpurun_out/
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 42 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 6 --warmup 3 --no-graph > gpurun_out/bench_under_ncu.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum --clock-control none -k regex:qgemm_kernel -s 14 -c 7 --csv --log-file gpurun_out/gemm_traffic_r1b.csv python tools/step_once.py 6 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:qgemm_kernel -s 14 -c 1 -f -o gpurun_out/ncu_qgemm_r1b python tools/step_once.py 4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:rowwise_quant -s 14 -c 1 -f -o gpurun_out/ncu_quant_r1b python tools/step_once.py 4 > /dev/null 2>&1
python bench.py --steps 1000 --warmup 20 > gpurun_out/bench_r1b_1000.json 2> gpurun_out/bench_r1b_1000.err
python bench.py --steps 100 --warmup 10 > gpurun_out/bench_r1b_100.json 2> gpurun_out/bench_r1b_100.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r1b_ref.json 2> gpurun_out/bench_r1b_ref.err
