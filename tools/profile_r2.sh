# ncu evidence for round 2; outputs under gpurun_out/ (summaries are copied to profiles/ by hand)
set -x
# launch list of the bench command (eager launches: 8 per step; skip weight-quant + warm-up launches of OUR kernels)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"qgemm|rowwise_quant" -s 71 -c 48 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 6 --warmup 3 --no-graph > gpurun_out/bench_under_ncu_r2.log 2>&1
# DRAM / L2 traffic of the step's four GEMM launches
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:qgemm_kernel -s 8 -c 4 --csv --log-file gpurun_out/gemm_traffic_r2.csv python tools/step_once.py 4 > /dev/null 2>&1
# full captures: the step's gate/up GEMM (2048 x 22016 x 4096), a 2048 x 11008 x 4096 GEMM, the step's first act-quant launch
ncu --set full --clock-control none --import-source on -k regex:qgemm_kernel -s 10 -c 1 -f -o gpurun_out/ncu_qgemm_gateup_r2 python tools/step_once.py 4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:qgemm -s 3 -c 1 -f -o gpurun_out/ncu_qgemm_11008_r2 python tools/one_gemm.py 2048 11008 4096 -1 5 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:rowwise_quant -s 8 -c 1 -f -o gpurun_out/ncu_quant_r2 python tools/step_once.py 4 > /dev/null 2>&1
ls -la gpurun_out/*_r2*
