# ncu evidence for the end of round 2; outputs under gpurun_out/ (summaries are copied to profiles/ by hand)
set -x
# launch list of the bench command (eager launches: 8 per step; skip weight-quant + warm-up launches of OUR kernels)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"qgemm|rowwise_quant" -s 71 -c 48 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 6 --warmup 3 --no-graph > gpurun_out/bench_under_ncu_r2b.log 2>&1
# DRAM / L2 traffic of the step's four GEMM launches and four act-quant launches
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:"qgemm_kernel|rowwise_quant" -s 16 -c 8 --csv --log-file gpurun_out/step_traffic_r2b.csv python tools/step_once.py 4 > /dev/null 2>&1
# full captures: the staged quantizer (2048 x 11008, the step's 4th act-quant), the step's gate/up GEMM
ncu --set full --clock-control none --import-source on -k regex:rowwise_quant_staged -s 2 -c 1 -f -o gpurun_out/ncu_quant_staged_r2b python tools/step_once.py 4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:qgemm_kernel -s 10 -c 1 -f -o gpurun_out/ncu_qgemm_gateup_r2b python tools/step_once.py 4 > /dev/null 2>&1
for f in ncu_quant_staged_r2b ncu_qgemm_gateup_r2b; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null
done
ls -la gpurun_out/*_r2b*
