for cfg in 0 1 11; do
ncu --set full --clock-control none --import-source on -k regex:qgemm_kernel -s 2 -c 1 -f -o gpurun_out/ncu_r1b_cfg$cfg python tools/one_gemm.py 2048 4096 4096 $cfg 3 > gpurun_out/ncu_r1b_cfg$cfg.log 2>&1
done
