#!/bin/bash
# round-2 probe 1: in-kernel SM clock at the headline shapes, and ncu --set full of ours vs cuBLASLt int8 on one shape
set -x
for it in 1 8 40; do
  python tools/clock_in_kernel.py 2048 11008 4096 1 $it
  python tools/clock_in_kernel.py 2048 11008 4096 8 $it
  python tools/clock_in_kernel.py 2048 4096 4096 11 $it
done
python tools/clock_in_kernel.py 8192 8192 8192 1 1
python tools/clock_in_kernel.py 8192 8192 8192 1 8
python tools/clock_in_kernel.py 2048 22016 4096 8 8
ncu --set full --clock-control none --import-source on -k regex:qgemm -s 3 -c 1 -o gpurun_out/ncu_r2_ours_11008 -f python tools/one_gemm.py 2048 11008 4096 1 5 > /dev/null 2>&1
ncu --set full --clock-control none -s 3 -c 1 -o gpurun_out/ncu_r2_intmm_11008 -f python tools/one_intmm.py 2048 11008 4096 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
