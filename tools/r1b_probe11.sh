ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:qgemm_kernel -s 1 -c 1 --csv --log-file gpurun_out/big_traffic_a.csv python tools/one_gemm.py 8192 16384 16384 -1 2 > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:qgemm_kernel -s 1 -c 1 --csv --log-file gpurun_out/big_traffic_b.csv python tools/one_gemm.py 4096 4096 16384 -1 2 > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:qgemm_kernel -s 1 -c 1 --csv --log-file gpurun_out/big_traffic_c.csv python tools/one_gemm.py 8192 8192 8192 -1 2 > /dev/null 2>&1
python tools/prof_gemm.py 8192 16384 16384 -1 3 0
python tools/prof_gemm.py 8192 16384 16384 1 3 0
python tools/prof_gemm.py 4096 4096 16384 -1 10 0
python tools/prof_gemm.py 4096 4096 16384 1 10 0
python tools/prof_gemm.py 4096 4096 16384 0 10 0
