B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'^(k_ir_fft|k_cmac|k_cmac_static|k_ifft_ola|k_x_fft|k_mix|k_amb_partial)$' -s 36 -c 12 --csv --log-file gpurun_out/r02_traffic.csv $B > gpurun_out/l.log 2>&1
