B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'(^|[ :])k_[a-z]' -s 192 -c 64 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^k_ir_fft$' -s 24 -c 1 -f -o gpurun_out/prof_r02_irfft $B > gpurun_out/n1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^k_cmac$' -s 24 -c 1 -f -o gpurun_out/prof_r02_cmac $B > gpurun_out/n2.log 2>&1
ncu --set full --clock-control none -k regex:'^k_ifft_ola$' -s 24 -c 1 -f -o gpurun_out/prof_r02_ifft $B > gpurun_out/n3.log 2>&1
ncu --set full --clock-control none -k regex:'^k_cmac_static$' -s 24 -c 1 -f -o gpurun_out/prof_r02_cmacs $B > gpurun_out/n4.log 2>&1
ls gpurun_out/prof_r02_*.ncu-rep
