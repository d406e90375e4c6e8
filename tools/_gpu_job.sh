B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'(^|[ :])k_[a-z]' -s 192 -c 64 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/l.log 2>&1
