B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --scenes-per-gpu 64"
ALR_RING_MB=64 ncu --set full --clock-control none --import-source on -k regex:k_mov_fused -s 3 -c 1 -f -o gpurun_out/prof_fused_r64 $B > gpurun_out/ncu1.log 2>&1
ALR_RING_MB=1024 ALR_LOOKAHEAD=8 ncu --set full --clock-control none --import-source on -k regex:k_mov_fused -s 3 -c 1 -f -o gpurun_out/prof_fused_r1024 $B > gpurun_out/ncu2.log 2>&1
ALR_RING_MB=96 ALR_LOOKAHEAD=2 ncu --set full --clock-control none -k regex:k_mov_fused -s 3 -c 1 -f -o gpurun_out/prof_fused_r96 $B > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log
ls -la gpurun_out/*.ncu-rep
