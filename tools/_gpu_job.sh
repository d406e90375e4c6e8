export ALR_WATCHDOG_MS=120000
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_round2.py tests/test_gpu_ambience.py -m gpu -q -x -k "small_rir or dry_window or (fused and not tight) or seeded or missing_seed or mixes_like" > gpurun_out/race.txt 2>&1
grep -E "Error|Warning|hazard|at alr|at .*\.cuh|Saved host" gpurun_out/race.txt | head -60
