timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_t2.log
cat gpurun_out/r02_t2.log
