timeout 900 python -m pytest tests/test_gpu_ambience.py tests/test_dropin.py -m gpu -q 2>&1 | tail -15
