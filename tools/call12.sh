timeout 1500 python -m pytest tests/test_gpu_training.py tests/test_gpu_fullsize.py -m gpu -q --timeout 900 --durations=12 2>&1 | tail -30
