timeout 900 python -m pytest tests/test_training_gpu.py tests/test_abi.py -x -q 2>&1 | tail -5
