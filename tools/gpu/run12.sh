python -m pytest tests/test_gpu_path.py -m gpu -x -q 2>&1 | tail -12
