set -x
python -m pytest tests/test_gpu_baseline_size.py tests/test_gpu_scene.py -m gpu -x -q 2>&1 | tail -25
