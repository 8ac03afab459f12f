python -m pytest tests/test_gpu_scene.py tests/test_gpu_property.py::test_scene_random_frames tests/test_gpu_fused.py -m gpu -x -q 2>&1 | tail -3
python tools/scene_bench.py 256 0
TOD_WEIGHTS_VARIANT=1 python tools/scene_bench.py 256 0
TOD_STAMP_IMPL=1 python tools/scene_bench.py 256 0
