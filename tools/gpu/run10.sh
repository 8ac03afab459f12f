python -m pytest tests/test_gpu_scene.py tests/test_gpu_property.py::test_scene_random_frames tests/test_gpu_fused.py -m gpu -x -q 2>&1 | tail -3
python tools/scene_bench.py 256 0,16,32,64
TOD_WEIGHTS_VARIANT=0 python tools/scene_bench.py 256 0,32
TOD_WEIGHTS_VARIANT=2 python tools/scene_bench.py 256 0,32
