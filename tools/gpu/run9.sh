python -m pytest tests/test_gpu_scene.py tests/test_gpu_property.py::test_scene_random_frames tests/test_gpu_fused.py -m gpu -x -q 2>&1 | tail -3
for v in 1 0 2 9; do echo variant $v; TOD_WEIGHTS_VARIANT=$v python tools/scene_probe.py 256; done
