python -m pytest tests/test_gpu_scene.py -m gpu -x -q 2>&1 | tail -3
for v in 1 0 2 9 8 10; do echo variant $v; TOD_WEIGHTS_VARIANT=$v python tools/scene_probe.py 256; done
