set -x
python -m pytest tests/test_gpu_scene.py -m gpu -x -q 2>&1 | tail -15
for i in 0 1; do TOD_STAMP_IMPL=$i python tools/scene_probe.py 256; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/scene_launches_r2.csv python tools/scene_probe.py 64 > gpurun_out/ncu_scene.log 2>&1
grep -E "land_kernel|stamp_|merge_kernel|weights_kernel|balls" gpurun_out/scene_launches_r2.csv | awk -F'","' '{print $5, $NF}' | tail -12
