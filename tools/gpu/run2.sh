set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/scene_probe.py 64 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'land_kernel|stamp_pruned|merge_kernel|weights_kernel' -s 4 -c 4 -o gpurun_out/scene_r2 python tools/scene_probe.py 64 > gpurun_out/ncu_scene_full.log 2>&1
tail -3 gpurun_out/ncu_scene_full.log
