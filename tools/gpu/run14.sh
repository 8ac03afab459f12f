python -m pytest tests/test_gpu_scene.py tests/test_gpu_determinism.py tests/test_gpu_property.py::test_scene_random_frames -m gpu -x -q 2>&1 | tail -3
python tools/scene_bench.py 256 0
python tools/scene_probe.py 64 > gpurun_out/plain_scene.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'land_kernel|stamp_pruned|weights_literal' -s 3 -c 3 -o gpurun_out/scene_r2_final python tools/scene_probe.py 64 > gpurun_out/ncu_scene_final.log 2>&1
ncu -i gpurun_out/scene_r2_final.ncu-rep --page raw --csv > gpurun_out/scene_r2_final_raw.csv 2>/dev/null
python tools/scene_probe.py 64 > gpurun_out/plain_scene2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/scene_launches_r2.csv python tools/scene_probe.py 64 > gpurun_out/ncu_scene_l.log 2>&1
python tools/ab_step.py 64 1 > gpurun_out/plain_step.log 2>&1 && ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed -s 1140 -c 95 --csv --log-file gpurun_out/step_dram_r2.csv python tools/ab_step.py 64 1 > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log; wc -l gpurun_out/step_dram_r2.csv
