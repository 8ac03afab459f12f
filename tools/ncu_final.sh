set -x
# (1) plain runs first (a number printed under ncu is never a bench value)
python tools/ab_step.py 64 50 1 > gpurun_out/plain_final.log 2>&1 || exit 1
# (2) launch list of one bench run (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2_final.csv python bench.py --steps 2 --warmup 3 --no-scene --fused 0 --no-cpu-baseline --pipeline 1 --sustain 0 > gpurun_out/ncu_launches.log 2>&1
# (3) whole step, caches not flushed: DRAM bytes, time, tensor pipe, instructions of all launches of one 64-tile step
ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread -s 1056 -c 88 --csv --log-file gpurun_out/step_metrics_r2_final.csv python tools/ab_step.py 64 1 1 > gpurun_out/ncu_step.log 2>&1
# (4) --set full of the dominant kernels of one step: the 56x56 3x3 pair layer (launch 59), the 112x112 expand layer (3), a 28x28 pair layer (52), depthwise (1), nms (85/86)
ncu --set full --import-source on --clock-control none --cache-control none -k regex:"conv_tc_pair|conv_tc_flc|depthwise3x3_lean|nms_kernel" -s 36 -c 12 -o /tmp/top_r2 -f python tools/ab_step.py 64 1 1 > gpurun_out/ncu_top.log 2>&1
ncu -i /tmp/top_r2.ncu-rep --page raw --csv > gpurun_out/top_r2_final_raw.csv
ls -la /tmp/top_r2.ncu-rep gpurun_out/
