python tools/ab_step.py 64 1 > gpurun_out/plain_step.log 2>&1 && ncu --set full --cache-control none --clock-control none --import-source on -s 1059 -c 1 -o gpurun_out/op3_r2 python tools/ab_step.py 64 1 > gpurun_out/ncu_op3.log 2>&1
ncu -i gpurun_out/op3_r2.ncu-rep --page raw --csv > gpurun_out/op3_r2_raw.csv 2>/dev/null
ncu -i gpurun_out/op3_r2.ncu-rep --page source --csv > gpurun_out/op3_r2_source.csv 2>/dev/null
ls -la gpurun_out/op3_r2*
