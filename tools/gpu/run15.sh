for d in 2 3 4 6; do python bench.py --steps 100 --warmup 5 --no-scene --fused 0 --no-cpu-baseline --sustain 0 --pipeline $d 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('pipeline', $d, 'value %.0f' % d['value'], 'ms %.4f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], 'single %.0f' % d['single_stream']['value'])
"; done
python tools/ab_step.py 64 1 > gpurun_out/plain_step.log 2>&1 && ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed -s 1056 -c 88 --csv --log-file gpurun_out/step_dram_r2.csv python tools/ab_step.py 64 1 > gpurun_out/ncu_step.log 2>&1
wc -l gpurun_out/step_dram_r2.csv
