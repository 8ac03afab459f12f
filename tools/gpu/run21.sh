python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r2_b.json 2> gpurun_out/bench_r2_b.err; echo rc=$?; tail -c 400 gpurun_out/bench_r2_b.err
