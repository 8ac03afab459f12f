set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err; echo rc=$?
tail -c 600 gpurun_out/bench_r2_a.err
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
