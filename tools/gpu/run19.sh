python -m pytest tests/test_gpu_graph.py tests/test_gpu_conv_tc.py tests/test_gpu_baseline_size.py::test_config2_all_64_bench_tiles_against_oracle tests/test_gpu_property.py -m gpu -x -q 2>&1 | tail -4
for i in 1 2; do
TOD_TC_FL=0 python tools/ab_step.py 64 300
python tools/ab_step.py 64 300
done
