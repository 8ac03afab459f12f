python -m pytest tests/test_gpu_graph.py tests/test_gpu_conv_tc.py tests/test_gpu_baseline_size.py::test_config2_all_64_bench_tiles_against_oracle tests/test_gpu_property.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
TOD_TC_RELU=0 TOD_RELU_TAB=0 TOD_TC_FL=0 python tools/ab_step.py 64 300
python tools/ab_step.py 64 300
done
python tools/layer_table.py > gpurun_out/layers_r2_a.txt 2>&1; tail -5 gpurun_out/layers_r2_a.txt
