set -x
nproc; nvidia-smi topo -m 2>/dev/null | head -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; echo rc=$?
tail -c 1500 gpurun_out/bench_r2_n2.err
