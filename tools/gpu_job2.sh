cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r4a_bench_n8.json 2> gpurun_out/r4a_bench_n8.err; tail -1 gpurun_out/r4a_bench_n8.err; head -c 260 gpurun_out/r4a_bench_n8.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 10 --warmup 3 --workload vrptw100_stochastic --no-cpu-baseline > gpurun_out/r4a_bench_n8_cfg4.json 2> gpurun_out/r4a_bench_n8_cfg4.err; head -c 260 gpurun_out/r4a_bench_n8_cfg4.json; echo
