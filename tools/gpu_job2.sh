cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r3l_bench_n2.json 2> gpurun_out/r3l_bench_n2.err; tail -2 gpurun_out/r3l_bench_n2.err; head -c 400 gpurun_out/r3l_bench_n2.json; echo
