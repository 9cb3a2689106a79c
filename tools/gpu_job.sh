cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -x -p no:cacheprovider > gpurun_out/r3e_tests.log 2>&1; tail -5 gpurun_out/r3e_tests.log
python bench.py --all-workloads --steps 5 --warmup 3 > gpurun_out/r3e_bench_all.json 2> gpurun_out/r3e_bench_all.err; tail -3 gpurun_out/r3e_bench_all.err
