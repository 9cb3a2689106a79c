cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -x -p no:cacheprovider > gpurun_out/r3x_tests.log 2>&1; tail -3 gpurun_out/r3x_tests.log
python bench.py --all-workloads --steps 5 --warmup 3 > gpurun_out/r3x_bench_all.json 2> gpurun_out/r3x_bench_all.err; tail -2 gpurun_out/r3x_bench_all.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r3x_bench_headline.json 2> gpurun_out/r3x_bench_headline.err; head -c 250 gpurun_out/r3x_bench_headline.json; echo
UPS=1 timeout 60 python tools/phase_clocks.py vrptw100 14208 tc4x 48 > gpurun_out/r3x_clocks_final.txt 2>&1
python __graft_entry__.py smoke 2>&1 | tail -1
