cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout 100 --timeout-method=thread -x -k "beam" -p no:cacheprovider > gpurun_out/r3n_tests.log 2>&1; tail -3 gpurun_out/r3n_tests.log
BEAM=10 timeout 120 python tools/phase_clocks.py vrptw50 2960 tc3 0 > gpurun_out/r3n_beam_clocks.txt 2>&1
cat gpurun_out/r3n_beam_clocks.txt | tail -8
timeout 200 python bench.py --workload vrptw50_beam10 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r3n_bench_beam.json 2>gpurun_out/r3n_bench_beam.err; head -c 300 gpurun_out/r3n_bench_beam.json; echo
