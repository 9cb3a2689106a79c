cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r6_tests.log 2>&1; tail -4 gpurun_out/r6_tests.log
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 200 python bench.py > gpurun_out/r6_bench_headline.json 2> gpurun_out/r6_bench_headline.err; head -c 300 gpurun_out/r6_bench_headline.json; echo
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r6_bench_reference.json 2> gpurun_out/r6_bench_reference.err; head -c 200 gpurun_out/r6_bench_reference.json; echo
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r6_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r6_ncu.log 2>&1; tail -c 300 gpurun_out/r6_ncu.log; wc -l gpurun_out/r6_launches_bench.csv
