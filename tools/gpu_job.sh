cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -x -p no:cacheprovider > gpurun_out/r3j_tests.log 2>&1; tail -3 gpurun_out/r3j_tests.log
timeout 300 python -m pytest tests/test_gpu_parity_large.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -E "^\[parity|passed|failed" > gpurun_out/r3j_agreement.txt; cat gpurun_out/r3j_agreement.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r3j_bench_headline.json 2> gpurun_out/r3j_bench_headline.err; tail -2 gpurun_out/r3j_bench_headline.err; head -c 600 gpurun_out/r3j_bench_headline.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3j_bench_reference.json 2> gpurun_out/r3j_bench_reference.err; head -c 400 gpurun_out/r3j_bench_reference.json; echo
python __graft_entry__.py smoke 2>&1 | tail -2
