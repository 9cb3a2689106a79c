cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 40 python tools/time_variant.py libvrpb200.so tc4x 14208 greedy ups > gpurun_out/r3v_templ.txt 2>&1 || { echo HUNG; tail -3 gpurun_out/r3v_templ.txt; exit 0; }
timeout 40 python tools/time_variant.py libvrpb200.so auto 8192 stochastic uniform vrptw100 0 >> gpurun_out/r3v_templ.txt 2>&1
timeout 40 python tools/time_variant.py libvrpb200.so auto 4096 greedy uniform vrptw20 0 >> gpurun_out/r3v_templ.txt 2>&1
cat gpurun_out/r3v_templ.txt
timeout 500 python -m pytest tests -m gpu -q --timeout 100 --timeout-method=thread -x -k "stochastic or sampling or fast" -p no:cacheprovider > gpurun_out/r3v_tests.log 2>&1; tail -3 gpurun_out/r3v_tests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r3v_bench_headline.json 2> gpurun_out/r3v_bench_headline.err; head -c 250 gpurun_out/r3v_bench_headline.json; echo
