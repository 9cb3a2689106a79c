cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 40 python tools/time_variant.py libvrpb200.so tc4x 14208 greedy ups > gpurun_out/r3w_templ.txt 2>&1 || { echo HUNG; tail -3 gpurun_out/r3w_templ.txt; exit 0; }
timeout 40 python tools/time_variant.py libvrpb200.so auto 8192 stochastic uniform vrptw100 0 >> gpurun_out/r3w_templ.txt 2>&1
timeout 40 python tools/time_variant.py libvrpb200.so auto 4096 greedy uniform vrptw20 0 >> gpurun_out/r3w_templ.txt 2>&1
cat gpurun_out/r3w_templ.txt
timeout 500 python -m pytest tests -m gpu -q --timeout 100 --timeout-method=thread -x -k "stochastic or sampling or fast" -p no:cacheprovider > gpurun_out/r3w_tests.log 2>&1; tail -3 gpurun_out/r3w_tests.log
UPS=1 timeout 60 python tools/phase_clocks.py vrptw100 14208 tc4x 48 2>&1 | grep "end-of-step\|arg-max\|total\|near-tie"; STOCHASTIC=1 timeout 60 python tools/phase_clocks.py vrptw100 8192 tc4x 32 2>&1 | grep "rows fast\|total"
