cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 40 python tools/time_variant.py libvrpb200.so tc4x 14208 greedy ups > gpurun_out/r3o_atomic.txt 2>&1 || { echo HUNG; tail -3 gpurun_out/r3o_atomic.txt; exit 0; }
{
timeout 40 python tools/time_variant.py libvrpb200.so auto 8192 stochastic uniform vrptw100 0 2>&1 | tail -1
timeout 40 python tools/time_variant.py libvrpb200.so auto 4096 greedy uniform vrptw20 0 2>&1 | tail -1
timeout 40 python tools/time_variant.py libvrpb200.so auto 128 greedy uniform vrp10 0 2>&1 | tail -1
UPS=1 timeout 60 python tools/phase_clocks.py vrptw100 14208 tc4x 48 2>&1 | tail -32
} >> gpurun_out/r3o_atomic.txt 2>&1
cat gpurun_out/r3o_atomic.txt
timeout 400 python -m pytest tests/test_gpu_rollout.py -m gpu -q --timeout 60 --timeout-method=thread -x -k "tc4x or auto or fast or rows_per" -p no:cacheprovider > gpurun_out/r3o_tests.log 2>&1; tail -4 gpurun_out/r3o_tests.log
