cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
timeout 100 python tools/time_variant.py libvrpb200.so auto 128 greedy uniform vrp10 0 2>&1 | tail -1
timeout 100 python tools/time_variant.py libvrpb200.so auto 2000 greedy uniform vrptw20 0 2>&1 | tail -1
timeout 100 python tools/time_variant.py libvrpb200.so tc4x 2000 greedy uniform vrptw20 16 2>&1 | tail -1
timeout 100 python tools/time_variant.py libvrpb200.so tc5 8192 stochastic uniform vrptw100 16 2>&1 | tail -1
timeout 100 python tools/time_variant.py libvrpb200.so tc5 8192 stochastic uniform vrptw100 32 2>&1 | tail -1
} > gpurun_out/r3d_small.txt 2>&1
cat gpurun_out/r3d_small.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -x -p no:cacheprovider > gpurun_out/r3d_tests.log 2>&1; tail -5 gpurun_out/r3d_tests.log
