cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
timeout 200 python tools/time_variant.py libvrpb200.so tc4x 14208 greedy ups 2>&1 | tail -1
timeout 200 python tools/time_variant.py libvrpb200_nothr.so tc4x 14208 greedy ups 2>&1 | tail -1
timeout 200 python tools/time_variant.py libvrpb200.so tc5 14208 greedy ups 2>&1 | tail -1
timeout 200 python tools/time_variant.py libvrpb200.so tc4x 8192 stochastic uniform 2>&1 | tail -1
timeout 200 python tools/time_variant.py libvrpb200.so tc5 8192 stochastic uniform 2>&1 | tail -1
timeout 200 python tools/time_variant.py libvrpb200.so tc4x 65536 stochastic uniform 2>&1 | tail -1
timeout 200 python tools/time_variant.py libvrpb200.so tc5 65536 stochastic uniform 2>&1 | tail -1
} > gpurun_out/r2m_variants.txt 2>&1
cat gpurun_out/r2m_variants.txt
timeout 600 python -m pytest tests/test_gpu_rollout.py -m gpu -q --timeout 120 --timeout-method=thread -x -k "tc5" -p no:cacheprovider > gpurun_out/r2m_tests.log 2>&1; tail -5 gpurun_out/r2m_tests.log
