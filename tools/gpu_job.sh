cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 40 python tools/time_variant.py libvrpb200.so auto 8192 stochastic uniform vrptw100 0 > gpurun_out/r3i_samp.txt 2>&1 || { echo HUNG; tail -3 gpurun_out/r3i_samp.txt; exit 0; }
timeout 40 python tools/time_variant.py libvrpb200.so auto 65536 stochastic uniform vrptw100 0 >> gpurun_out/r3i_samp.txt 2>&1
cat gpurun_out/r3i_samp.txt
timeout 300 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_host_api.py -m gpu -q --timeout 60 --timeout-method=thread -x -k "stochastic or critic or train_step" -p no:cacheprovider > gpurun_out/r3i_tests.log 2>&1; tail -4 gpurun_out/r3i_tests.log
