cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_host_api.py -m gpu -q --timeout 120 --timeout-method=thread -x -p no:cacheprovider > gpurun_out/r3y_tests.log 2>&1; tail -5 gpurun_out/r3y_tests.log
