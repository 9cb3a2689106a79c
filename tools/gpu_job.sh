cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_host_api.py -m gpu -q --timeout 120 --timeout-method=thread -x -p no:cacheprovider -k "reward or min_trucks or ups_old" > gpurun_out/r5b_tests.log 2>&1; tail -5 gpurun_out/r5b_tests.log
