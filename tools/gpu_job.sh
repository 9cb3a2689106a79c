cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 140 python -m pytest tests/test_gpu_host_api.py tests/test_gpu_train.py -m gpu -q -x --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/r10_host_train.log 2>&1; tail -5 gpurun_out/r10_host_train.log
