cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/gpu_train_pending.py tests/test_gpu_train.py -m gpu -q --timeout 180 --timeout-method=thread -p no:cacheprovider > gpurun_out/r9_train_tests.log 2>&1; tail -40 gpurun_out/r9_train_tests.log
