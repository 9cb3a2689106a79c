cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 180 --timeout-method=thread -p no:cacheprovider > gpurun_out/r5e_train_tests.log 2>&1; tail -40 gpurun_out/r5e_train_tests.log
timeout 200 python -m pytest tests/test_gpu_host_api.py -m gpu -q --timeout 120 --timeout-method=thread -x -p no:cacheprovider -k "train_step or critic" > gpurun_out/r5e_host.log 2>&1; tail -5 gpurun_out/r5e_host.log
