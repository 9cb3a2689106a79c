cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/gpu_train_pending.py -m gpu -q --timeout 180 --timeout-method=thread -p no:cacheprovider -k "adam or train_step" > gpurun_out/r5d_train_tests.log 2>&1; tail -30 gpurun_out/r5d_train_tests.log
