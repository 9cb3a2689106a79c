cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 180 --timeout-method=thread -p no:cacheprovider > gpurun_out/r8_train_tests.log 2>&1; tail -3 gpurun_out/r8_train_tests.log
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -1
VRPB200_TRAIN_CPU=0 timeout 100 python tools/time_train_step.py vrptw100 128 2>&1 | tail -1
