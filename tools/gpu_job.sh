cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 180 --timeout-method=thread -p no:cacheprovider > gpurun_out/r7_train_tests.log 2>&1; tail -5 gpurun_out/r7_train_tests.log
rm -f gpurun_out/r7_train_timing.txt
for cfg in "vrptw100 128" "vrptw100 1024"; do VRPB200_TRAIN_CPU=0 timeout 100 python tools/time_train_step.py $cfg 2>&1 | tail -1 >> gpurun_out/r7_train_timing.txt; done; cat gpurun_out/r7_train_timing.txt
