cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/r6_train_step.txt
for cfg in "vrptw100 128" "vrptw50 128" "vrp10 128"; do timeout 150 python tools/time_train_step.py $cfg 2>&1 | tail -2 >> gpurun_out/r6_train_step.txt; done; cat gpurun_out/r6_train_step.txt
