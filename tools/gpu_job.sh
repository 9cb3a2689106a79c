cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 100 python tools/smoke_tc5.py > gpurun_out/r2b_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r2b_smoke.txt
cat gpurun_out/r2b_smoke.txt
if grep -q "smoke rc=0" gpurun_out/r2b_smoke.txt; then
  UPS=1 timeout 300 python tools/phase_clocks.py vrptw100 14208 tc5 48 > gpurun_out/r2b_phase_clocks.txt 2>&1; cat gpurun_out/r2b_phase_clocks.txt
  timeout 900 python -m pytest tests/test_gpu_rollout.py -m gpu -q --timeout 300 -k "tc5" -p no:cacheprovider > gpurun_out/r2b_tests.log 2>&1; tail -15 gpurun_out/r2b_tests.log
fi
