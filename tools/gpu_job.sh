cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rollout.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -k "whole_batch or stochastic" > gpurun_out/r2p_tests.log 2>&1; tail -40 gpurun_out/r2p_tests.log
