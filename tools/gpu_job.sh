cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_host_api.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -k "critic or train_step or stepwise_matches" > gpurun_out/r2o_tests.log 2>&1; tail -40 gpurun_out/r2o_tests.log
