cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_host_api.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -k "glimpse or ups_old or upsold or min_trucks" > gpurun_out/r2n_tests.log 2>&1; tail -40 gpurun_out/r2n_tests.log
