cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 40 python __graft_entry__.py smoke 2>&1 | tail -1
