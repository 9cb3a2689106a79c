cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
timeout 120 python tools/time_variant.py libvrpb200_notok.so tc6 14208 greedy ups 2>&1 | tail -1
UPS=1 timeout 200 python tools/phase_clocks.py vrptw100 14208 tc6 48 2>&1 | tail -50
} > gpurun_out/r2r_clocks.txt 2>&1
cat gpurun_out/r2r_clocks.txt
