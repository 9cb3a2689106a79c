cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
for rb in 32 48; do
timeout 100 python tools/time_variant.py libvrpb200.so tc5 8192 stochastic uniform vrptw100 $rb 2>&1 | tail -1
timeout 100 python tools/time_variant.py libvrpb200.so tc5 4096 greedy uniform vrptw20 $rb 2>&1 | tail -1
done
timeout 100 python tools/time_variant.py libvrpb200.so auto 8192 stochastic uniform vrptw100 0 2>&1 | tail -1
timeout 100 python tools/time_variant.py libvrpb200.so auto 4096 greedy uniform vrptw20 0 2>&1 | tail -1
timeout 100 python tools/time_variant.py libvrpb200.so auto 128 greedy uniform vrp10 0 2>&1 | tail -1
timeout 100 python tools/time_variant.py libvrpb200.so auto 2000 greedy uniform vrptw20 0 2>&1 | tail -1
timeout 100 python tools/time_variant.py libvrpb200.so tc4x 2000 greedy uniform vrptw20 16 2>&1 | tail -1
} > gpurun_out/r2u_rb_sweep2.txt 2>&1
cat gpurun_out/r2u_rb_sweep2.txt
python bench.py --all-workloads --steps 5 --warmup 3 > gpurun_out/r2u_bench_all.json 2> gpurun_out/r2u_bench_all.err; tail -3 gpurun_out/r2u_bench_all.err
