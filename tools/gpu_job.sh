cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r3k_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3k_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r3k_ncu_bench.log 2>&1
tail -2 gpurun_out/r3k_ncu_bench.log | head -c 300; echo
python tools/time_variant.py libvrpb200.so tc4x 14208 greedy ups > gpurun_out/r3k_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rollout4 -s 2 -c 1 -o gpurun_out/r3k_prof python tools/time_variant.py libvrpb200.so tc4x 14208 greedy ups > gpurun_out/r3k_ncu.log 2>&1
tail -2 gpurun_out/r3k_ncu.log; ls -la gpurun_out/r3k_*
