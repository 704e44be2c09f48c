set -x
cd $GRAFT_REPO_ROOT
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo rc=$?
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv $B > gpurun_out/ncu_l.log 2>&1; echo rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:edge_.*_stream_kernel -s 6 -c 4 -f -o gpurun_out/prof_stream_final $B > gpurun_out/ncu_s.log 2>&1; echo rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:edge_.*_pair_kernel -s 3 -c 3 -f -o gpurun_out/prof_pair_final $B > gpurun_out/ncu_p.log 2>&1; echo rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 14 -c 14 -f -o gpurun_out/prof_gemm_final $B > gpurun_out/ncu_g.log 2>&1; echo rc=$?
ls -la gpurun_out/*final*
