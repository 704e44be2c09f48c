cd ${GRAFT_REPO_ROOT:-.}
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/tests_final.log; cat gpurun_out/tests_final.log
timeout 300 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo bench rc=$?
bash tools/ab_packed_fp32.sh 2>&1 | tail -4
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:edge_.*_stream_kernel -s 6 -c 4 -f -o gpurun_out/prof_stream_final2 $B > gpurun_out/ncu_s2.log 2>&1; echo rc=$?
timeout 200 ncu --set full --clock-control none --import-source on -k regex:edge_.*_pair_kernel -s 3 -c 3 -f -o gpurun_out/prof_pair_final2 $B > gpurun_out/ncu_p2.log 2>&1; echo rc=$?
