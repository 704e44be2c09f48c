# Round-end evidence on one B200 box (gpurun -- bash tools/final_profile.sh): GPU tests, the default bench line, the same-box
# A/B of packed vs scalar fp32 (needs libgatx_scalar.so, see tools/ab_packed_fp32.sh), the ncu launch list of one epoch and
# ncu --set full captures of the streaming / pair edge kernels and the tcgen05 GEMMs.  Summaries: tools/launch_list.py,
# tools/ncu_summary.py -> profiles/.
cd ${GRAFT_REPO_ROOT:-.}
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/tests_final.log; cat gpurun_out/tests_final.log
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo bench rc=$?
[ -f graph-attention-network-gatv2-_b200/libgatx_scalar.so ] && bash tools/ab_packed_fp32.sh 2>&1 | tail -4
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv $B > gpurun_out/ncu_l.log 2>&1; echo rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:edge_.*_stream_kernel -s 6 -c 4 -f -o gpurun_out/prof_stream_final $B > gpurun_out/ncu_s.log 2>&1; echo rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:edge_.*_pair_kernel -s 3 -c 3 -f -o gpurun_out/prof_pair_final $B > gpurun_out/ncu_p.log 2>&1; echo rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 14 -c 14 -f -o gpurun_out/prof_gemm_final $B > gpurun_out/ncu_g.log 2>&1; echo rc=$?
