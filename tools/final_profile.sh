# Round-end evidence on one B200 box (gpurun -- bash tools/final_profile.sh [tag]): GPU tests, the default bench line, the ncu
# launch list of one epoch (products and arxiv) and ncu --set full captures of the streaming / pair edge kernels and the
# tcgen05 GEMMs.  Summaries: tools/launch_list.py, tools/ncu_summary.py -> profiles/.
T=${1:-r2}
cd ${GRAFT_REPO_ROOT:-.}
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${T}_tests_final.log; cat gpurun_out/${T}_tests_final.log
timeout 900 python bench.py > gpurun_out/${T}_bench_final.json 2> gpurun_out/${T}_bench_final.err; echo bench rc=$?
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-same-config"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_products.csv $B > gpurun_out/ncu_l.log 2>&1; echo rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:edge_.*_stream_kernel -s 6 -c 4 -f -o gpurun_out/${T}_prof_stream $B > gpurun_out/ncu_s.log 2>&1; echo rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:edge_.*_pair_kernel -s 3 -c 3 -f -o gpurun_out/${T}_prof_pair $B > gpurun_out/ncu_p.log 2>&1; echo rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 13 -c 13 -f -o gpurun_out/${T}_prof_gemm $B > gpurun_out/ncu_g.log 2>&1; echo rc=$?
A="python bench.py --workload arxiv --steps 2 --warmup 3 --no-cpu-baseline --no-same-config"
GATX_CUDA_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_arxiv.csv $A > gpurun_out/ncu_a.log 2>&1; echo rc=$?
python bench.py --workload arxiv --steps 20 --no-same-config --no-cpu-baseline > gpurun_out/${T}_bench_arxiv.json 2>/dev/null
