# Round-end evidence on one B200 box (gpurun -- bash tools/final_profile.sh [tag] [tests]): the default bench line, the ncu
# launch list of one epoch (products and arxiv) and ncu --set full captures of the streaming / pair edge kernels and the
# tcgen05 GEMMs, summarised ON the box (tools/ncu_summary.py, tools/launch_list.py): only the text summaries travel back
# (the .ncu-rep files exceed gpurun's 64 MiB return limit).  Pass "tests" as the second argument to run the GPU tests first.
T=${1:-r2}
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out
if [ "$2" = "tests" ]; then python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/${T}_tests_final.log; cat $O/${T}_tests_final.log; fi
timeout 900 python bench.py > $O/${T}_bench_final.json 2> $O/${T}_bench_final.err; echo bench rc=$?
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-same-config"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_products.csv $B > $O/ncu_l.log 2>&1; echo rc=$?
python tools/launch_list.py $O/${T}_launches_products.csv > $O/${T}_launches_products_epoch.txt 2>&1
cap() {  # name, kernel regex, skip, count
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o /tmp/${T}_$1 $B > $O/ncu_$1.log 2>&1; echo rc=$?
  python tools/ncu_summary.py /tmp/${T}_$1.ncu-rep > $O/${T}_ncu_$1.txt 2>&1
}
cap stream 'edge_.*_stream_kernel' 6 4
cap pair 'edge_.*_pair_kernel' 3 3
cap gemm 'gemm_tf32' 13 13
A="python bench.py --workload arxiv --steps 2 --warmup 3 --no-cpu-baseline --no-same-config"
GATX_CUDA_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_arxiv.csv $A > $O/ncu_a.log 2>&1; echo rc=$?
python tools/launch_list.py $O/${T}_launches_arxiv.csv > $O/${T}_launches_arxiv_epoch.txt 2>&1
python bench.py --workload arxiv --steps 20 --no-same-config --no-cpu-baseline > $O/${T}_bench_arxiv.json 2>/dev/null
du -sh $O
