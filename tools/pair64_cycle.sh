# One validation cycle for the 64-float pair kernels (gpurun -- bash tools/pair64_cycle.sh): parity tests that reach them,
# the arxiv bench line + launch list + one ncu --set full capture (summarised on the box), and a short products line.
T=${1:-r2_pair64}
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out
python -m pytest tests -m gpu -q -k "pair64 or chunk_sizes or forward_backward_parity or arxiv or op_edge or baseline_config" > $O/${T}_tests_full.log 2>&1; (grep -m 12 -E "^(E  |FAILED|ERROR)" $O/${T}_tests_full.log; tail -3 $O/${T}_tests_full.log) > $O/${T}_tests.log; cat $O/${T}_tests.log; rm -f $O/${T}_tests_full.log
python bench.py --workload arxiv --steps 20 --no-same-config --no-cpu-baseline > $O/${T}_bench_arxiv.json 2> $O/${T}_bench_arxiv.err; echo arxiv rc=$?
python bench.py --steps 5 --warmup 3 --no-same-config --no-cpu-baseline > $O/${T}_bench_products.json 2> $O/${T}_bench_products.err; echo products rc=$?
A="python bench.py --workload arxiv --steps 2 --warmup 3 --no-cpu-baseline --no-same-config"
GATX_CUDA_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_arxiv.csv $A > $O/ncu_a.log 2>&1; echo rc=$?
python tools/launch_list.py $O/${T}_launches_arxiv.csv > $O/${T}_launches_arxiv_epoch.txt 2>&1
GATX_CUDA_GRAPH=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'edge_.*_pair_kernel' -s 3 -c 3 -f -o /tmp/${T}_pair $A > $O/ncu_pair.log 2>&1; echo rc=$?
python tools/ncu_summary.py /tmp/${T}_pair.ncu-rep > $O/${T}_ncu_pair_kernels_arxiv.txt 2>&1
rm -f $O/${T}_launches_arxiv.csv
du -sh $O
