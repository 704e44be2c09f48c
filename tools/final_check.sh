# Round-end check on one B200 (gpurun -- bash tools/final_check.sh [tag]): every GPU test, then the default bench line.
T=${1:-r2_final}
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out
python -m pytest tests -m gpu -q > $O/${T}_tests_full.log 2>&1
(grep -m 20 -E "^(E  |FAILED|ERROR)" $O/${T}_tests_full.log | cut -c1-300; tail -3 $O/${T}_tests_full.log) > $O/${T}_tests.log; cat $O/${T}_tests.log; rm -f $O/${T}_tests_full.log
timeout 900 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo bench rc=$?
python -c "
import json; j=json.load(open('$O/${T}_bench.json'))
print(j['ms_per_step'], j['value'], j['e2e'], j['roofline']['frac'], j['roofline']['edge_passes_aggregate'], j.get('same_config'), j['clocks'])"
