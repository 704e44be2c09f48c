# A/B of the pass-1 row staging (bulk copy vs register prefetch) and the race check of the 128-float pair kernels
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out
L=$PWD/graph-attention-network-gatv2-_b200
echo "== race check, 128-float last layer, default lib" > $O/ab_rowldg.log
python tools/debug_pair64.py race 4 128 2>&1 | tail -4 >> $O/ab_rowldg.log
echo "== race check, 128-float last layer, rowldg" >> $O/ab_rowldg.log
GATX_LIB=$L/libgatx_rowldg.so python tools/debug_pair64.py race 4 128 2>&1 | tail -4 >> $O/ab_rowldg.log
for v in "" _rowldg; do
  GATX_LIB=$L/libgatx$v.so python bench.py --steps 5 --warmup 3 --no-same-config --no-cpu-baseline > $O/ab_products$v.json 2>/dev/null
  GATX_LIB=$L/libgatx$v.so python bench.py --workload arxiv --steps 20 --no-same-config --no-cpu-baseline > $O/ab_arxiv$v.json 2>/dev/null
done
GATX_LIB=$L/libgatx_pairtma.so python bench.py --workload arxiv --steps 20 --no-same-config --no-cpu-baseline > $O/ab_arxiv_pairtma.json 2>/dev/null
python - <<'PY' >> gpurun_out/ab_rowldg.log
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        j = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    ks = j.get("roofline", {}).get("kernels") or j.get("kernels") or []
    print(f, "ms/step %.3f" % j["ms_per_step"], "sm_mhz", j.get("clocks", {}).get("sm_mhz"))
    for k in ks:
        if k["layer"] == 2: print("    ", k["kernel"], "%.3f ms %.0f GB/s" % (k["ms"], k["gbs"]))
PY
cat $O/ab_rowldg.log
