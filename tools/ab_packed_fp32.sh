#!/bin/bash
# Same-box A/B of the packed fp32x2 edge kernels (FFMA2 / FADD2 / FMUL2) against the scalar build of the same arithmetic:
#   GATX_VARIANT=scalar GATX_EXTRA_FLAGS=-DGATX_SCALAR_FP32 python graph-attention-network-gatv2-_b200/build.py   (here, once)
#   gpurun -- bash tools/ab_packed_fp32.sh        -> gpurun_out/ab_packed_*.json, one summary line per run
cd ${GRAFT_REPO_ROOT:-.}
P=graph-attention-network-gatv2-_b200
for round in 1 2; do
  for v in scalar packed; do
    lib=$P/libgatx.so; [ $v = scalar ] && lib=$P/libgatx_scalar.so
    GATX_LIB=$PWD/$lib python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_packed_${v}_${round}.json 2> gpurun_out/ab_packed_${v}_${round}.err
    python - <<PY
import json
d = json.load(open("gpurun_out/ab_packed_${v}_${round}.json"))
k = {(x["kernel"][5:12], x["layer"]): round(x["ms"], 2) for x in d["roofline"]["all_edge_kernels"]}
print("${v} ${round}: epoch %.2f ms  edge_fwd %.2f edge_bwd %.2f  sm %.0f MHz  loss %.9f  %s" % (d["ms_per_step"], d["phase_ms_per_epoch"]["edge_fwd"], d["phase_ms_per_epoch"]["edge_bwd"], d["clocks"]["sm_mhz"], d["final_loss"], k))
PY
  done
done
