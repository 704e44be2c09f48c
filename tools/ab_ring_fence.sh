# A/B for the next round (gpurun -- bash tools/ab_ring_fence.sh): does a fence.proxy.async before every per-edge refill of the
# bulk-copy rings (-DGATX_RING_FENCE, edge_stream.cu) cost epoch time?  Builds the variant on the box, runs both benches.
cd ${GRAFT_REPO_ROOT:-.}
L=$PWD/graph-attention-network-gatv2-_b200
(cd $L && GATX_VARIANT=ringfence GATX_EXTRA_FLAGS="-DGATX_RING_FENCE" python build.py > /dev/null)
for v in "" _ringfence; do
  GATX_LIB=$L/libgatx$v.so python bench.py --steps 10 --warmup 3 --no-same-config --no-cpu-baseline > gpurun_out/ab_ring$v.json 2>/dev/null
  python -c "
import json; j=json.load(open('gpurun_out/ab_ring$v.json')); print('$v', j['ms_per_step'], j['clocks']['sm_mhz'], [round(k['ms'],2) for k in j['roofline']['all_edge_kernels']])"
done
