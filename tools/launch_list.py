"""Launch list of one epoch from `ncu --metrics gpu__time_duration.sum --csv` (profiles/*_launches_*.txt):
python tools/launch_list.py launches.csv FIRST_KERNEL_OF_EPOCH EPOCH_INDEX > profiles/....txt
The epoch starts at the EPOCH_INDEX-th (0-based) launch of pack_weights_kernel that follows an optimizer_kernel (or the
first one) and ends with the next optimizer_kernel."""
import csv, collections, io, re, sys

txt = open(sys.argv[1]).read()
rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
rows = [r for r in rows if r.get("Metric Name") == "gpu__time_duration.sum"]
short = lambda n: re.sub(r"\(.*", "", n.replace("void ", "").replace("gatx::", "").replace("<unnamed>::", "").replace("unnamed>::", ""))
names = [short(r["Kernel Name"]) for r in rows]
ms = [float(r["Metric Value"]) * (1e-6 if r["Metric Unit"] == "ns" else 1e-3 if r["Metric Unit"] in ("us", "usecond") else 1.0) for r in rows]
ends = [i for i, n in enumerate(names) if n.startswith("optimizer_kernel")]
k = int(sys.argv[2]) if len(sys.argv) > 2 else len(ends) - 1
lo = ends[k - 1] + 1 if k > 0 else next(i for i, n in enumerate(names) if n.startswith("pack_weights"))
hi = ends[k] + 1
tot = sum(ms[lo:hi])
print("# total %.2f ms, %d launches (epoch %d of the capture; cold-cache, serialised: compare shares, not absolutes)" % (tot, hi - lo, k))
for n, t in zip(names[lo:hi], ms[lo:hi]):
    print("%9.3f ms  %5.1f%%  %s" % (t, 100 * t / tot, n))
agg = collections.OrderedDict()
for n, t in zip(names[lo:hi], ms[lo:hi]):
    a = agg.setdefault(n, [0.0, 0])
    a[0] += t
    a[1] += 1
print("# by kernel")
for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%9.3f ms  %5.1f%%  x%d  %s" % (t, 100 * t / tot, c, n))
