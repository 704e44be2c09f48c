"""Per-unit SASS instruction mix from an `ncu --page source --csv` dump: python tools/sass_mix.py file.csv UNITS"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2])
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ei0 = hdr.index("Instructions Executed")
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[ei0].isdigit() and r[0].startswith("0x")]
ai, ei, si = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[ei]) for r in data)
print("static instructions", len(data), "executed", tot, "per unit %.1f" % (tot / units))
hot = [r for r in data if int(r[ei]) > 0.5 * units]
print("hot instructions", len(hot), "-> %.1f per unit" % (sum(int(r[ei]) for r in hot) / units))
def opc(s):
    t = s.split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
ops, st = collections.Counter(), collections.Counter()
for r in hot:
    ops[opc(r[ai])] += int(r[ei]) / units
    st[opc(r[ai])] += int(r[si])
print("mix:", ", ".join("%s %.1f" % kv for kv in ops.most_common(40)))
print("stall samples:", st.most_common(12))
