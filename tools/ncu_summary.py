"""Text summary of an `ncu --set full` report, one block per captured launch (the format of profiles/*_ncu_*.txt):
python tools/ncu_summary.py file.ncu-rep [file2.ncu-rep ...] > profiles/....txt"""
import csv, io, re, subprocess, sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
    "launch__shared_mem_per_block_dynamic",
]
STALL = re.compile(r"smsp__average_warps?_issue_stalled_(\w+)_per_issue_active\.ratio$")


def summarize(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out[out.index('"ID"'):])))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        name = r[col["Kernel Name"]].replace("void ", "").replace("gatx::", "")
        print("%-56s grid= %s" % (name[:56], r[col["Grid Size"]].strip("()").split(",")[0]))
        for m in METRICS:
            i = col.get(m)
            if i is not None and r[i] != "":
                print("    %-70s %s %s" % (m, r[i], units[i]))
        st = []
        for h, i in col.items():
            mm = STALL.search(h)
            if mm and r[i] not in ("", "0"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 0.15:
                    st.append("%s %.2f" % (mm.group(1), v))
        print("    stalls per issue: " + ", ".join(sorted(st)))


if __name__ == "__main__":
    for rep in sys.argv[1:]:
        summarize(rep)
