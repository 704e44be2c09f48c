"""Static SASS statistics per kernel: python tools/sass_count.py <regex> -> instructions, S2R/S2UR count, and the span
(in instructions) of every backward branch (a loop body) that contains an LDS."""
import re, subprocess, sys
so = "graph-attention-network-gatv2-_b200/libgatx.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(sys.argv[1])
for blk in txt.split("Function : ")[1:]:
    name = blk.split("\n", 1)[0]
    if not pat.search(name):
        continue
    ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", blk)
    addr = {int(a, 16): i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.\w+)* (?:\w+, )?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt in addr and addr[tgt] < i:
                body = ins[addr[tgt]:i + 1]
                if any("LDS" in b for _, b in body):
                    loops.append((len(body), sum("S2R" in b or "S2UR" in b for _, b in body)))
    short = re.sub(r"^_ZN4gatx\d+_GLOBAL__N__\w+?_cu_[0-9a-f]+\d\d", "", name)[:60]
    print("%-62s instr %5d  S2R/S2UR %2d  loops(len,s2r) %s" % (short, len(ins), sum("S2R" in t or "S2UR" in t for _, t in ins), loops))
