"""Development aid: dynamic SASS opcode mix of one kernel from an .ncu-rep captured with --import-source on.
Usage: python scripts/ncu_instmix.py report.ncu-rep kernel_regex [top]"""
import csv, io, subprocess, sys
from collections import Counter
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "-k", "regex:" + sys.argv[2],
                      "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
c, cs = Counter(), Counter()
tot = tots = 0
for r in rows[2:]:
    if len(r) <= iE or not r[iE].isdigit():
        continue
    toks = r[iS].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0].rstrip(";")
    c[op] += int(r[iE]); cs[op] += int(r[iSm]); tot += int(r[iE]); tots += int(r[iSm])
print("total warp instructions", tot, "samples", tots)
for op, n in c.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 20):
    print(f"{op:12s} {n:14d} {100 * n / tot:5.1f}%   samples {100 * cs[op] / max(tots, 1):5.1f}%")
