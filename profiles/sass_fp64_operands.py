"""FP64 issue-cycle estimate of a kernel's RK4 step loop from its SASS: usage sass_fp64_operands.py <obj> <kernel> [min_len max_len]
A DFMA/DMUL/DADD with <= 2 distinct register operands issues every 2 clocks per scheduler, one with 3 distinct register
operands every 3 (measured: profiles/r1l/fp64_operands.txt, 64.0 vs 42.6 DFMA lanes/clk/SM)."""
import re, subprocess, sys
from collections import Counter
obj, kern = sys.argv[1], sys.argv[2]
lo_len, hi_len = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (150, 100000)
out = subprocess.check_output(["cuobjdump", "-sass", "-fun", kern, obj], text=True)
rows = []
for l in out.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        rows.append((int(m.group(1), 16), m.group(2).strip()))
# the step loop = the INNERMOST backward branch whose body holds FP64 work (shortest loop within the length window)
best = None
for a, t in rows:
    m = re.search(r"BRA(\.U)?\s+(!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
    if m and int(m.group(3), 16) < a:
        n = (a - int(m.group(3), 16)) // 16 + 1
        if lo_len <= n <= hi_len and (best is None or n < best[2]):
            best = (int(m.group(3), 16), a, n)
lo, hi, n = best
c, cyc, other = Counter(), 0, 0
for a, t in rows:
    if not (lo <= a <= hi):
        continue
    m = re.match(r"(@!?U?P\d\s+)?(DFMA|DMUL|DADD|DSETP)\S*\s+(.*)", t)
    if not m:
        other += 1
        continue
    srcs = [o.strip() for o in m.group(3).split(",")][1:]
    regs = {re.sub(r"[-|~]", "", o).split(".")[0] for o in srcs if re.match(r"[-|~]*R\d+", o)}
    c[(m.group(2), len(regs))] += 1
    cyc += 3 if len(regs) >= 3 else 2
nf = sum(c.values())
print("%s: step loop %s-%s, %d instructions: %d FP64 (%d with 3 distinct register operands), %d other" % (kern, hex(lo), hex(hi), n, nf, sum(v for k, v in c.items() if k[1] >= 3), other))
print("  FP64 issue cycles per warp-step >= %d (2 per instruction would be %d); by (op, distinct regs): %s" % (cyc, 2 * nf, sorted(c.items())))
