"""Static SASS mix of a kernel's hot loop (largest backward branch): usage sass_count.py <obj> <kernel>"""
import re, subprocess, sys
from collections import Counter
obj, kern = sys.argv[1], sys.argv[2]
out = subprocess.check_output(["cuobjdump", "-sass", "-fun", kern, obj], text=True)
ins = []
for l in out.split('\n'):
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
best = None
def has_lds(lo, hi): return any(lo <= a <= hi and t.startswith("LDS") for a, t in ins)
for a, t in ins:
    m = re.search(r"BRA (0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        # the fast path is the largest loop that reads the shared-memory sincos table and calls nothing
        if tgt < a and has_lds(tgt, a) and sum(1 for x, y in ins if tgt <= x <= a and (y.startswith("UMOV") or "CALL" in y)) < 8 and (best is None or a - tgt > best[1] - best[0]): best = (tgt, a)
if best is None:
    for a, t in ins:
        m = re.search(r"BRA (0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a and (best is None or a - int(m.group(1), 16) > best[1] - best[0]): best = (int(m.group(1), 16), a)
c = Counter()
for a, t in ins:
    if best[0] <= a <= best[1]:
        op = t.split()[0] if not t.startswith('@') else t.split()[1]
        c[op.split('.')[0]] += 1
n = sum(c.values()); dp = sum(v for k, v in c.items() if k in ('DFMA', 'DMUL', 'DADD', 'DSETP'))
print(f"{kern}: total static {len(ins)}; hot loop {hex(best[0])}-{hex(best[1])}: {n} instrs, {dp} fp64 ->", sorted(c.items(), key=lambda x: -x[1])[:12])
