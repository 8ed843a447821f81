"""Static SASS view of a kernel: every backward branch (loop) with its instruction mix, FP64 count and the issue clocks of
the model measured in profiles/r2a and r2u (FP64 = 2 clk, 3 with three distinct register operands — 2.3 when one of them
carries the .reuse flag, i.e. is the previous instruction's operand in the same slot —, LDS 2, everything else 1).
usage: sass_loops.py <obj|cubin|so> <kernel> [--dump lo hi]"""
import re, subprocess, sys
from collections import Counter
obj, kern = sys.argv[1], sys.argv[2]
out = subprocess.check_output(["cuobjdump", "-sass", "-fun", kern, obj], text=True)
ins = []
for l in out.split('\n'):
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
def opname(t):
    p = t.split()
    return (p[1] if p[0].startswith('@') else p[0]).split('.')[0]
def fp64(t): return opname(t) in ('DFMA', 'DMUL', 'DADD', 'DSETP')
def three_reg(t):
    if opname(t) != 'DFMA': return False
    ops = t.split(None, 2 if t.startswith('@') else 1)[-1]
    regs = set(re.findall(r"\bR(\d+)\b", ops.split(',', 1)[1])) if ',' in ops else set()
    return len(regs) >= 3 and 'c[' not in ops
def clocks(lo, hi, skip=()):
    n = f = t3 = t3r = lds = 0
    c = Counter()
    for a, t in ins:
        if lo <= a <= hi and not any(s0 <= a <= s1 for s0, s1 in skip):
            n += 1; c[opname(t)] += 1
            if fp64(t): f += 1
            if three_reg(t):
                t3 += 1
                if '.reuse' in t: t3r += 1
            if opname(t) == 'LDS' and not t.startswith('@!PT'): lds += 1
    return n, f, t3, c, t3r, lds
if len(sys.argv) > 3 and sys.argv[3] == '--dump':
    lo, hi = int(sys.argv[4], 16), int(sys.argv[5], 16)
    for a, t in ins:
        if lo <= a <= hi: print(hex(a), t)
    sys.exit(0)
print(f"{kern}: {len(ins)} instructions")
for a, t in ins:
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a and a - int(m.group(1), 16) > 64:
        lo = int(m.group(1), 16)
        n, f, t3, c, t3r, lds = clocks(lo, a)
        print(f"  loop {hex(lo)}..{hex(a)}: {n} instr, {f} FP64 ({t3} three-register, {t3r} of them .reuse), {lds} LDS, issue clocks {n + f + (t3 - t3r) + 0.3 * t3r + lds:.0f};  " +
              ", ".join(f"{k} {v}" for k, v in c.most_common(10)))
