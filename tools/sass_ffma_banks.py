"""Static estimate of FFMA register-bank pressure in a kernel's SASS (no GPU needed).

Model (B300_MICROARCH.md "RF banking"): a 3-source FFMA needs max(1, #distinct fresh even regs, #distinct fresh odd
regs) dispatch cycles; an operand is not fresh when the previous instruction carried `.reuse` in the same slot for the
same register.  usage: cuobjdump -sass x.o | python tools/sass_ffma_banks.py <kernel-substring>"""
import re
import sys

pat = re.compile(r"^\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?(\S+)\s+(.*?);")
want = sys.argv[1] if len(sys.argv) > 1 else ""
inside = False
prev_reuse = {}
cyc = n = 0
other = 0
hist = {}
for line in sys.stdin:
    if "Function :" in line:
        inside = want in line
        continue
    if not inside:
        continue
    m = pat.match(line)
    if not m:
        continue
    op, args = m.group(2), m.group(3)
    if not op.startswith("FFMA"):
        prev_reuse = {}
        other += 1
        continue
    srcs = [a.strip() for a in args.split(",")][1:4]
    fresh_even, fresh_odd = set(), set()
    new_reuse = {}
    for slot, s in enumerate(srcs):
        r = re.match(r"-?\|?(R\d+)", s)
        if not r:
            continue
        reg = int(r.group(1)[1:])
        if ".reuse" in s:
            new_reuse[slot] = reg
        if prev_reuse.get(slot) == reg:
            continue
        (fresh_even if reg % 2 == 0 else fresh_odd).add(reg)
    c = max(1, len(fresh_even), len(fresh_odd))
    hist[c] = hist.get(c, 0) + 1
    cyc += c
    n += 1
    prev_reuse = new_reuse
print("FFMA %d, est. dispatch cycles %d (%.3f per FFMA), other instructions %d, histogram %s" % (n, cyc, cyc / max(n, 1), other, hist))
