#!/usr/bin/env python3
"""Static instruction profile of a kernel: SASS instructions per source line (nvdisasm --print-line-info on the
cubin of a -lineinfo object), optionally restricted to an address range (the main loop).  No GPU needed: this is
how the experimental cuts at the end of round 1 were chosen and sized when the round's GPU budget was spent.

  cuobjdump -xelf all bliss-rs_b200/csrc/spectral.o && nvdisasm --print-line-info spectral.sm_100a.cubin > lines.txt
  python scripts/static_sass.py lines.txt pvoc512_kernelILb1ELb0ELb0ELb0 [0x1a20 0x6990]   # address range = the frame loop
"""
import re
import sys
from collections import Counter, defaultdict


def parse(path, sub):
    txt = open(path).read().split("\n")
    start = [i for i, l in enumerate(txt) if l.startswith(".text.") and sub in l][0]
    cur, ins = None, []
    for l in txt[start + 1:]:
        if l.startswith("//-----"):
            break
        m = re.match(r'\s*//## File "(.*?)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            t = m.group(2).split()
            if t[0].startswith("@"):
                t = t[1:]
            ins.append((int(m.group(1), 16), t[0].split(".")[0], " ".join(t), cur))
    return ins


def main():
    path, sub = sys.argv[1], sys.argv[2]
    lo, hi = (int(sys.argv[3], 16), int(sys.argv[4], 16)) if len(sys.argv) > 4 else (0, 1 << 30)
    ins = [x for x in parse(path, sub) if lo <= x[0] <= hi]
    per, perop = Counter(), defaultdict(Counter)
    for a, op, t, cur in ins:
        per[cur] += 1
        perop[cur][op] += 1
    ops = Counter(op for _, op, _, _ in ins)
    print("%d instructions in [%#x, %#x]" % (len(ins), lo, min(hi, ins[-1][0])))
    print("opcodes:", ", ".join("%s %d" % kv for kv in ops.most_common(24)))
    print()
    print("| instructions | file:line | top opcodes |")
    print("|---|---|---|")
    for (f, ln), c in sorted(per.items(), key=lambda x: -x[1])[:32]:
        print("| %d | %s:%d | %s |" % (c, f, ln, ", ".join("%s %d" % kv for kv in perop[(f, ln)].most_common(3))))


if __name__ == "__main__":
    main()
