#!/usr/bin/env python3
"""tools/sass_loops.py -- opcode census of the biggest loop of every warp role of a strip-kernel instantiation.
Usage: cuobjdump -sass -fun <mangled name> build/kernel_strip.cu.o | python tools/sass_loops.py
Roles are the SASS regions between the PMTRIG markers (pmevent 1..4 = P, A, B, C); a loop is a backward branch and its
target; the report lists, per role, the largest loops with their instruction counts, the packed (two-FMA-pipe-cycle)
instructions among them and the opcode histogram."""
import re, sys
from collections import Counter

PACKED = ("FFMA2", "FADD2", "FMUL2")

def main():
    ins = []
    for l in sys.stdin:
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;', l)
        if m:
            txt = re.sub(r'^@!?U?P\d+\s+', '', m.group(2))
            ins.append((int(m.group(1), 16), txt))
    marks = [i for i, (_, t) in enumerate(ins) if t.startswith("PMTRIG")] + [len(ins)]
    names = ["P", "A", "B", "C"]
    for r in range(len(marks) - 1):
        reg = ins[marks[r]:marks[r + 1]]
        lo, hi = reg[0][0], reg[-1][0]
        loops = []
        for pc, t in reg:
            m = re.match(r'BRA(?:\.\w+)*\s+(?:\w+,\s*)?0x([0-9a-f]+)', t)
            if m:
                tgt = int(m.group(1), 16)
                if lo <= tgt < pc:
                    loops.append((pc - tgt, tgt, pc))
        loops.sort(reverse=True)
        print("role %s: %d instructions" % (names[r] if r < 4 else "?", len(reg)))
        for span, tgt, pc in loops[:int(sys.argv[1]) if len(sys.argv) > 1 else 2]:
            body = [t.split()[0].split('.')[0] for a, t in reg if tgt <= a <= pc]
            c = Counter(body)
            npk = sum(c[k] for k in PACKED)
            print("  loop %04x..%04x: %d instr, %d packed (%d issue+pipe slots): %s" % (tgt, pc, len(body), npk, len(body) + npk,
                  ", ".join("%s %d" % kv for kv in c.most_common(24))))

if __name__ == "__main__":
    main()
