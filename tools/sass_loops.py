#!/usr/bin/env python
"""Opcode histogram of the loops of one kernel in a cuobjdump -sass listing.

    cuobjdump -sass -fun '<mangled>' lib.so | python tools/sass_loops.py [min_len]

A loop = a backward branch; prints, for every loop body longer than min_len instructions, the
instruction count and the opcode histogram (predicated instructions are counted under their opcode)."""
import collections
import re
import sys

min_len = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ins = []
pat = re.compile(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);")
for line in sys.stdin:
    m = pat.match(line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, text) in enumerate(ins):
    m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", text)
    if not m:
        continue
    tgt = int(m.group(1), 16)
    if tgt >= a or tgt not in addr_index:
        continue
    j = addr_index[tgt]
    body = ins[j:i + 1]
    if len(body) < min_len:
        continue
    hist = collections.Counter()
    for _, t in body:
        t = re.sub(r"^@!?U?P\w+\s+", "", t)
        hist[t.split()[0].split(".")[0]] += 1
    groups = {"fp64": ("DADD", "DMUL", "DFMA", "DSETP"), "shuffle": ("SHFL",), "smem": ("LDS", "STS", "LDGSTS", "LDGDEPBAR", "DEPBAR"),
              "global": ("LDG", "STG", "LD", "ST"), "select": ("FSEL", "SEL", "ISETP", "FSETP", "PLOP3", "LOP3"),
              "int/addr": ("IMAD", "IADD3", "VIADD", "LEA", "SHF", "MOV", "VIADDMNMX", "IABS", "UMOV", "UIADD3", "ULEA", "UIMAD", "CS2R")}
    print("loop 0x%04x..0x%04x: %d instructions" % (tgt, a, len(body)))
    for g, ops in groups.items():
        print("   %-9s %4d   %s" % (g, sum(hist[o] for o in ops), " ".join("%s=%d" % (o, hist[o]) for o in ops if hist[o])))
    rest = {o: c for o, c in hist.items() if not any(o in ops for ops in groups.values())}
    print("   other     %4d   %s" % (sum(rest.values()), " ".join("%s=%d" % kv for kv in sorted(rest.items(), key=lambda kv: -kv[1]))))
