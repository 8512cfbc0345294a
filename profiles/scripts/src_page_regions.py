#!/usr/bin/env python3
"""ncu source page (--page source --csv --print-source sass) of one kernel -> executed warp instructions and stall samples
per code region (kernel body / each out-of-line function, split at RET) and per opcode.
usage: src_page_regions.py page.csv [--body-ops]"""
import csv, sys, re, collections

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
ins = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    text = r[col["Source"]].strip()
    t = re.sub(r'^@!?U?P\d+\s+', '', text)
    op = t.split()[0].rstrip(';')
    ins.append((int(r[0], 16), op, text, int(r[col["Instructions Executed"]]), int(r[col["# Samples"]]),
                int(r[col["stall_wait"]]), int(r[col["stall_math"]]), int(r[col["stall_long_sb"]])))
base = ins[0][0]
regions, cur = [], []
for i in ins:
    cur.append(i)
    if i[1].startswith("RET") or i[1].startswith("EXIT") and False:
        regions.append(cur); cur = []
if cur:
    regions.append(cur)
tot = sum(i[3] for i in ins); tots = sum(i[4] for i in ins)
print(f"total executed warp instructions {tot:.4g}, samples {tots}")
for k, reg in enumerate(regions):
    ex = sum(i[3] for i in reg); sm = sum(i[4] for i in reg)
    wide = sum(i[3] for i in reg if i[1].startswith("IMAD.WIDE"))
    print(f"region {k}: 0x{reg[0][0]-base:05x}..0x{reg[-1][0]-base:05x} static {len(reg):5d} executed {100.0*ex/tot:5.1f}% "
          f"samples {100.0*sm/tots:5.1f}%  IMAD.WIDE share {100.0*wide/max(ex,1):4.1f}%  calls~{reg[-1][3]:.3g}")
if "--body-ops" in sys.argv:
    for k, reg in enumerate(regions):
        ex = sum(i[3] for i in reg)
        if ex < 0.02 * tot:
            continue
        h = collections.Counter(); s = collections.Counter()
        for i in reg:
            h[i[1]] += i[3]; s[i[1]] += i[4]
        print(f"-- region {k} opcodes (share of ALL executed / of ALL samples)")
        for op, c in h.most_common(14):
            print(f"   {op:24s} {100.0*c/tot:5.2f}%  {100.0*s[op]/tots:5.2f}%")
