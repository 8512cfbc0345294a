#!/usr/bin/env python3
"""ncu SASS source page -> samples / executed instructions per 0x200-byte block of the kernel body, top stall reasons, and
the most-stalled single instructions.  usage: src_page_blocks.py page.csv [body_end_hex] [ntop]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
end = int(sys.argv[2], 16) if len(sys.argv) > 2 else 1 << 60
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 30
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]; col = {n: i for i, n in enumerate(hdr)}
ins = [r for r in rows[h + 1:] if len(r) >= len(hdr) and r[0].startswith('0x')]
base = int(ins[0][0], 16)
tot = sum(int(r[col['# Samples']]) for r in ins)
stall_cols = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
body = [r for r in ins if int(r[0], 16) - base < end]
blk = collections.OrderedDict()
for r in body:
    a = (int(r[0], 16) - base) // 0x200 * 0x200
    d = blk.setdefault(a, [0, 0, collections.Counter()])
    d[0] += int(r[col['# Samples']]); d[1] += int(r[col['Instructions Executed']])
    for n in stall_cols:
        d[2][n] += int(r[col[n]])
print("block  samples%  exec(M)  top stalls")
for a, d in blk.items():
    if d[0] < 0.003 * tot: continue
    top = ", ".join(f"{n[6:]} {100*v/tot:.1f}" for n, v in d[2].most_common(3))
    print(f"0x{a:05x} {100*d[0]/tot:6.2f}% {d[1]/1e6:8.1f}  {top}")
print("--- top instructions")
top = sorted(body, key=lambda r: -int(r[col['# Samples']]))[:ntop]
for r in sorted(top, key=lambda r: int(r[0], 16)):
    st = sorted(((int(r[col[n]]), n[6:]) for n in stall_cols), reverse=True)[:1]
    print(f"0x{int(r[0],16)-base:05x} {100*int(r[col['# Samples']])/tot:5.2f}% ex={int(r[col['Instructions Executed']])/1e6:6.1f}M "
          f"{r[col['Source']].strip()[:64]:64s} {st[0][1]} {100*st[0][0]/tot:.2f}")
