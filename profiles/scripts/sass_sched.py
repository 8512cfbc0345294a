#!/usr/bin/env python3
"""Decode `cuobjdump -sass` text of one function: opcode histogram and the scheduler's stall counts per address range.
usage: sass_sched.py file.sass [lo_hex hi_hex] [--list]"""
import re, sys, collections

def parse(path):
    ins = []
    cur = None
    for line in open(path):
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/', line)
        if m:
            cur = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), None]
            ins.append(cur)
            continue
        m = re.match(r'\s*/\* (0x[0-9a-f]{16}) \*/', line)
        if m and cur is not None and cur[3] is None:
            cur[3] = int(m.group(1), 16)
    return ins

def main():
    path = sys.argv[1]
    args = [a for a in sys.argv[2:] if not a.startswith('--')]
    lo, hi = (int(args[0], 16), int(args[1], 16)) if len(args) >= 2 else (0, 1 << 60)
    ins = [i for i in parse(path) if lo <= i[0] < hi]
    hist = collections.Counter()
    stall = collections.Counter()
    for addr, text, w0, w1 in ins:
        t = re.sub(r'^@!?U?P\d+\s+', '', text)
        op = t.split()[0]
        st = (w1 >> 41) & 0xf if w1 is not None else 0
        hist[op] += 1
        stall[op] += st
        if '--list' in sys.argv:
            print(f"{addr:06x} st={st:2d} y={(w1>>45)&1} wr={(w1>>46)&7} rd={(w1>>49)&7} wait={(w1>>52)&0x3f:02x}  {text}")
    n = sum(hist.values())
    print(f"{n} instructions, total stall count {sum(stall.values())}")
    for op, c in hist.most_common(40):
        print(f"  {op:28s} {c:6d} {100.0*c/n:5.1f}%  avg stall {stall[op]/c:4.1f}")

main()
