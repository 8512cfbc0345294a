"""ncu --metrics gpu__time_duration.sum --csv launch list -> per-kernel table (markdown) and, with --seq NAME, the
launch-by-launch durations between two launches of kernel NAME (one iteration of the step)."""
import csv, sys, collections, re

path = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = []
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    us = v / 1000.0 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1000.0 if u in ("ms", "msecond") else v
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"<.*", "", name)
    launches.append((name, us))
if "--seq" in sys.argv:
    key = sys.argv[sys.argv.index("--seq") + 1]
    idx = [i for i, (n, _) in enumerate(launches) if key in n]
    k = int(sys.argv[sys.argv.index("--seq") + 2]) if len(sys.argv) > sys.argv.index("--seq") + 2 else len(idx) - 2
    a, b = idx[k], idx[k + 1]
    for n, us in launches[a:b]:
        print("%-60s %10.1f us" % (n[-60:], us))
    print("total %.1f us in %d launches" % (sum(us for _, us in launches[a:b]), b - a))
else:
    tot = collections.defaultdict(lambda: [0, 0.0])
    for n, us in launches:
        tot[n][0] += 1
        tot[n][1] += us
    total = sum(v[1] for v in tot.values())
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for n, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.3f | %.1f%% |" % (n, c, us / 1000.0, 100.0 * us / total))
