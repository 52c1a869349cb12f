"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) as a markdown share table:
python tools/ncu_summary.py gpurun_out/X_launches.csv "title" > profiles/X_launches_summary.md"""
import csv, re, sys
from collections import defaultdict

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "ncu launch list")
rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
agg = defaultdict(list)
ours = 0.0
for r in rows:
    name = r[4]
    mine = "<unnamed>::" in name
    short = re.sub(r"\(.*", "", name.replace("void ", "").replace("<unnamed>::", "").replace("at::", ""))
    us = int(r[14]) / 1e3
    agg[short[:100]].append(us)
    if mine:
        ours += us
tot = sum(sum(v) for v in agg.values())
print("# %s\n" % title)
print("%d launches, total kernel time %.1f ms; kernels of libd3feat_b200.so: %.1f ms (%.0f%%).  Per-launch times are "
      "cold-cache and serialised under ncu: read SHARES, not absolutes.\n" % (len(rows), tot / 1e3, ours / 1e3, 100 * ours / tot))
print("| share | total us | launches | avg us | max us | kernel |\n|---|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:45]:
    print("| %.1f%% | %.0f | %d | %.1f | %.1f | `%s` |" % (100 * sum(v) / tot, sum(v), len(v), sum(v) / len(v), max(v), k))
