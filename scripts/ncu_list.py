"""Per-kernel averages of an ncu launch list (--metrics gpu__time_duration.sum --csv): python scripts/ncu_list.py <csv>..."""
import collections, csv, re, sys
for path in sys.argv[1:]:
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        m = re.search(r"(\w+)(<[^(]*)?\(", r[ki]); name = m.group(1) if m else r[ki][:30]
        v = float(r[vi].replace(",", "")); v = v / 1e3 if r[ui].startswith("n") else v
        agg.setdefault(name, []).append(v)
    print(path)
    for k, v in agg.items():
        if "elementwise" in k or "reduce" in k: continue
        print(f"  {k:24s} n={len(v):3d} avg {sum(v)/len(v):9.2f} us  min {min(v):9.2f}")
