"""Top source lines (instructions executed, stall samples) of one kernel of an ncu report.
usage: python scripts/ncu_hot.py <report.ncu-rep> <kernel regex> [top_n]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; data = []; seen_kernels = 0
for r in rows:
    if len(r) > 5 and r[0] == 'Line No':
        seen_kernels += 1
        if seen_kernels > 1: break
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit(): data.append(r)
ie = hdr.index('Instructions Executed'); sm = hdr.index('# Samples')
tot = sum(int(d[ie]) for d in data); tots = sum(int(d[sm]) for d in data)
print("warp instructions", tot, "samples", tots)
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
agg = {}
for d in data:
    for i in stall_cols:
        if d[i].isdigit(): agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + int(d[i])
print("stalls:", ", ".join(f"{k} {v / max(tots,1) * 100:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
for d in sorted(data, key=lambda d: -int(d[ie]))[:top]:
    st = {hdr[i][6:]: int(d[i]) for i in stall_cols if d[i].isdigit() and int(d[i]) > 0}
    t3 = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{int(d[ie]) / tot * 100:5.1f}% inst {int(d[sm]) / max(tots,1) * 100:5.1f}% smp L{d[0]}: {d[1].strip()[:80]}  {t3}")
