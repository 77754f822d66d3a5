"""Summarise an ncu report per CUDA source line (instructions executed, stall samples).
usage: python scripts/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file = None; hdr = None; data = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples")
        st = {k: int(r[i]) for i, k in enumerate(hdr) if k.startswith("stall_") and "Not Issued" not in k and r[i].isdigit()}
        data.append((int(r[ie]), int(r[sm]), cur_file, int(r[0]), r[1].strip(), st))
tot = sum(d[0] for d in data); tots = sum(d[1] for d in data)
print("total warp-instructions", tot, "stall samples", tots)
agg = {}
for d in data:
    for k, v in d[5].items(): agg[k] = agg.get(k, 0) + v
print("stalls:", ", ".join(f"{k[6:]} {v/tots*100:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for d in sorted(data, key=lambda d: -d[1])[:top]:
    main = max(d[5].items(), key=lambda kv: kv[1])[0][6:] if d[5] else ""
    print(f"{d[0]/tot*100:5.1f}% inst {d[1]/tots*100:5.1f}% smp [{main:10s}] {d[2]}:{d[3]}: {d[4][:100]}")
