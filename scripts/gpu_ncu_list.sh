#!/bin/bash
# usage (through gpurun): bash scripts/gpu_ncu_list.sh <tag> <workload> [lib] — ncu launch list (gpu__time_duration) of a quick bench run
TAG=${1:-ncu}; WL=${2:-anymal_trot_block}; LIB=${3:-towr_b200/libtowr_b200.so}; OUT=gpurun_out/$TAG; mkdir -p $OUT
TWB_LIB=$PWD/$LIB timeout 300 python bench.py --quick --workload $WL --steps 3 --warmup 3 > $OUT/plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain.log; exit 1; }
TWB_LIB=$PWD/$LIB timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$WL.csv python bench.py --quick --workload $WL --steps 3 --warmup 3 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_list.py $OUT/launches_$WL.csv; exit 0
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open("$OUT/launches_$WL.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0].split("<")[0].split("::")[-1]
    v = float(r[vi].replace(",", "")); v = v / 1e3 if r[ui] in ("ns", "nsecond") else v
    agg.setdefault(name, []).append(v)
for k, v in agg.items(): print(f"{k:20s} n={len(v):3d} avg {sum(v)/len(v):9.2f} us  min {min(v):9.2f}")
PY
