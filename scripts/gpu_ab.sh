#!/bin/bash
# usage (through gpurun): bash scripts/gpu_ab.sh <tag> <workload> <batch|0> lib1 lib2 ... — bench --quick + ncu launch list per library on ONE box (A/B)
TAG=$1; WL=$2; B=$3; shift; shift; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
BATCH=""; [ "$B" != "0" ] && BATCH="--batch $B"
for so in "$@"; do
  echo "== $(basename $so) $WL $BATCH" | tee -a $OUT/exp.txt
  for rep in 1 2; do
  TWB_LIB=$PWD/$so timeout 300 python bench.py --quick --workload $WL $BATCH --steps 30 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('step_us', round(d['ms_per_step']*1e3,2), 'best', round(d['ms_best']*1e3,2), 'Mevals', round(d['value']/1e6,2), 'frac', round(d['frac'],4))" | tee -a $OUT/exp.txt
  done
  TWB_LIB=$PWD/$so timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $OUT/launches_$(basename $so .so).csv python bench.py --quick --workload $WL $BATCH --steps 3 --warmup 3 > $OUT/ncu.log 2>&1
  python scripts/ncu_list.py $OUT/launches_$(basename $so .so).csv | tee -a $OUT/exp.txt
done
