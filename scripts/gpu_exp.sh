#!/bin/bash
# usage (through gpurun): bash scripts/gpu_exp.sh <tag> [variant tags...] — per-kernel event profile + pipeline timing of variant builds
TAG=${1:-exp}; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ $# -gt 0 ]; then LIST=""; for t in "$@"; do LIST="$LIST towr_b200/variants/libtowr_b200_$t.so"; done; else LIST=$(ls towr_b200/variants/*.so); fi
for so in $LIST; do
  echo "== $(basename $so) $TWB_EXTRA" | tee -a $OUT/exp.txt
  TWB_LIB=$PWD/$so timeout 300 python bench.py --quick --steps 50 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('step_us', round(d['ms_per_step']*1e3,2), 'best', round(d['ms_best']*1e3,2), 'Mevals', round(d['value']/1e6,2), 'frac', round(d['frac'],4))" | tee -a $OUT/exp.txt
  TWB_PROFILE=1 TWB_LIB=$PWD/$so timeout 300 python bench.py --quick --steps 20 --warmup 3 2>&1 >/dev/null | grep "twb profile" | tee -a $OUT/exp.txt
done
