#!/bin/bash
# usage (through gpurun): bash scripts/gpu_exp.sh <tag> — per-kernel event profile + pipeline timing of every variant build
TAG=${1:-exp}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for so in towr_b200/variants/*.so; do
  echo "== $(basename $so)" | tee -a $OUT/exp.txt
  TWB_LIB=$PWD/$so timeout 300 python bench.py --quick --steps 50 --warmup 5 2>&1 | tail -1 | tee -a $OUT/exp.txt
  TWB_PROFILE=1 TWB_LIB=$PWD/$so timeout 300 python bench.py --quick --steps 20 --warmup 3 2>&1 >/dev/null | grep "twb profile" | tee -a $OUT/exp.txt
done
