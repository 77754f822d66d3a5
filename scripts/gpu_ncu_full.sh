#!/bin/bash
# usage (through gpurun): bash scripts/gpu_ncu_full.sh <tag> <workload> <kernel regex> [launch-skip] [launch-count] — one ncu --set full capture of a quick bench run
TAG=${1:-full}; WL=${2:-anymal_trot_block}; RE=${3:-RomNodeOut}; SKIP=${4:-6}; CNT=${5:-3}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python bench.py --quick --workload $WL --steps 3 --warmup 3 > $OUT/plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RE" --launch-skip $SKIP --launch-count $CNT -o $OUT/full_$WL python bench.py --quick --workload $WL --steps 3 --warmup 3 > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
