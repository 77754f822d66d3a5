#!/bin/bash
# usage (through gpurun): bash scripts/gpu_env_ab.sh <tag> <workload> "<ENV=.. ENV=..>" "<ENV...>" ... — bench --quick per environment setting on ONE box
TAG=$1; WL=$2; shift; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
for envs in "$@"; do
  echo "== [$envs] $WL" | tee -a $OUT/exp.txt
  for rep in 1 2 3; do
    env $envs timeout 300 python bench.py --quick --workload $WL --steps 50 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('step_us', round(d['ms_per_step']*1e3,2), 'best', round(d['ms_best']*1e3,2), 'Mevals', round(d['value']/1e6,2), 'frac', round(d['frac'],4))" | tee -a $OUT/exp.txt
  done
done
