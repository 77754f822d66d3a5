#!/bin/bash
# usage (through gpurun): bash scripts/gpu_sweep.sh <tag>   — parity tests on the default build, then kernel timing of every variant
TAG=${1:-sweep}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
echo "default: $(timeout 300 python bench.py --quick --steps 50 --warmup 5 2>&1 | tail -1)" | tee -a $OUT/sweep.txt
for so in towr_b200/variants/*.so; do
  for env in "" $SWEEP_ENVS; do
    echo "$(basename $so) $env: $(env $env TWB_LIB=$PWD/$so timeout 300 python bench.py --quick --steps 50 --warmup 5 2>&1 | tail -1)" | tee -a $OUT/sweep.txt
  done
done
TWB_PROFILE=1 timeout 300 python bench.py --quick --steps 20 --warmup 3 > /dev/null 2> $OUT/profile_default.txt; tail -6 $OUT/profile_default.txt
