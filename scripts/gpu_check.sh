#!/bin/bash
# One GPU-box pass: parity tests, bench line, per-kernel event profile, ncu launch list, ncu full capture.
# usage (through gpurun): bash scripts/gpu_check.sh <tag> [skip_tests]
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt
if [ -z "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log
  tail -5 $OUT/pytest.log
fi
timeout 600 python bench.py --steps 50 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json
TWB_PROFILE=1 timeout 300 python bench.py --quick --steps 20 --warmup 3 > $OUT/profile_quick.json 2> $OUT/profile_kernels.txt; echo "profile rc=$?"
cat $OUT/profile_kernels.txt | tail -12
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv python bench.py --quick --steps 3 --warmup 3 > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'DynOut|RomOut|NodeOut|SplineKernel|TransposeIn|ConstOut|Eval' --launch-skip 12 --launch-count 6 -o $OUT/full python bench.py --quick --steps 3 --warmup 3 > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
