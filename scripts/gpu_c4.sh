#!/bin/bash
# usage (through gpurun): bash scripts/gpu_c4.sh <tag> [workload] [libs...] — parity tests of the duration-optimised configs, then bench + ncu launch list per library
TAG=${1:-c4}; WL=${2:-hyq_gallop_gap}; shift; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
LIBS=${@:-towr_b200/libtowr_b200.so}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "config4 or durations or hopper or config2 or flags or config3" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log
tail -5 $OUT/pytest.log
for so in $LIBS; do
  echo "== $(basename $so) $WL" | tee -a $OUT/exp.txt
  TWB_LIB=$PWD/$so timeout 300 python bench.py --quick --workload $WL --steps 30 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('step_us', round(d['ms_per_step']*1e3,2), 'best', round(d['ms_best']*1e3,2), 'Mevals', round(d['value']/1e6,2), 'frac', round(d['frac'],4))" | tee -a $OUT/exp.txt
  bash scripts/gpu_ncu_list.sh $TAG/ncu_$(basename $so .so) $WL $so | tee -a $OUT/exp.txt
done
