#!/bin/bash
# One GPU-box pass of the round's final evidence.  usage (through gpurun): bash scripts/gpu_final.sh <tag>
TAG=${1:-final}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1; nproc >> $OUT/gpu.txt
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log; tail -3 $OUT/pytest.log
cp gpurun_out/parity.json $OUT/parity.json 2>/dev/null
timeout 900 python bench.py --steps 50 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json | cut -c1-600
# the other configs (device-resident quick lines)
for spec in "biped_walk_stairs 4096" "biped_walk_stairs 16384" "hyq_gallop_gap 4096" "hyq_gallop_gap 32768" "anymal_trot_mixed 4096" "anymal_trot_mixed 8192" "anymal_trot_block 32768"; do
  set -- $spec
  echo "== $1 B=$2" | tee -a $OUT/other_configs.txt
  timeout 600 python bench.py --quick --workload $1 --batch $2 --steps 30 --warmup 5 2>&1 | tail -1 | tee -a $OUT/other_configs.txt
done
# ncu launch lists and full captures: config 2 (shipped), config 4 (shipped and the library before this session: PhaseJac)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv env TWB_NO_CONFIG5=1 python bench.py --quick --steps 3 --warmup 3 > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'DynOut|RomNodeOut|TransposeIn' --launch-skip 9 --launch-count 3 -o $OUT/full python bench.py --quick --steps 3 --warmup 3 > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_config4.csv python bench.py --quick --workload hyq_gallop_gap --steps 3 --warmup 3 > $OUT/ncu_launches4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'DynOut|RomNodeOut|DynTailOut|PhaseJac' --launch-skip 12 --launch-count 4 -o $OUT/full_config4 python bench.py --quick --workload hyq_gallop_gap --steps 3 --warmup 3 > $OUT/ncu_full4.log 2>&1; echo "ncu full config4 rc=$?"
TWB_LIB=$PWD/towr_b200/variants/libtowr_b200_before.so timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_config4_before.csv python bench.py --quick --workload hyq_gallop_gap --steps 3 --warmup 3 > $OUT/ncu_launches4b.log 2>&1
TWB_LIB=$PWD/towr_b200/variants/libtowr_b200_before.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:'PhaseJac' --launch-skip 3 --launch-count 1 -o $OUT/full_config4_before python bench.py --quick --workload hyq_gallop_gap --steps 3 --warmup 3 > $OUT/ncu_full4b.log 2>&1; echo "ncu full config4 before rc=$?"
ls -la $OUT
