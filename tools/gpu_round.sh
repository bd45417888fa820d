#!/bin/bash
# One GPU-box visit: GPU parity tests, the bench line, the ncu launch list of the same command and
# one `ncu --set full` capture of one RK stage.  Usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG'
TAG=${1:-r01x}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=15 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_512_1gpu.json 2> $OUT/bench_512_1gpu.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_512_1gpu.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_rhs_z|k_spec_z|k_fwd|k_inv|k_flux|k_cfl" -s 28 -c 11 \
  -o $OUT/stage_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ncu -i $OUT/stage_full.ncu-rep --page raw --csv > $OUT/stage_full_raw.csv 2>/dev/null
ls -la $OUT
tail -5 $OUT/pytest_gpu.log
cat $OUT/bench_512_1gpu.json
