#!/bin/bash
# Final GPU visit of a round (1 GPU): full GPU test suite, bench line with CPU baseline, reference arm, other
# configurations, cuFFT comparison, launch list, ncu --set full of one stage.
TAG=${1:-r01i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -q --durations=8 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_512_1gpu.json 2> $OUT/bench_512_1gpu.err
LAPS_TUNE_RHS=0 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_rhs0.json 2> $OUT/bench_rhs0.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 300 python tools/bench_configs.py > $OUT/configs.jsonl 2> $OUT/configs.err
timeout 300 python tools/cufft_compare.py 512 > $OUT/cufft_compare.json 2> $OUT/cufft_compare.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_512_1gpu.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_rhs_z|k_spec_z|k_fwd|k_inv|k_flux|k_cfl" -s 30 -c 10 \
  -o $OUT/stage_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ncu -i $OUT/stage_full.ncu-rep --page raw --csv > $OUT/stage_full_raw.csv 2>/dev/null
ls -la $OUT
tail -12 $OUT/pytest_gpu.log
cat $OUT/bench_512_1gpu.json
python - <<PY
import json
for f in ("bench_512_1gpu","bench_rhs0"):
    try:
        d=json.loads(open("$OUT/%s.json"%f).read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), d["roofline"]["time_share"])
    except Exception as e: print(f, "failed", e)
PY
cat $OUT/configs.jsonl $OUT/cufft_compare.json
for f in $OUT/*.err; do tail -n 3 $f; done
