#!/bin/bash
# Second GPU visit: fused flux + forward x pass against the separate kernels, GPU tests, other configs.
TAG=${1:-r01c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
LAPS_TUNE_FUSEX=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_fused.json 2> $OUT/bench_fused.err
LAPS_TUNE_FUSEX=0 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_unfused.json 2> $OUT/bench_unfused.err
( time timeout 900 python -m pytest tests -m gpu -q --durations=10 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
timeout 300 python tools/bench_configs.py > $OUT/configs.jsonl 2> $OUT/configs.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_flux_fwd_x|k_flux|k_inv_x" -s 6 -c 3 \
  -o $OUT/fused_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ncu -i $OUT/fused_full.ncu-rep --page raw --csv > $OUT/fused_full_raw.csv 2>/dev/null
ls -la $OUT
tail -5 $OUT/pytest_gpu.log
python - <<PY
import json
for f in ("bench_fused","bench_unfused"):
    try:
        d=json.load(open("$OUT/%s.json"%f)); print(f, d["ms_per_step"], d["roofline"]["time_share"], d["roofline"]["per_kernel_GBps"])
    except Exception as e: print(f, "failed", e)
PY
cat $OUT/configs.jsonl; tail -3 $OUT/configs.err
