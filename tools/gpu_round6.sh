#!/bin/bash
# Sixth GPU visit (1 GPU): circular column pruning + kz pruning of the state traffic.
TAG=${1:-r01g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_$name.json 2> $OUT/bench_$name.err; }
run default LAPS_X=0
run nocircle_nokz LAPS_TUNE_CIRCLE=0 LAPS_TUNE_KZPRUNE=0
run nokz LAPS_TUNE_KZPRUNE=0
( time timeout 900 python -m pytest tests -m gpu -q -x ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_rhs_z|k_spec_z|k_fwd_y|k_inv_y" -s 12 -c 5 \
  -o $OUT/z_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ncu -i $OUT/z_full.ncu-rep --page raw --csv > $OUT/z_full_raw.csv 2>/dev/null
ls -la $OUT
tail -6 $OUT/pytest_gpu.log
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d["ms_per_step"],2), d["roofline"]["time_share"], {k:int(v) for k,v in d["roofline"]["per_kernel_GBps"].items()})
    except Exception as e: print(f, "failed", e)
PY
for f in $OUT/*.err; do tail -n 3 $f; done
