#!/bin/bash
# Fourth GPU visit (1 GPU): symmetric tensor (default), z-chunked flux/forward-x interleave with L2 persistence.
TAG=${1:-r01e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_$name.json 2> $OUT/bench_$name.err; }
run default LAPS_X=0
run nosym LAPS_TUNE_SYM=0
run zchunk1 LAPS_TUNE_ZCHUNK=1
run zchunk2 LAPS_TUNE_ZCHUNK=2
run zchunk4 LAPS_TUNE_ZCHUNK=4
run zchunk2_nopersist LAPS_TUNE_ZCHUNK=2 LAPS_TUNE_L2PERSIST=0
run rcg2 LAPS_TUNE_RCG=2
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x ) > $OUT/pytest_parity.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_parity.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_flux|k_fwd_x" -s 40 -c 24 --csv --log-file $OUT/zchunk2_launches.csv \
  env LAPS_TUNE_ZCHUNK=2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/zchunk2_ncu.log 2>&1
ls -la $OUT
tail -4 $OUT/pytest_parity.log
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d["ms_per_step"],2), d["roofline"]["time_share"])
    except Exception as e: print(f, "failed", e)
PY
for f in $OUT/*.err; do tail -n 3 $f; done
