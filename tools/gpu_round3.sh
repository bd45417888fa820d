#!/bin/bash
# Third GPU visit (2 GPUs): mass-flux-from-state, group barriers; fused variant again; 2-GPU parity + bench.
TAG=${1:-r01d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_default.json 2> $OUT/bench_default.err
LAPS_TUNE_FUSEX=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_fused.json 2> $OUT/bench_fused.err
LAPS_TUNE_MASS=0 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_nomass.json 2> $OUT/bench_nomass.err
( time timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x ) > $OUT/pytest_multirank.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_multirank.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 6 --warmup 3 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x ) > $OUT/pytest_parity.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_parity.log
ls -la $OUT
tail -4 $OUT/pytest_multirank.log; tail -4 $OUT/pytest_parity.log
python - <<PY
import json
for f in ("bench_default","bench_fused","bench_nomass","bench_2gpu"):
    try:
        d=json.loads(open("$OUT/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["roofline"]["time_share"], {k:int(v) for k,v in d["roofline"]["per_kernel_GBps"].items()})
    except Exception as e: print(f, "failed", e)
PY
tail -3 $OUT/*.err
