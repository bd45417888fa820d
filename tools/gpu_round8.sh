#!/bin/bash
# 2-GPU validation of the multi-rank synchronisation (run with gpurun --gpus 2).
TAG=${1:-r01j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 400 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x ) > $OUT/pytest_multirank.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_multirank.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err
tail -4 $OUT/pytest_multirank.log
python - <<PY
import json
d=json.loads(open("$OUT/bench_2gpu.json").read().strip().splitlines()[-1]); print("2gpu", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), d["state_finite"], d["roofline"]["time_share"])
PY
tail -n 3 $OUT/bench_2gpu.err
