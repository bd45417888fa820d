#!/bin/bash
# 2-GPU check of the experimental cyclic ky ownership (run with gpurun --gpus 2).
TAG=${1:-r01k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export LAPS_TUNE_CYCLIC=1
( time timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x ) > $OUT/pytest_multirank_cyclic.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_multirank_cyclic.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 > $OUT/bench_2gpu_cyclic.json 2> $OUT/bench_2gpu_cyclic.err
tail -4 $OUT/pytest_multirank_cyclic.log
python - <<PY
import json
d=json.loads(open("$OUT/bench_2gpu_cyclic.json").read().strip().splitlines()[-1]); print("2gpu cyclic", round(d["ms_per_step"],2), d["state_finite"], d["roofline"]["time_share"])
PY
tail -n 3 $OUT/bench_2gpu_cyclic.err
