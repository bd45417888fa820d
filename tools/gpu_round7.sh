#!/bin/bash
# Multi-GPU visit: multi-rank parity tests over NVLink and the bench at N GPUs (run with gpurun --gpus N).
TAG=${1:-r01h}
N=${2:-4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x ) > $OUT/pytest_multirank.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_multirank.log
for n in 2 $N; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > $OUT/bench_${n}gpu.json 2> $OUT/bench_${n}gpu.err
done

ls -la $OUT
tail -5 $OUT/pytest_multirank.log
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_*gpu.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), d["roofline"]["time_share"])
    except Exception as e: print(f, "failed", e)
PY
for f in $OUT/*.err; do tail -n 4 $f; done
