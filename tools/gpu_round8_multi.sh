#!/bin/bash
# First multi-GPU visit after the switch of the default Fourier-row ownership (run with gpurun --gpus N, N = 4 or 8):
# multi-rank parity tests in both forms, then the bench at N GPUs with the round-robin rows (default) and with the
# reference's decompose_1d slabs (LAPS_TUNE_CYCLIC=0) — time per step, the per-kernel shares and the NVLink egress.
TAG=${1:-r02a}
N=${2:-4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x ) > $OUT/pytest_multirank.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_multirank.log
run() {  # name, env assignment
  env $2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_${N}gpu_$1.json 2> $OUT/bench_${N}gpu_$1.err
}
run cyclic LAPS_TUNE_CYCLIC=1
run slabs LAPS_TUNE_CYCLIC=0
run default LAPS_NOOP=1
tail -5 $OUT/pytest_multirank.log
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_*gpu_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        nv=d.get("nvlink") or {}
        print(f.split('/')[-1], round(d["ms_per_step"],2), "ms/step; decomposition:", d["config"]["decomposition"])
        print("   shares", d["roofline"]["time_share"])
        print("   nvlink", {k: round(v["egress_GBps"]) for k, v in nv.get("per_kernel", {}).items()}, "over the step", round(nv.get("egress_GBps_over_the_step", 0)))
    except Exception as e: print(f, "failed", e)
PY
for f in $OUT/*.err; do tail -n 4 $f; done
