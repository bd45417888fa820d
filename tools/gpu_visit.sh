#!/bin/bash
# One GPU visit = a list of named stages (replaces the per-round scripts of round 1).  Results go to gpurun_out/<tag>/.
#   tools/gpu_visit.sh <tag> <stage> [<stage> ...]           (under gpurun; add --gpus N to gpurun for the multi-GPU stages)
# stages: tests | tests_multi | odd | bench | bench_ref | ab | configs | launches | ncu_stage | multi<N> | cfg5
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
trun() { n=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29$((500 + RANDOM % 400)) "$@"; }
for stage in "$@"; do
  echo "== stage $stage ($(date +%T))"
  case $stage in
    tests)        ( time timeout 1500 python -m pytest tests -m gpu -q --durations=10 --maxfail=8 ) > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -15 $OUT/pytest_gpu.log ;;
    tests_new)    ( time timeout 900 python -m pytest tests/test_gpu_driver.py tests/test_gpu_parity.py -m gpu -q -k "driver or async or rhs_kernel or cfl_screen or two_stream or 2d_tree_parity or 2d_library" ) > $OUT/pytest_new.log 2>&1; tail -5 $OUT/pytest_new.log ;;
    tests_multi)  ( time timeout 1200 python -m pytest tests/test_gpu_multirank.py -m gpu -q --durations=10 ) > $OUT/pytest_multirank_${NG}gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_multirank_${NG}gpu.log; tail -8 $OUT/pytest_multirank_${NG}gpu.log ;;
    odd)          ( time timeout ${ODD_T:-100} python -m pytest tests/test_gpu_z_odd_factor_lines.py -m gpu -q --durations=8 --maxfail=20 ) > $OUT/pytest_odd.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_odd.log; tail -12 $OUT/pytest_odd.log
                  timeout ${ODD_B:-50} python bench.py --n 384 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_384_1gpu.json 2> $OUT/bench_384_1gpu.err; cut -c 1-400 $OUT/bench_384_1gpu.json; tail -c 300 $OUT/bench_384_1gpu.err ;;
    bench)        timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_512_1gpu.json 2> $OUT/bench_512_1gpu.err; tail -c 600 $OUT/bench_512_1gpu.err ;;
    bench_ref)    timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; cat $OUT/bench_reference.json ;;
    ab)           timeout 900 python tools/ab_tune.py > $OUT/ab_tune.jsonl 2> $OUT/ab_tune.err; cut -c 1-260 $OUT/ab_tune.jsonl; tail -3 $OUT/ab_tune.err ;;
    configs)      for c in 1 2 3; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_config$c.json 2> $OUT/bench_config$c.err; done ;;
    launches)     timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_512_1gpu.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > $OUT/launches_bench.log 2>&1 ;;
    ncu_stage)    timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_rhs_z|k_spec_z|k_fwd|k_inv|k_flux|k_cfl" -s 60 -c 12 -o $OUT/stage_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > $OUT/ncu_full.log 2>&1
                  ncu -i $OUT/stage_full.ncu-rep --page raw --csv > $OUT/stage_full_raw.csv 2>/dev/null
                  python tools/ncu_traffic.py $OUT/stage_full_raw.csv 4 512 1 "profiles/${TAG}_ncu_stage.md (ncu --set full, 512^3 Hall + expanding box, 1 B200; dram__bytes_read.sum + dram__bytes_write.sum per launch)" $OUT/traffic.json $OUT/ncu_stage_table.md > /dev/null
                  rm -f $OUT/stage_full.ncu-rep; cat $OUT/ncu_stage_table.md ;;
    ncu_z)        timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_rhs_z|k_inv_y" -s 8 -c 2 -o $OUT/zpass python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > $OUT/ncu_z.log 2>&1
                  ncu -i $OUT/zpass.ncu-rep --page raw --csv > $OUT/zpass_raw.csv 2>/dev/null; ls -la $OUT/zpass.ncu-rep ;;
    multi*)       n=${stage#multi}; trun $n bench.py --gpus $n --steps 10 --warmup 3 > $OUT/bench_512_${n}gpu.json 2> $OUT/bench_512_${n}gpu.err; tail -c 400 $OUT/bench_512_${n}gpu.err
                  LAPS_TUNE_OVERLAP=${ALT_OVERLAP:-0} trun $n bench.py --gpus $n --steps 10 --warmup 3 --no-parity > $OUT/bench_512_${n}gpu_overlap${ALT_OVERLAP:-0}.json 2> $OUT/bench_512_${n}gpu_overlap${ALT_OVERLAP:-0}.err ;;
    abmulti*)     n=${stage#abmulti}; LAPS_TUNE_STAGING=1 trun $n tools/ab_tune.py --rounds 3 --steps 5 --variants serial=overlap:0 form1=overlap:1 form2=overlap:2 \
                    form1_c2=overlap:1,ovl_chunks:2 form1_c4=overlap:1,ovl_chunks:4 form1_y8=overlap:1,ovl_y:8 form1_y32=overlap:1,ovl_y:32 \
                    form1_z16=overlap:1,ovl_z:16 form1_z0=overlap:1,ovl_z:0 form2_c4=overlap:2,ovl_chunks:4 > $OUT/ab_tune_${n}gpu.jsonl 2> $OUT/ab_tune_${n}gpu.err
                  grep "^{" $OUT/ab_tune_${n}gpu.jsonl | cut -c 1-200; tail -3 $OUT/ab_tune_${n}gpu.err ;;
    tests_nccl)   ( time timeout 1200 python -m pytest tests/test_gpu_multirank.py -m gpu -q --durations=10 -k "not connect_local and not exchange_wait" ) > $OUT/pytest_multirank_nccl_${NG}gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_multirank_nccl_${NG}gpu.log; tail -8 $OUT/pytest_multirank_nccl_${NG}gpu.log ;;
    cfg5)         trun $NG bench.py --gpus $NG --config 5 --steps 5 --warmup 3 > $OUT/bench_config5_${NG}gpu.json 2> $OUT/bench_config5_${NG}gpu.err; tail -c 400 $OUT/bench_config5_${NG}gpu.err ;;
    *)            echo "unknown stage $stage" ;;
  esac
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f.split('/')[-1], "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"].get("ms_per_step", 0), 3), "parity", (d.get("parity") or {}).get("ok"),
              "k0", (d.get("k0_mode") or {}).get("ok"), "step_frac", round(r.get("step_frac", 0), 3), "top", r.get("kernel"), round(r.get("frac", 0), 3))
        print("   frac", r.get("per_kernel_frac"))
        print("   share", r.get("time_share"))
        nv = d.get("nvlink") or {}
        if nv: print("   nvlink", {k: round(v["egress_GBps"]) for k, v in nv.get("per_kernel", {}).items()}, "over the step", round(nv.get("egress_GBps_over_the_step", 0)))
    except Exception as e:
        print(f, "failed", e)
PY
ls -la $OUT | head -40
