#!/bin/bash
# Round-2 GPU call driver: sections are run in the order given.
#   usage: gpurun --timeout 1500 -- 'bash tools/gpu_r02.sh <tag> [parity] [tests] [smoke] [bench] [ref] [workloads] [launches] [kmetrics] [full] [convbench]'
# Everything lands under gpurun_out/ with the tag in its name; the files worth keeping are copied to profiles/ by hand.
set -u
TAG=${1:-r02}; shift || true
WHAT=" ${*:-tests smoke bench} "
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu_$TAG.txt 2>&1
has() { [[ "$WHAT" == *" $1 "* ]]; }
if has parity; then
  MEDNET_PARITY_REPORT=$OUT/parity_$TAG python -m pytest tests/test_parity_gpu.py -m gpu -q -s > $OUT/parity_$TAG.log 2>&1; echo "parity rc=$?"
  grep -v "^    \|^  " $OUT/parity_$TAG.log | tail -40
fi
if has tests; then
  python -m pytest tests -m gpu -x -q > $OUT/t_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/t_gpu_$TAG.log
  tail -12 $OUT/t_gpu_$TAG.log
fi
if has smoke; then
  python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke_$TAG.log
fi
if has bench; then
  python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
fi
if has ref; then
  python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"; cat $OUT/bench_ref_$TAG.json
fi
if has workloads; then
  for W in ${WORKLOADS:-cfg1 cfg2 cfg2_res cfg3_ce cfg5 cfg5_res res32 cfg4 cfg4_o32}; do
    python bench.py --workload $W --steps 5 --warmup 3 ${WL_FLAGS:---no-cpu-baseline} > $OUT/bench_${W}_$TAG.json 2> $OUT/bench_${W}_$TAG.err
    echo "$W rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_${W}_$TAG.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print("   ", d["metric"], "%.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "conv frac", r.get("frac"), "share", r.get("share_of_step"))
except Exception as e:
    print("   no line:", e)
PY
    tail -2 $OUT/bench_${W}_$TAG.err
  done
fi
if has convbench; then
  python tools/conv_bench.py ${CONVBENCH_ARGS:-} > $OUT/convbench_$TAG.txt 2>&1; echo "convbench rc=$?"; cat $OUT/convbench_$TAG.txt | tail -40
fi
CMD="python bench.py --workload ${NCU_WORKLOAD:-cfg3} --steps 1 --warmup 3 --no-cpu-baseline"
if has launches || has kmetrics || has full; then
  $CMD > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.log; exit 1; }
fi
if has launches; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_a_$TAG.log 2>&1
  echo "ncu launches rc=$?"
  python tools/launch_summary.py $OUT/launches_$TAG.csv > $OUT/launches_$TAG.txt 2>&1; head -40 $OUT/launches_$TAG.txt
fi
if has kmetrics; then
  M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,launch__registers_per_thread,launch__grid_size
  ncu --clock-control none --metrics $M -k regex:"${KREGEX:-^(conv|wgrad|gn_|upcat|maxpool|dice|adam|act_|pack_|cast_|transpose_|ce_|hm_|layout|border|affine|split)}" -c ${KCOUNT:-300} --csv --log-file $OUT/kernels_$TAG.csv $CMD > $OUT/ncu_d_$TAG.log 2>&1
  echo "ncu per-kernel metrics rc=$?"
  python tools/kernel_metrics.py $OUT/kernels_$TAG.csv > $OUT/kernels_$TAG.txt 2>&1; head -60 $OUT/kernels_$TAG.txt
fi
if has full; then
  # the dominant kernel with the full set + source: launches 14..16 of a step = decoder-join coarse part (UPCONV_F), its
  # full-resolution skip part (kd-merged, fp32 addend) and the kd-merged 64 -> 64 @ 128^3 layer
  ncu --set full --clock-control none --import-source on -k regex:conv3_tc_kernel -s 13 -c 3 -f -o $OUT/conv3_tc_$TAG $CMD > $OUT/ncu_b_$TAG.log 2>&1
  echo "ncu conv full rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 0 -c 2 -f -o $OUT/wgrad_tc_$TAG $CMD > $OUT/ncu_c_$TAG.log 2>&1
  echo "ncu wgrad full rc=$?"
  for K in conv3_tc wgrad_tc; do
    ncu -i $OUT/${K}_$TAG.ncu-rep --page raw --csv > $OUT/${K}_${TAG}_raw.csv 2>/dev/null
  done
  ls -la $OUT/*.ncu-rep
fi
du -sh $OUT
