#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), then ncu launch list + captures of the hot kernels.
# Keeps gpurun_out/ small (the 64 MiB pull limit): .ncu-rep files are exported to CSV on the box and deleted, except
# one 2-launch report of the conv kernel with source.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag> [tests|bench|ncu ...]'
set -u
TAG=${1:-r01}; shift || true
WHAT="${*:-tests bench ncu}"
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu_$TAG.txt 2>&1
if [[ " $WHAT " == *" tests "* ]]; then
  python -m pytest tests -m gpu -x -q > $OUT/t_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/t_gpu_$TAG.log
  tail -15 $OUT/t_gpu_$TAG.log
  python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
fi
if [[ " $WHAT " == *" bench "* ]]; then
  python bench.py --steps 8 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
  python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"; cat $OUT/bench_ref_$TAG.json
fi
[[ " $WHAT " == *" ncu "* ]] && WHAT="$WHAT launches kmetrics full"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
if [[ " $WHAT " == *" launches "* || " $WHAT " == *" kmetrics "* || " $WHAT " == *" full "* ]]; then
  $CMD > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.log; exit 1; }
fi
if [[ " $WHAT " == *" launches "* ]]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_a_$TAG.log 2>&1
  echo "ncu launches rc=$?"
fi
if [[ " $WHAT " == *" kmetrics "* ]]; then
  M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,launch__registers_per_thread,launch__grid_size
  # one training step's worth of every kernel of ours (-k matches the unqualified function name), metrics only
  ncu --clock-control none --metrics $M -k regex:'^(conv|wgrad|gn_|upcat|maxpool|dice|adam|act_|pack_|cast_|transpose_|ce_|hm_)' -c 300 --csv --log-file $OUT/kernels_$TAG.csv $CMD > $OUT/ncu_d_$TAG.log 2>&1
  echo "ncu per-kernel metrics rc=$?"
fi
if [[ " $WHAT " == *" full "* ]]; then
  # the dominant kernel with the full set + source (2 launches of the widest layer: skip the first 20 conv launches)
  ncu --set full --clock-control none --import-source on -k regex:conv3_tc_kernel -s 20 -c 2 -f -o $OUT/conv3_tc_$TAG $CMD > $OUT/ncu_b_$TAG.log 2>&1
  echo "ncu conv full rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 10 -c 1 -f -o $OUT/wgrad_tc_$TAG $CMD > $OUT/ncu_c_$TAG.log 2>&1
  echo "ncu wgrad full rc=$?"
fi
du -sh $OUT; ls -la $OUT | grep $TAG
