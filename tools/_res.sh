OUT=gpurun_out
CMD="python bench.py --workload res32 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/plain_res32.log 2>&1 || { tail -5 $OUT/plain_res32.log; exit 1; }
tail -1 $OUT/plain_res32.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_res32.csv $CMD > $OUT/ncu_res32.log 2>&1
