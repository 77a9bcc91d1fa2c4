#!/bin/bash
# One multi-GPU box call (gpurun --gpus 8): tile-sharded inference at 1/2/4/8 GPUs, the 5-level nets at 1/2/4/8,
# cfg-3 at 8, and the 2-GPU equivalence checks.   usage: gpurun --gpus 8 --timeout 900 -- 'bash tools/multi_gpu_r02.sh <tag>'
set -u
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
run() {  # workload, ngpus, extra flags
  local W=$1 N=$2; shift 2
  local F=$OUT/bench_${W}_n${N}_$TAG
  if [ "$N" = "1" ]; then
    python bench.py --workload $W --gpus 1 --steps 6 --warmup 3 "$@" > $F.json 2> $F.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
      bench.py --workload $W --gpus $N --steps 6 --warmup 3 "$@" > $F.json 2> $F.err
  fi
  echo "$W N=$N rc=$?"; python - <<PY
import json
try:
    d = json.loads([l for l in open("$F.json") if l.startswith("{")][-1])
    print("   ", d["metric"], "%.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"])
except Exception as e:
    print("   no line:", e)
PY
}
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_check.py > $OUT/dp_check_$TAG.log 2>&1; echo "dp_check rc=$?"; tail -1 $OUT/dp_check_$TAG.log
for N in 1 2 4 8; do run cfg4 $N --no-cpu-baseline; done
for N in 1 2 4 8; do run cfg5 $N --no-cpu-baseline; done
for N in 1 8; do run cfg5_res $N --no-cpu-baseline; done
for N in 1 8; do run cfg4_o32 $N --no-cpu-baseline; done
run cfg3 8 --no-cpu-baseline
