#!/bin/bash
# Developer helper: multi-GPU measurements of one gpurun call.  usage: tools/scale_sweep.sh N  (N = GPUs of the box)
N=$1; OUT=gpurun_out/r02_scale; mkdir -p $OUT
run() { name=$1; shift; python bench.py --gpus $N --no-cpu "$@" 2> $OUT/$name.err | grep '^{' > $OUT/$name.json; python - <<PY
import json
try:
    d=json.load(open("$OUT/$name.json")); print("$name", d["n_gpus"], d["scaling"], round(d["value"]), "pts/s", round(d["ms_per_step"],3), "ms", "frac_burst", round(d["roofline"]["frac_burst"],3))
except Exception as e: print("$name FAILED", e)
PY
}
if [ "$N" = "8" ]; then
  run burgers100d_weak_n8 --workload burgers100d --scaling weak
  run burgers100d_strong_n8 --workload burgers100d --scaling strong
  run burgers1000d_m256_n8 --workload burgers1000d --mc 256
  run burgers1000d_m4096_n8 --workload burgers1000d --mc 4096 --steps 2
  run hjb100d_weak_n8 --workload hjb100d
  python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29533 tests/dp_worker.py $OUT/dp_fit_n8.json > $OUT/dp_fit_n8.log 2>&1; cat $OUT/dp_fit_n8.json | head -c 400; echo
elif [ "$N" = "4" ]; then
  run burgers100d_strong_n4 --workload burgers100d --scaling strong
  run burgers1000d_m256_n4 --workload burgers1000d --mc 256
  python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29533 tests/dp_worker.py $OUT/dp_fit_n4.json > $OUT/dp_fit_n4.log 2>&1; cat $OUT/dp_fit_n4.json | head -c 300; echo
elif [ "$N" = "2" ]; then
  run burgers1000d_m256_n2 --workload burgers1000d --mc 256
else
  run burgers1000d_m1024_n1 --workload burgers1000d --mc 1024
  run burgers1000d_m4096_n1 --workload burgers1000d --mc 4096 --steps 2
  run gbm100d_n1 --workload gbm100d
fi
