#!/bin/bash
# what the driver does at round end: the bench at N = 1, 2, 4, 8 on one box
for n in 1 2 4 8; do
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-stored 2>gpurun_out/scale_err_$n.log | tail -1 > gpurun_out/scale_$n.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 10 --warmup 3 2>gpurun_out/scale_err_$n.log | tail -1 > gpurun_out/scale_$n.json
  fi
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open("gpurun_out/scale_%s.json" % n))
    print("N=%s value=%.2f evals/s ms=%.2f e2e=%.2f clocks=%s" % (n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]))
except Exception as exc:
    print("N=%s failed: %r" % (n, exc)); print(open("gpurun_out/scale_err_%s.log" % n).read()[-1500:])
PY
done
