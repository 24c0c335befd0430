#!/bin/bash
# what the driver does at round end: the bench at N = 1, 2, 4, 8 on one box (strong scaling of the 1k x 1M alignment)
mkdir -p gpurun_out
for n in ${SCALE_NS:-1 2 4 8}; do
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-stored --no-configs 2>gpurun_out/scale_err_$n.log | tail -1 > gpurun_out/scale_$n.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 10 --warmup 3 2>gpurun_out/scale_err_$n.log | tail -1 > gpurun_out/scale_$n.json
  fi
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open("gpurun_out/scale_%s.json" % n))
    w = d.get("weak_scaling") or {}
    c5 = (d.get("configs") or {}).get("cfg5") or {}
    print("N=%s strong value=%.2f evals/s ms=%.3f kernel_ms=%.3f e2e=%.2f (%.3f ms) weak=%s lnl=%r parity=%s cfg5_sweep_ms=%s clocks=%s" % (
        n, d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"],
        w.get("value_1M_pattern_evals_per_s"), d["lnl"], (d.get("sharded_parity") or {}).get("ok"), c5.get("sweep_ms"), d["clocks"]))
except Exception as exc:
    print("N=%s failed: %r" % (n, exc)); print(open("gpurun_out/scale_err_%s.log" % n).read()[-1500:])
PY
done
