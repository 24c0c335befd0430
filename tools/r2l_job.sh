#!/bin/bash
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600+$1+$2)) bench.py --gpus $1 --steps 10 --warmup 3 --no-weak --no-configs --chunks $2 2>gpurun_out/r2l_err_$1_$2.log | tail -1 > gpurun_out/r2l_$1_$2.json
python - $1 $2 <<'PY'
import json, sys
n, ch = sys.argv[1:3]
try:
    d = json.load(open("gpurun_out/r2l_%s_%s.json" % (n, ch))); e = d["e2e"]
    print("N=%s chunks=%s value=%.2f ms=%.3f | e2e=%.2f (%.3f ms, %.1f GB/s/GPU, %s) other=%.2f (%.3f ms)" % (n, ch, d["value"], d["ms_per_step"], e["value"], e["ms_per_step"], e["host_to_device_gbs_per_gpu"], e["tip_code_format"][:12], e["other_format"]["value"], e["other_format"]["ms_per_step"]))
except Exception as exc:
    print("N=%s chunks=%s failed %r" % (n, ch, exc)); print(open("gpurun_out/r2l_err_%s_%s.log" % (n, ch)).read()[-800:])
PY
}
run 8 0; run 8 4; run 8 16; run 4 0; run 4 8
