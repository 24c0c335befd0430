#!/bin/bash
# eight GPUs, final build: strong scaling of the headline alignment (in-kernel peer sums over NVLink), trimmed bench
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 8 --steps 20 --warmup 3 --no-stored --no-configs --no-cpu-baseline --no-weak 2> gpurun_out/r3g_err8.log | tail -1 > gpurun_out/r3g_bench_8gpu.json; echo "bench8 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29812 bench.py --gpus 4 --steps 20 --warmup 3 --no-stored --no-configs --no-cpu-baseline --no-weak 2> gpurun_out/r3g_err4.log | tail -1 > gpurun_out/r3g_bench_4gpu.json; echo "bench4 rc=$?"
python - <<'PY'
import json
for n in (8, 4):
    try:
        d = json.load(open("gpurun_out/r3g_bench_%dgpu.json" % n)); e = d["e2e"]
        print("N=%d value=%.2f ms=%.4f kernel_ms=%.4f e2e=%.2f lnl=%r rank_sum=%s parity=%s clocks=%s" % (n, d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], e["value"], d["lnl"], str(d.get("rank_sum"))[:50], (d.get("sharded_parity") or {}).get("ok"), d["clocks"]))
    except Exception as exc:
        print("N=%d failed" % n, exc); print(open("gpurun_out/r3g_err%d.log" % n).read()[-1500:])
PY
