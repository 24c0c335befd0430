#!/bin/bash
# two GPUs: the new large tests, the distributed tests over NCCL (+ in-kernel peer sums), the bench at N = 2
timeout 600 python -m pytest tests/test_gpu_large.py tests/test_gpu_distributed.py -m gpu -x -q > gpurun_out/r3d_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r3d_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 3 --no-stored --no-configs --no-cpu-baseline 2> gpurun_out/r3d_bench2.err | tail -1 > gpurun_out/r3d_bench_2gpu.json; echo "bench2 rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r3d_bench_2gpu.json")); e = d["e2e"]
    print("N=2 value=%.2f ms=%.4f kernel_ms=%.4f e2e=%.2f lnl=%r scaling=%s rank_sum=%s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], e["value"], d["lnl"], d["scaling"], str(d.get("rank_sum"))[:60]))
    print("weak", json.dumps(d.get("weak_scaling"))[:300])
    print("sharded_parity", d.get("sharded_parity"))
except Exception as exc:
    print("failed", exc); print(open("gpurun_out/r3d_bench2.err").read()[-1500:])
PY
