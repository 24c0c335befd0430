#!/bin/bash
# two GPUs: the distributed tests over NCCL + NVLink peer sums, then the bench at N=2 with the in-kernel sum and with all_reduce
for peer in 1 0 1 0; do
  PHB_PEER_SUM=$peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600+peer)) bench.py --gpus 2 --steps 20 --warmup 3 --no-stored --no-configs --no-cpu-baseline --no-weak 2>gpurun_out/r2q_err_$peer.log | tail -1 > gpurun_out/r2q_bench_peer$peer.json
  python - $peer <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/r2q_bench_peer%s.json" % sys.argv[1])); e = d["e2e"]
    print("peer=%s value=%.2f ms=%.4f kernel_ms=%.4f e2e=%.2f one_at_a_time=%.2f lnl=%r rank_sum=%s" % (sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], e["value"], e["one_at_a_time"]["value"], d["lnl"], d["rank_sum"][:40]))
except Exception as exc:
    print("failed", exc); print(open("gpurun_out/r2q_err_%s.log" % sys.argv[1]).read()[-1500:])
PY
done
