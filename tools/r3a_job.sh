#!/bin/bash
# store-walk build variants on one box (v0 = HEAD, vA = whole-tile staging + rescale behind the shapes, vB = whole-tile staging +
# rescale inside the shapes), then the GPU tests and the cfg5 shard with the branchy sum-table step
for v in v0 vA vB v0 vA vB; do
  PHB_LIBRARY=$PWD/phylo_utils_b200/libphylo_b200_$v.so python tools/store_probe.py --tag $v | tee -a gpurun_out/r3a_store_variants.jsonl
done
python -m pytest tests -m gpu -x -q > gpurun_out/r3a_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r3a_pytest.log
python tools/bench_configs.py cfg5 --reps 5 2>&1 | tee gpurun_out/r3a_configs.jsonl | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print({k: d[k] for k in ('lnl_ms', 'prune_kernel_ms', 'up_pass_ms', 'derivative_pass_ms', 'sweep_ms', 'lnl') if k in d}, d.get('parity'))
"
