#!/bin/bash
# DMMA kernels with the category count known at compile time and strength-reduced row copies
python -m pytest tests -m gpu -x -q > gpurun_out/r3b_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r3b_pytest.log
python tools/bench_configs.py cfg3 cfg4 --reps 5 2>&1 | tee gpurun_out/r3b_configs.jsonl | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print(d.get('config', '')[:40], {k: (round(d[k], 3) if isinstance(d[k], float) else d[k]) for k in ('lnl_ms', 'prune_kernel_ms', 'up_pass_ms', 'first_derivative_pass_ms', 'derivative_pass_ms', 'sweep_ms', 'lnl') if k in d}, d.get('fp64'), d.get('parity'))
"
