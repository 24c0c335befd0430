#!/bin/bash
# symmetric-P-block lnL-only walk: parity (whole GPU suite) and A/B timing
(python tools/strong_probe.py --tag sym --sizes 125000,250000,1000000; PHB_PAIR_FULL_P=1 python tools/strong_probe.py --tag fullP --sizes 125000,250000,1000000) > gpurun_out/r2f_probe.jsonl 2> gpurun_out/r2f_probe.err
cat gpurun_out/r2f_probe.jsonl | cut -c1-260; tail -c 400 gpurun_out/r2f_probe.err
timeout 1300 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo pytest rc=$?; tail -12 gpurun_out/r2f_pytest.log
python tools/bench_configs.py cfg1 cfg3 cfg4 --reps 3 > gpurun_out/r2f_cfg.jsonl 2> gpurun_out/r2f_cfg.err
python -c "
import json
for l in open('gpurun_out/r2f_cfg.jsonl'):
    d=json.loads(l); print(d['config'], d['lnl'], d.get('lnl_ms'), d.get('prune_kernel_ms'), d.get('lnl_only_us_per_eval'), d['parity']['ok'])
"
