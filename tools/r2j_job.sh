#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest.log 2>&1; echo pytest rc=$?; tail -6 gpurun_out/r2j_pytest.log
python bench.py --steps 10 --warmup 3 --no-stored --no-configs > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo bench rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], 'nibble', d['e2e']['two_codes_per_byte']['value'], 'lnl', d['lnl'], d['e2e']['lnl'])
"
for v in 0 2; do PHB_MMA_VARIANT=$v python tools/bench_configs.py cfg3 cfg4 --reps 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('variant $v', d['config'][:5], d['lnl'], d['lnl_ms'], d['prune_kernel_ms'], d.get('up_pass_ms'), d['parity']['ok'])
"; done
