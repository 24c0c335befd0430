#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2n_pytest.log 2>&1; echo pytest rc=$?; tail -8 gpurun_out/r2n_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-stored --no-configs --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo bench rc=$?; tail -c 600 gpurun_out/r2n_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2n_bench.json').read().strip().splitlines()[-1]); e=d['e2e']
print('value', d['value'], d['ms_per_step'], '| e2e', e['value'], e['ms_per_step'], e['tip_code_format'][:10], 'one-at-a-time', e['one_at_a_time']['value'], '| other', e['other_format']['value'], e['other_format']['one_at_a_time']['value'], 'lnl', d['lnl'], e['lnl'])
"
