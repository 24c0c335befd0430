#!/bin/bash
python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo bench rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/r2g_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['rel_diff'])
print({k:(v['evals_per_s'], v['roofline']['frac']) for k,v in d['with_stored_partials'].items()})
for k,v in d['configs'].items(): print(k, v.get('value'), v.get('lnl_ms'), v.get('up_pass_ms'), v.get('derivative_pass_ms'), v['parity']['ok'])
"
P="python tools/profile_prune.py --taxa 1000 --patterns 1000000 --evals 3 --lnl-only"
$P > gpurun_out/r2g_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dna_pair_kernel -s 1 -c 1 -f -o gpurun_out/r02b_pair_lnl_sym $P > gpurun_out/r2g_ncu.log 2>&1
ls -la gpurun_out/r02b_pair_lnl_sym.ncu-rep
