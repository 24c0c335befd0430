#!/bin/bash
# one GPU call: fp64 probes, cfg3/cfg4 timings, the whole GPU test-suite
python -c "
import ctypes
from phylo_utils_b200._lib import lib, check
for kind in (0,1,2,3):
    out=ctypes.c_double(0); check(lib().phb_op_fp64_peak(0, kind, ctypes.byref(out))); print('kind', kind, out.value)
" > gpurun_out/r2e_peaks.txt 2>&1; cat gpurun_out/r2e_peaks.txt
python tools/bench_configs.py cfg4 cfg3 --reps 3 > gpurun_out/r2e_cfg.jsonl 2> gpurun_out/r2e_cfg.err; tail -c 300 gpurun_out/r2e_cfg.err
python -c "
import json
for l in open('gpurun_out/r2e_cfg.jsonl'):
    d=json.loads(l); print(d['config'], d['lnl'], d['lnl_ms'], d['prune_kernel_ms'], d.get('up_pass_ms'), d.get('first_derivative_pass_ms'), d.get('derivative_pass_ms'), d['parity']['ok'])
"
timeout 1300 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/r2e_pytest.log
