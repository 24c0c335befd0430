#!/bin/bash
# state of HEAD on one GPU: GPU tests, the full default bench line, launch list, ncu capture of the shipped headline kernel
python -m pytest tests -m gpu -x -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2v_pytest.log
python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2v_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "kernel", d["roofline"]["kernel"][:60], d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r2v_bench.err").read()[-2000:])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2v_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-stored --no-e2e > gpurun_out/r2v_ncu_list.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:dna_pair -s 2 -c 1 -f -o gpurun_out/r2v_pair python tools/profile_prune.py --taxa 1000 --patterns 1000000 --evals 3 --lnl-only > gpurun_out/r2v_ncu_full.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_summary.py gpurun_out/r2v_pair.ncu-rep > gpurun_out/r2v_pair_summary.txt 2>&1; head -30 gpurun_out/r2v_pair_summary.txt
ls -la gpurun_out/
