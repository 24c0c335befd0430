#!/bin/bash
# rescale behind the row shapes (no register shuffles) + whole-tile staging in the store / pre-order walks
python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2z_pytest.log
python tools/strong_probe.py --sizes 125000,250000,1000000 --tag "rescale-after-join" | tee gpurun_out/r2z_probe.jsonl
python tools/bench_configs.py cfg5 cfg1 --reps 5 2>&1 | tee gpurun_out/r2z_configs.jsonl | cut -c1-900
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2z_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "lnl", d["lnl"])
    print("stored", json.dumps(d.get("with_stored_partials"))[:1500])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r2z_bench.err").read()[-2000:])
PY
