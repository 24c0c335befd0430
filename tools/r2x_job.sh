#!/bin/bash
# (record of a dropped experiment: the one-category-per-lane layout and its PHB_PAIR_NO_CAT switch are not in the tree - DESIGN.md 3.1, profiles/r02h_cat_layout_experiment.txt)
# one category per lane (CAT) against the round-2a layout: GPU tests, then the shard-size probe with and without it
python -m pytest tests -m gpu -x -q > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2x_pytest.log
for cat in 0 1; do
  PHB_PAIR_NO_CAT=$cat python tools/strong_probe.py --sizes 125000,250000,1000000 --tag "no_cat=$cat" | tee -a gpurun_out/r2x_probe.jsonl
done
PHB_PAIR_PPT=2 python tools/strong_probe.py --sizes 125000 --tag "cat ppt2" | tee -a gpurun_out/r2x_probe.jsonl
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-stored > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2x_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "lnl", d["lnl"], d["cpu_baseline"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r2x_bench.err").read()[-2000:])
PY
