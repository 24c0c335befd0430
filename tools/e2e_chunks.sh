#!/bin/bash
# e2e throughput of bench.py as a function of the host->device pipeline depth
python tools/h2d_probe.py
for c in 4 8 16 32; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stored --chunks $c 2>&1 | tail -1 > /tmp/line.json
  python - "$c" <<'PY'
import json, sys
d = json.load(open("/tmp/line.json"))
print("chunks", sys.argv[1], "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "ms", round(d["e2e"]["ms_per_step"], 2))
PY
done
