#!/bin/bash
for v in 0 1 2; do echo "variant $v"; PHB_MMA_VARIANT=$v python tools/bench_configs.py cfg3 cfg4 2>&1 | tail -2 | python -c "
import sys, json
for line in sys.stdin:
    d = json.loads(line); print('  ', d['config'], 'prune_ms', round(d['prune_ms'],2), 'TFLOPs', round(d['fp64_TFLOPs'],2), 'lnl', d['lnl'])
"; done
