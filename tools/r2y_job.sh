#!/bin/bash
# ncu capture of the CAT walk at the headline size: summary + per-line hot spots
ncu --set full --clock-control none --import-source on -k regex:dna_pair -s 2 -c 1 -f -o gpurun_out/r2y_cat python tools/profile_prune.py --taxa 1000 --patterns 1000000 --evals 3 --lnl-only > gpurun_out/r2y_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/r2y_cat.ncu-rep > gpurun_out/r2y_cat_summary.txt 2>&1
python tools/ncu_hotspots.py gpurun_out/r2y_cat.ncu-rep 40 > gpurun_out/r2y_cat_hotspots.txt 2>&1
cat gpurun_out/r2y_cat_summary.txt
