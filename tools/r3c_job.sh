#!/bin/bash
# final captures of the round (summaries, per-line hot spots and per-opcode profiles are made on the box; the .ncu-rep files
# are too big to travel back together): headline walk, pre-order walk + edge kernel on the cfg5 shard, the two DMMA lnL kernels;
# then the full default bench line and its launch list
T=/tmp/r3c; mkdir -p $T
ncu --set full --clock-control none --import-source on -k regex:dna_pair -s 2 -c 1 -f -o $T/pair python tools/profile_prune.py --taxa 1000 --patterns 1000000 --evals 3 --lnl-only > gpurun_out/r3c_ncu_pair.log 2>&1; echo "ncu pair rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"dna_up_kernel|dna_edge_st_kernel" -c 2 -f -o $T/up python tools/bench_configs.py cfg5 --reps 1 > gpurun_out/r3c_ncu_up.log 2>&1; echo "ncu up rc=$?"
ncu --set full --clock-control none --import-source on -k regex:mma_prune -c 1 -f -o $T/cfg4 python tools/bench_configs.py cfg4 --reps 1 > gpurun_out/r3c_ncu_cfg4.log 2>&1; echo "ncu cfg4 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:mma_prune -c 1 -f -o $T/cfg3 python tools/bench_configs.py cfg3 --reps 1 > gpurun_out/r3c_ncu_cfg3.log 2>&1; echo "ncu cfg3 rc=$?"
for k in pair up cfg4 cfg3; do
  python tools/ncu_summary.py $T/$k.ncu-rep > gpurun_out/r3c_${k}_summary.txt 2>&1
  python tools/ncu_hotspots.py $T/$k.ncu-rep 30 > gpurun_out/r3c_${k}_hotspots.txt 2>&1
done
python tools/ncu_sass_profile.py $T/pair.ncu-rep $((15625*999)) dna_pair > gpurun_out/r3c_pair_sass_profile.txt 2>&1
python tools/ncu_sass_profile.py $T/up.ncu-rep $((1953*1999)) dna_up_kernel > gpurun_out/r3c_up_sass_profile.txt 2>&1
python tools/ncu_sass_profile.py $T/cfg4.ncu-rep $((391*99)) mma_prune > gpurun_out/r3c_cfg4_sass_profile.txt 2>&1
python tools/ncu_sass_profile.py $T/cfg3.ncu-rep $((782*498)) mma_prune > gpurun_out/r3c_cfg3_sass_profile.txt 2>&1
python bench.py > gpurun_out/r3c_bench.json 2> gpurun_out/r3c_bench.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3c_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-stored --no-e2e > gpurun_out/r3c_ncu_list.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r3c_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], d["roofline"]["kernel"][:70], d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r3c_bench.err").read()[-2000:])
PY
du -sh gpurun_out
