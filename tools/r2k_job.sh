#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_distributed.py -q > gpurun_out/r2k_dist.log 2>&1; echo dist rc=$?; tail -3 gpurun_out/r2k_dist.log
SCALE_NS="8 4" bash tools/scale_run.sh
