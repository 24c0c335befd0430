#!/bin/bash
# in-kernel peer sum: two ranks (one GPU: both on cuda:0 over CUDA IPC; two GPUs: NVLink), then the whole GPU suite
timeout 600 python -m pytest tests/test_gpu_distributed.py -q -x -s > gpurun_out/r2p_dist.log 2>&1; echo dist rc=$?; grep -a "peer_sums\|passed\|failed\|Error" gpurun_out/r2p_dist.log | tail -8
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2p_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2p_pytest.log
