#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r2i_pytest.log
bash tools/scale_run.sh
