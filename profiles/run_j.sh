#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_configs_gpu.py -m gpu -x -q -s > gpurun_out/pytest_cfg.log 2>&1; grep -E "PSNR|RGBA8|passed|failed|Error|error" gpurun_out/pytest_cfg.log | tail -30
python __graft_entry__.py smoke 2>&1 | tail -2
