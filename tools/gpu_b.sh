#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "segment_compositing or alpha_importance" 2>&1 | tail -15
timeout 600 python -m pytest tests/test_stage2_gpu.py -q -m gpu -s 2>&1 | tail -25
