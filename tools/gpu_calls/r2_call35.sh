#!/bin/bash
python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | grep -E "Error|assert|passed|failed" | head -12
