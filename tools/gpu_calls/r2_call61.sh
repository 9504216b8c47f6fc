#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q -k "grid or swept or shard" 2>&1 | tail -3
python tools/prof_grid.py 1000000 2.0 gridonly 2>&1 | tail -1 | cut -c1-200
python tools/prof_grid.py 1000000 2.0 gridonly 2>&1 | tail -1 | cut -c1-130
