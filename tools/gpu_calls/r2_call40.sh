#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q 2>&1 | tail -2
python tools/prof_world.py 1000000 6 c3 2>&1 | tail -1 | cut -c1-150
python tools/prof_world.py 131072 20 c2 2>&1 | tail -1 | cut -c1-150
