#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_suptab.py -x -q 2>&1 | tail -2
python tools/prof_world.py 1000000 6 c3 2>&1 | tail -1 | cut -c1-150
python tools/prof_pairs.py mix 2000000 3 2>&1 | tail -1 | cut -c1-100
python tools/prof_pairs.py box 2000000 3 2>&1 | tail -1 | cut -c1-100
python tools/prof_pairs.py sphere 1000000 3 2>&1 | tail -1 | cut -c1-100
