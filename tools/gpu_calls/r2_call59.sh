#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -k "grid or swept" 2>&1 | tail -3
python tools/prof_ground.py 512 2>&1 | tail -1 | cut -c1-300
