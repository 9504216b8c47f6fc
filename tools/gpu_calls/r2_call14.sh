#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for n in 10000 40000; do python tools/prof_world.py $n 30 c2 2>&1 | tail -1 | cut -c1-140; done
python tools/prof_world.py 131072 10 c2 2>&1 | tail -1 | cut -c1-140
python tools/prof_world.py 200000 10 c2 2>&1 | tail -1 | cut -c1-140
