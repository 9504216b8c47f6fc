#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/prof_world.py 1000000 6 c3 2>&1 | tail -1 | cut -c1-170
python tools/prof_world.py 131072 20 c2 2>&1 | tail -1 | cut -c1-170
python tools/prof_pairs.py mix 2000000 3 2>&1 | tail -1 | cut -c1-100
python tools/prof_pairs.py box 2000000 3 2>&1 | tail -1 | cut -c1-100
