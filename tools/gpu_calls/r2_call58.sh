#!/bin/bash
python tools/prof_ground.py 512 2>&1 | tail -2 | cut -c1-300
python tools/prof_ground.py 128 2>&1 | tail -1 | cut -c1-300
