#!/bin/bash
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-300
