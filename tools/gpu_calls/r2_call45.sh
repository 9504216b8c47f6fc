#!/bin/bash
for m in mix sphere box; do NETPHYS_B200_LIB=$PWD/build/lib_e2dbg.so python tools/e2_debug_pairs.py $m 2>&1 | tail -1; done
