#!/bin/bash
NETPHYS_B200_LIB=$PWD/build/lib_e2dbg.so python tools/e2_debug.py 300000 2>&1 | tail -3
