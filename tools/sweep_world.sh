#!/bin/bash
for g in 32 16 8; do for e in 32 16 8; do echo -n "G_gjk=$g G_epa=$e : "; NP_G_GJK=$g NP_G_EPA=$e python tools/prof_world.py ${1:-10000} 6 c2 | tail -1 | cut -c1-150; done; done
