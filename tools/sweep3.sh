#!/bin/bash
for n in 10000 100000; do for g in 8 32; do for e in 32; do echo -n "world $n G_gjk=$g G_epa=$e : "; NP_G_GJK=$g NP_G_EPA=$e python tools/prof_world.py $n 5 c2 | tail -1 | cut -c50-170; done; done; done
