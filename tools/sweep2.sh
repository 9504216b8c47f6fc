#!/bin/bash
for mix in box mix; do for g in 1 4; do for e in 4 8 32; do echo -n "pairs $mix G_gjk=$g G_epa=$e : "; NP_G_GJK=$g NP_G_EPA=$e python tools/prof_pairs.py $mix 1000000 3 | tail -1 | cut -c1-90; done; done; done
for n in 100000 1000000; do for g in 1 8 32; do for e in 4 8 32; do echo -n "world $n G_gjk=$g G_epa=$e : "; NP_G_GJK=$g NP_G_EPA=$e python tools/prof_world.py $n 4 c3 | tail -1 | cut -c50-170; done; done; done
