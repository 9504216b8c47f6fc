#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): tests, the headline bench, its reference arm, then the two ncu passes of B200_PROFILING.md.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2a_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2a_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err; echo "ref rc=$?"
python tools/prof_world.py 1000000 6 c3 > gpurun_out/r2a_plain_1m.log 2>&1; cat gpurun_out/r2a_plain_1m.log
python tools/prof_world.py 10000 8 c2 > gpurun_out/r2a_plain_10k.log 2>&1; cat gpurun_out/r2a_plain_10k.log
ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 200 --csv --log-file gpurun_out/r2a_launches_1m.csv python tools/prof_world.py 1000000 4 c3 > gpurun_out/r2a_ncu1.log 2>&1
ncu --set full --clock-control none --cache-control none --graph-profiling node --import-source on -k regex:"k_gjk|k_epa|k_integrate|k_resolve|k_link" -s 27 -c 9 -o gpurun_out/r2a_full_1m python tools/prof_world.py 1000000 4 c3 > gpurun_out/r2a_ncu2.log 2>&1
tail -2 gpurun_out/r2a_ncu2.log
