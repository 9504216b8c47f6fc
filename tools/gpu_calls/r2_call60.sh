#!/bin/bash
# final validation at HEAD: the GPU suite, smoke(), and a full ncu capture of the grid broad phase's kernels
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
NP_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"k_body_boxes|k_grid_count|k_grid_fill|k_cand" -s 5 -c 5 -o gpurun_out/r4n_broad -f python tools/prof_grid.py 1000000 2.0 gridonly > gpurun_out/r4n_ncu.log 2>&1
tail -1 gpurun_out/r4n_ncu.log
