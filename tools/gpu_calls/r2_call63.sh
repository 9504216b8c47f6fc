#!/bin/bash
# the libraries rebuilt from scratch by build(): GPU suite and smoke once more
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
