#!/bin/bash
# round 2, GPU call 25 (1 GPU): re-entry validation of the committed state (full GPU suite without -x, smoke)
mkdir -p gpurun_out
echo "== full GPU suite =="; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/call25_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/call25_tests.log
echo "== smoke =="; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
