#!/bin/bash
# One GPU session of the development loop: GPU tests, beam-only throughput on L2-cold inputs, parity sweep.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [sweep_pairs] [extra]
TAG=${1:-x}; SWEEP=${2:-192}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/${TAG}_pytest.txt
cat gpurun_out/${TAG}_pytest.txt
( POB_DEBUG_VERBOSE=1 POB_PROF_UNIQUE=1024 timeout 600 python tools/prof_pair.py 2664 2 2>&1 | tail -6 ) > gpurun_out/${TAG}_prof.txt
cat gpurun_out/${TAG}_prof.txt
if [ "$SWEEP" -gt 0 ]; then
  ( timeout 900 python tools/parity_sweep.py $SWEEP 12000 2>&1 | tail -3 ) > gpurun_out/${TAG}_sweep.txt
  cat gpurun_out/${TAG}_sweep.txt
fi
