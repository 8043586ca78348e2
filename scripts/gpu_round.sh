#!/bin/bash
# One gpurun call: exact-mode tests first (bounded), then stats.  Usage: scripts/gpu_round.sh <tag> [pytest -k expression] [frames]
tag=${1:-x}
kexp=${2:-exact}
nfr=${3:-8}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q -k "$kexp" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
FIESTA_DEBUG_X=1 timeout 100 python scripts/xstat.py lidar512 $nfr > gpurun_out/${tag}_xstat512.log 2>&1; echo "xstat rc=$?"
grep -v "^\[x\] gen\|seeds" gpurun_out/${tag}_xstat512.log | tail -32
