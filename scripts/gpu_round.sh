#!/bin/bash
# One gpurun call: exact-mode tests first (bounded), then stats.  Usage: scripts/gpu_round.sh <tag> [pytest -k expression]
tag=${1:-x}
kexp=${2:-exact}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "$kexp" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
FIESTA_DEBUG_X=1 timeout 300 python scripts/xstat.py lidar512 8 > gpurun_out/${tag}_xstat512.log 2>&1; echo "xstat rc=$?"
grep -v "^\[x\] gen" gpurun_out/${tag}_xstat512.log | tail -32
