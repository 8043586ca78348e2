#!/bin/bash
# Parameter sweep of the order-exact relaxation (work-list thresholds), 8 lidar512 frames each.
mkdir -p gpurun_out
for cfg in "16384 16384" "16384 65536" "65536 16384" "65536 65536"; do
  set -- $cfg
  FIESTA_DEBUG_X=1 FIESTA_X_DENSE=$1 FIESTA_X_SMALL=$2 timeout 120 python scripts/xstat.py lidar512 8 > gpurun_out/sweep_$1_$2.log 2>&1
  echo "dense_min=$1 small_max=$2: $(grep -o "'ms_update_esdf': [0-9.]*" gpurun_out/sweep_$1_$2.log | awk '{s+=$2} END {print s/7}') ms avg over frames 1-7"
done
grep "phases" gpurun_out/sweep_16384_16384.log | tail -2
