#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py --workload lidar1024 --steps 10 --warmup 3 --cpu-frames 2 > gpurun_out/r2_lidar1024.json 2> gpurun_out/r2_lidar1024.err; echo "lidar1024 rc=$?"; tail -c 1500 gpurun_out/r2_lidar1024.json; tail -3 gpurun_out/r2_lidar1024.err
timeout 900 python bench.py --workload stress256 --steps 3 --warmup 1 --cpu-frames 1 > gpurun_out/r2_stress256.json 2> gpurun_out/r2_stress256.err; echo "stress256 rc=$?"; tail -c 1500 gpurun_out/r2_stress256.json; tail -3 gpurun_out/r2_stress256.err
