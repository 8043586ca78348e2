#!/bin/bash
# One gpurun call: launch list of the default bench, full ncu capture of the dominant kernels on a reduced grid, sanitizer runs.
mkdir -p gpurun_out
B="--no-cpu-baseline --late-window 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 $B > gpurun_out/r02_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_x_relax -s 2 -c 1 -f -o gpurun_out/r02_xrelax python bench.py --workload lidar256 --steps 2 --warmup 2 --other-frames 0 $B > gpurun_out/r02_ncu_xrelax.log 2>&1; echo "ncu k_x_relax rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ray_resolve -s 2 -c 1 -f -o gpurun_out/r02_rayresolve python bench.py --steps 2 --warmup 2 --other-frames 0 $B > gpurun_out/r02_ncu_ray.log 2>&1; echo "ncu k_ray_resolve rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r02_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r02_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02_racecheck.log
ls -la gpurun_out/r02_*
