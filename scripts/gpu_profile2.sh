#!/bin/bash
# One gpurun call: launch list of the default bench, full ncu capture of k_x_relax on the 512^3 workload, memcheck of the smoke entry.
mkdir -p gpurun_out
B="--no-cpu-baseline --late-window 0 --no-host-mirror"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 1 $B > gpurun_out/r02b_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_x_relax -s 3 -c 1 -f -o gpurun_out/r02b_xrelax512 python bench.py --steps 2 --warmup 2 --other-frames 0 $B > gpurun_out/r02b_ncu_xrelax512.log 2>&1; echo "ncu k_x_relax rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r02b_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02b_memcheck.log
ls -la gpurun_out/r02b_*
