#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r02_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r02_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r02_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_edges.py -m gpu -x -q -k "local_map or exact" > gpurun_out/r02_memcheck_edges.log 2>&1; echo "memcheck edges rc=$?"; tail -4 gpurun_out/r02_memcheck_edges.log
