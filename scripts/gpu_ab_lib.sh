#!/bin/bash
# One gpurun call: GPU tests with the in-tree library, then the default EXACT bench alternating between two prebuilt libraries
# (ab_tmp/libfiesta_<name>.so copied over the in-tree one).  Usage: scripts/gpu_ab_lib.sh <tag> <name> <name> ...
tag=$1; shift
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
k=0
for v in "$@"; do
  k=$((k+1))
  cp ab_tmp/libfiesta_$v.so fiesta_b200/lib/libfiesta_b200.so
  timeout 600 python bench.py --no-cpu-baseline --late-window 0 --other-frames 0 --no-host-mirror > gpurun_out/${tag}_${v}_$k.json 2> gpurun_out/${tag}_${v}_$k.err
  echo "$v rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_${v}_$k.json"))
print("  ms/frame", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "k_x_relax", round(d["kernels"]["k_x_relax"]["ms_per_step"],3), "exp_equal", d["expansions_equal_reference"])
PY
done
