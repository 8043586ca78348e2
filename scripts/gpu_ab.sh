#!/bin/bash
# One gpurun call: GPU test suite, then the default bench (EXACT only, no CPU leg) once per value of an environment switch.
# Usage: scripts/gpu_ab.sh <tag> <ENV_NAME> <value> [value ...]
tag=$1; var=$2; shift 2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
for v in "$@"; do
  env $var=$v timeout 600 python bench.py --no-cpu-baseline --late-window 0 --other-frames 0 --no-host-mirror > gpurun_out/${tag}_${var}_${v}.json 2> gpurun_out/${tag}_${var}_${v}.err
  echo "$var=$v rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_${var}_${v}.json"))
print("  ms/frame", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "k_x_relax", round(d["kernels"]["k_x_relax"]["ms_per_step"],3), "exp_equal", d["expansions_equal_reference"])
PY
done
