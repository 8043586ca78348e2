#!/bin/bash
# Last GPU call of the round: tests on the in-tree library (falls back to ab_tmp/libfiesta_<fallback>.so if they fail), short A/B
# benches of the prebuilt libraries, the full default bench, the ncu launch list and one full capture of k_x_relax.
# Usage: scripts/gpu_final.sh <tag> <fallback> <name> [name ...]
tag=$1; fb=$2; shift 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
cp fiesta_b200/lib/libfiesta_b200.so /tmp/intree.so
B="--no-cpu-baseline --late-window 0 --other-frames 0 --no-host-mirror"
for v in "$@"; do
  cp ab_tmp/libfiesta_$v.so fiesta_b200/lib/libfiesta_b200.so
  timeout 300 python bench.py $B > gpurun_out/${tag}_ab_$v.json 2> gpurun_out/${tag}_ab_$v.err
  echo "$v rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_ab_$v.json"))
print("  ms/frame", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "k_x_relax", round(d["kernels"]["k_x_relax"]["ms_per_step"],3), "exp_equal", d["expansions_equal_reference"])
PY
done
if [ $rc -eq 0 ]; then cp /tmp/intree.so fiesta_b200/lib/libfiesta_b200.so; echo "final runs: in-tree library"; else cp ab_tmp/libfiesta_$fb.so fiesta_b200/lib/libfiesta_b200.so; echo "final runs: FALLBACK $fb"; fi
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("FULL ms/frame", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "value %.4e"%d["value"], "parity", d.get("parity",{}).get("exact"), "late", d.get("late_window",{}).get("ms_per_step"), "fast", d.get("fast_mode",{}).get("ms_per_step"), "cpu %.3e"%d["cpu_baseline"]["value"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --late-window 0 --no-host-mirror > gpurun_out/${tag}_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_x_relax -s 3 -c 1 -f -o gpurun_out/${tag}_xrelax512 python bench.py --steps 2 --warmup 2 $B > gpurun_out/${tag}_ncu_xrelax512.log 2>&1; echo "ncu k_x_relax rc=$?"
