#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --workload depth256 > gpurun_out/r2_depth256.json 2> gpurun_out/r2_depth256.err; echo "depth256 rc=$?"
timeout 1500 python bench.py --workload lidar1024 --steps 10 --warmup 3 --cpu-frames 2 > gpurun_out/r2_lidar1024.json 2> gpurun_out/r2_lidar1024.err; echo "lidar1024 rc=$?"
timeout 900 python bench.py --workload stress256 --steps 3 --warmup 1 --cpu-frames 1 > gpurun_out/r2_stress256.json 2> gpurun_out/r2_stress256.err; echo "stress256 rc=$?"
python - <<'PY'
import json
for n in ("depth256","lidar1024","stress256"):
    d=json.load(open("gpurun_out/r2_%s.json"%n))
    print(n,"EXACT ms %.2f value %.3e e2e %.3e | FAST ms %.2f value %.3e | cpu %.3e | parity exact %s fast dist %d reach %d maxerr %.3f tie %d finite %d"%(d["ms_per_step"],d["value"],d["e2e"]["value"],d["fast_mode"]["ms_per_step"],d["fast_mode"]["value"],d["cpu_baseline"]["value"],[d["parity"]["exact"][k] for k in ("dist","cobs_tie","cobs_nontie","occ")],d["parity"]["fast"]["dist"],d["parity"]["fast"]["reach"],d["parity"]["fast"]["dist_max_err"],d["parity"]["fast"]["cobs_tie"],d["parity"]["fast"]["finite"]))
PY
