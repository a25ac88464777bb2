#!/bin/bash
# usage: tools/run_variants.sh <workload> <warmup> <steps> name...   (run on the GPU box; results in gpurun_out/var_<name>.json)
wl=$1; wu=$2; st=$3; shift 3
mkdir -p gpurun_out
for v in "$@"; do
  RBPHD_LIB=$PWD/monorfs_b200/_build/librbphd_$v.so timeout 600 python bench.py --workload $wl --warmup $wu --steps $st \
      --no-cpu-baseline --no-secondary --e2e-steps 1 > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  echo "$v rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/var_$v.json"))
    print("$v", d["config"]["launch_shape"], "ms/step", round(d["ms_per_step"],3), "kernel_ms", round(d["roofline"]["kernel_ms"],3), "comps", round(d["config"]["mean_components_per_particle"],1))
except Exception as e:
    print("$v failed", e)
PY
done
