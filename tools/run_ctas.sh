#!/bin/bash
# usage: tools/run_ctas.sh <workload> <warmup> <steps> <lib-variant or -> ctas...
wl=$1; wu=$2; st=$3; v=$4; shift 4
mkdir -p gpurun_out
lib=$PWD/monorfs_b200/_build/librbphd.so
[ "$v" != "-" ] && lib=$PWD/monorfs_b200/_build/librbphd_$v.so
for c in "$@"; do
  RBPHD_MAX_CTAS=$c RBPHD_LIB=$lib timeout 900 python bench.py --workload $wl --warmup $wu --steps $st \
      --no-cpu-baseline --no-secondary --e2e-steps 1 > gpurun_out/ctas_${v}_$c.json 2> gpurun_out/ctas_${v}_$c.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ctas_${v}_$c.json"))
    ph=d["roofline"]["phase_kcycles_per_particle"]
    print("$v ctas=$c", d["config"]["launch_shape"], "ms/step", round(d["ms_per_step"],3), "kcycles/particle", round(sum(ph.values()),1))
except Exception as e:
    print("$v $c failed", e)
PY
done
