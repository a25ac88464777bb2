#!/bin/bash
# Round-2 profile run (on the GPU box): default bench, FP64 peak, launch list, ncu --set full of the kernels.
# Results land in gpurun_out/; tools/ncu_summary.py turns the reports into profiles/*.md / kernel_counters.json.
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err; echo "bench rc=$?"
python -c "
import json, sys
sys.path.insert(0, '.')
from monorfs_b200 import capi
print(json.dumps(capi.bench_fp64(0, 4000)))" > gpurun_out/r02_fp64_peak.json 2> gpurun_out/r02_fp64_peak.err; echo "fp64 rc=$?"
CMD_L="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-check --e2e-steps 1"
$CMD_L > gpurun_out/r02_plain_l.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c4.csv $CMD_L > gpurun_out/r02_ncu_l.log 2>&1; echo "launch list rc=$?"
CMD_K="python bench.py --workload c4s --steps 2 --warmup 4 --settle 10 --no-secondary --no-cpu-baseline --no-parity-check --e2e-steps 1"
$CMD_K > gpurun_out/r02_plain_k.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_particle_update -s 15 -c 1 -f -o gpurun_out/r02_prof_c4s $CMD_K > gpurun_out/r02_ncu_k.log 2>&1; echo "ncu particle_update rc=$?"
CMD_T="python bench.py --workload c2x --steps 2 --warmup 3 --settle 0 --no-secondary --no-cpu-baseline --no-parity-check --e2e-steps 1"
$CMD_T > gpurun_out/r02_plain_t.log 2>&1 && \
ncu --set full --clock-control none -k regex:"k_copy_particles|k_normalize_resample" -s 6 -c 2 -f -o gpurun_out/r02_prof_tail $CMD_T > gpurun_out/r02_ncu_t.log 2>&1; echo "ncu tail rc=$?"
