#!/bin/bash
# After tools/profile_r02.sh ran on the GPU box: turn gpurun_out/r02_* into the tracked files under profiles/.
set -e
cp gpurun_out/r02_bench_c4.json profiles/r02_bench_c4.json
cp gpurun_out/r02_fp64_peak.json profiles/fp64_peak.json
cp gpurun_out/r02_launches_c4.csv profiles/r02_ncu_launches_c4.csv
head -n 10 profiles/r02_ncu_c4s_summary.md > /tmp/hdr_c4s.md && cp /tmp/hdr_c4s.md profiles/r02_ncu_c4s_summary.md
python tools/ncu_summary.py gpurun_out/r02_prof_c4s.ncu-rep --particles 1480 \
    --capture "k_particle_update, c4s (1 480 particles), 16th launch (saturated maps), final tree" \
    --json profiles/kernel_counters.json --md profiles/r02_ncu_c4s_summary.md > /dev/null
head -n 7 profiles/r02_ncu_tail.md > /tmp/hdr_tail.md && cp /tmp/hdr_tail.md profiles/r02_ncu_tail.md
for i in 0 1; do python tools/ncu_summary.py gpurun_out/r02_prof_tail.ncu-rep --kernel-index $i --capture "kernel $i of the capture" --md profiles/r02_ncu_tail.md > /dev/null; done
python tools/ncu_lines.py gpurun_out/r02_prof_c4s.ncu-rep 40 > profiles/r02_ncu_c4s_hot_lines.txt
