"""Summarise one `ncu --set full` capture of a kernel: a markdown table of the metrics the roofline discussion
uses, and (for k_particle_update) profiles/kernel_counters.json -- DRAM bytes and FP64 thread-instructions per
particle-frame, tagged with the hash of the device sources so that bench.py only uses them on the same tree.

    python tools/ncu_summary.py report.ncu-rep --particles 1480 --capture "c4s, launch 6" \
        [--json profiles/kernel_counters.json] [--md profiles/r02_ncu_xxx.md] [--kernel-index 0]
"""
import argparse
import csv
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 wavefronts % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit rate %"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate %"),
    ("lts__t_sectors.sum", "L2 sectors"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("smsp__sass_inst_executed_op_local_ld.sum", "local-memory loads (warp inst)"),
    ("smsp__sass_inst_executed_op_local_st.sum", "local-memory stores (warp inst)"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier (warps / issue)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: LG throttle"),
]


def load(rep, index):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    vals = rows[2 + index]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(d, key, default=None):
    if key not in d:
        return default
    try:
        return float(d[key][0].replace(",", ""))
    except ValueError:
        return default


def to_bytes(d, key):
    v, u = d.get(key, ("0", "byte"))
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--particles", type=int, default=0, help="particle-frames the captured launch processed")
    ap.add_argument("--capture", default="")
    ap.add_argument("--json", default="")
    ap.add_argument("--md", default="")
    ap.add_argument("--kernel-index", type=int, default=0)
    a = ap.parse_args()
    d = load(a.report, a.kernel_index)
    name = d.get("Kernel Name", ("?", ""))[0]
    lines = ["| metric | value |", "|---|---|", "| kernel | `%s` |" % name]
    for k, label in KEYS:
        if k in d:
            v, u = d[k]
            lines.append("| %s | %s %s |" % (label, v, u))
    cyc = num(d, "smsp__cycles_elapsed.avg") or num(d, "sm__cycles_elapsed.max")
    fp64 = None
    per = [num(d, "smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op) for op in ("dadd", "dmul", "dfma")]
    if cyc and all(p is not None for p in per):
        fp64 = sum(per) * cyc
        lines.append("| FP64 thread-instructions (DADD + DMUL + DFMA) | %.4g (DADD %.3g, DMUL %.3g, DFMA %.3g per cycle) |"
                     % (fp64, per[0], per[1], per[2]))
    dram = to_bytes(d, "dram__bytes_read.sum") + to_bytes(d, "dram__bytes_write.sum")
    dur_v, dur_u = d.get("gpu__time_duration.sum", ("0", "ms"))
    dur_s = float(dur_v.replace(",", "")) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(dur_u, 1e-3)
    if dur_s > 0:
        lines.append("| DRAM bytes read + written | %.4g GB (%.1f GB/s) |" % (dram / 1e9, dram / dur_s / 1e9))
    if a.particles:
        lines.append("| particle-frames in this launch | %d |" % a.particles)
        lines.append("| DRAM bytes per particle-frame | %.4g MB |" % (dram / a.particles / 1e6))
        if fp64:
            lines.append("| FP64 thread-instructions per particle-frame | %.4g |" % (fp64 / a.particles))
    text = "\n".join(lines) + "\n"
    print(text)
    if a.md:
        with open(a.md, "a") as fh:
            fh.write("\n### %s\n\n" % (a.capture or os.path.basename(a.report)) + text)
    if a.json and a.particles:
        from monorfs_b200 import build
        rec = {"source_hash": build.source_hash(), "capture": a.capture, "kernel": name, "particle_frames": a.particles,
               "duration_ms": dur_s * 1e3, "dram_bytes_per_particle_frame": dram / a.particles,
               "fp64_thread_inst_per_particle_frame": (fp64 / a.particles) if fp64 else None,
               "fp64_pipe_active_pct": num(d, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
               "warp_inst_per_particle_frame": (num(d, "smsp__inst_executed.sum") or 0) / a.particles}
        with open(a.json, "w") as fh:
            json.dump(rec, fh, indent=1)
            fh.write("\n")


if __name__ == "__main__":
    main()
