#!/usr/bin/env python
"""bench.py -- RB-PHD SLAM per-frame update throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c2|c3|tiny] [--impl reference]

One "step" = one frame = PHDNavigator.Update + SlamUpdate (PHD:295,323) over all particles of the
workload, including weight normalisation / ESS test / resampling when it triggers.
  value  particle-frames/s, whole job, device-resident loop (inputs already in HBM), CUDA events
  e2e    the same metric through the reference-facing C ABI with HOST buffers each frame
         (rbphd_update + rbphd_slam_update + best-map read-back), wall clock
N > 1 (torchrun): particles sharded by rank (strong scaling: the workload's particle count is fixed),
one weight allgather per frame, map migration only on resampling frames.
--impl reference: the oracle (CPU restatement of the C# reference; the C# itself cannot run in this
image) on all host cores, on a bounded particle sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle_frames_per_sec"
UNIT = "particle-frames/s"
MAPPING_ONLY = {"c3"}
RECORD_BYTES = 80          # algorithmic FP64 record: weight + mean[3] + symmetric cov[6]
PER_PARTICLE_BYTES = 128   # pose read+written (112) + weight read+written (16)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as fh:
            for line in fh:
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for nm, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(np.max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def algorithmic_bytes(counters):
    return RECORD_BYTES * (counters["comps_in"] + counters["comps_out"]) + PER_PARTICLE_BYTES * counters["particle_frames"]


def cpu_baseline(workload, sample_particles, frames, seed):
    """The oracle (a port: CPU restatement of the C# reference) on the host cores, bounded sample."""
    from oracle import orc
    from monorfs_b200 import synth
    wl = synth.WORKLOADS[workload]
    cores = os.cpu_count() or 1
    S = max(1, min(sample_particles, wl["P"]))
    sc = synth.make_workload(workload, seed=seed, P=S)
    sc.params["nthreads"] = cores
    cfg = orc.make_config(sc.params)
    mapping = workload in MAPPING_ONLY
    nav = orc.Navigator(cfg, S, sc.poses[0], only_mapping=mapping)
    for i in range(S):
        nav.set_pose(i, sc.poses[i])
        nav.set_map(i, sc.map_w, sc.map_m, sc.map_P)
    times = []
    for _ in range(frames):
        fr = sc.next_frame()
        t0 = time.perf_counter()
        if not mapping:
            nav.update(fr.reading, synth.DT, fr.gauss)
        nav.slam_update(fr.z, fr.u)
        times.append(time.perf_counter() - t0)
    nav.close()
    tsum = float(np.sum(times))
    return {"value": S * frames / tsum, "unit": UNIT, "cores": min(cores, S), "kind": "port",
            "sample": "%d of %d particles x %d frames of workload %s (%d comps x %d meas), %.1f s CPU wall; "
                      "throughput is per-particle work, independent of the particle count"
                      % (S, wl["P"], frames, workload, wl["N"], wl["M"], tsum),
            "seconds": tsum}


def run_reference(args, rank, world):
    if rank != 0:
        return
    wl_name = args.workload
    from monorfs_b200 import synth
    wl = synth.WORKLOADS[wl_name]
    cores = os.cpu_count() or 1
    # one step = one frame over a bounded particle sample sized for ~1-3 s per step
    per_pf = {"c4": 0.6, "c4s": 0.6, "c2": 0.05, "c3": 25.0, "tiny": 0.001}.get(wl_name, 0.1)
    S = int(max(cores, min(wl["P"], round(2.0 * cores / per_pf))))
    S = max(cores, (S // cores) * cores)
    from oracle import orc
    sc = synth.make_workload(wl_name, seed=synth.SEED, P=S)
    sc.params["nthreads"] = cores
    cfg = orc.make_config(sc.params)
    mapping = wl_name in MAPPING_ONLY
    nav = orc.Navigator(cfg, S, sc.poses[0], only_mapping=mapping)
    for i in range(S):
        nav.set_pose(i, sc.poses[i])
        nav.set_map(i, sc.map_w, sc.map_m, sc.map_P)
    times = []
    for f in range(args.warmup + args.steps):
        fr = sc.next_frame()
        t0 = time.perf_counter()
        if not mapping:
            nav.update(fr.reading, synth.DT, fr.gauss)
        nav.slam_update(fr.z, fr.u)
        dt = time.perf_counter() - t0
        if f >= args.warmup:
            times.append(dt)
    tsum = float(np.sum(times))
    value = S * len(times) / tsum
    sample = ("%d of %d particles per step (per-particle work; the C# cannot run here, this is the oracle port), "
              "%d host threads" % (S, wl["P"], cores))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tsum / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl_name, "particles": wl["P"], "components": wl["N"], "measurements": wl["M"],
                   "sampled_particles": S},
        "component_updates_per_sec": value * wl["N"] * wl["M"],
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c4", choices=["c4", "c2", "c3", "tiny", "c4s", "c4m"])
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="frames of the host-buffer pass (default min(steps, 10))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    import torch.distributed as dist
    from monorfs_b200 import capi, sharded, synth

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def run_workload(name, steps, warmup, e2e_steps, do_profile):
        wl = synth.WORKLOADS[name]
        P, N, M = wl["P"], wl["N"], wl["M"]
        mapping = name in MAPPING_ONLY   # config 3: independent mapping-only filters (PHD:297-300, 334)
        lo, hi = sharded.block_range(rank, world, P)
        Pl = hi - lo
        sc = synth.make_workload(name, seed=synth.SEED)
        nframes = warmup + steps
        frames = [sc.next_frame() for _ in range(nframes)]
        h = capi.Handle(sc.params, max_particles=Pl, max_components=2 * N, max_measurements=M, max_pairs=16 * M,
                        device=local_rank, resident_frames=nframes)
        nav = sharded.ShardedNavigator(h, P, rank, world, local_rank)

        def initial_state():
            h.reset(Pl, sc.poses[lo], sc.map_w, sc.map_m, sc.map_P)
            h.set_poses(sc.poses[lo:hi])

        # ---------------- device-resident pass: `value`
        initial_state()
        for f, fr in enumerate(frames):
            h.upload_frame_inputs(fr.gauss[lo:hi], fr.z, slot=f)
        h.synchronize()
        stream = torch.cuda.ExternalStream(h.stream, device=local_rank)
        for f in range(warmup):
            nav.frame(frames[f].reading, synth.DT, M, frames[f].u, slot=f, only_mapping=mapping)
        h.synchronize()
        h.counters(reset=True)
        launches0 = h.kernel_launches
        if do_profile:
            h.profile_enable(steps)
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        for f in range(warmup, nframes):
            nav.frame(frames[f].reading, synth.DT, M, frames[f].u, slot=f, only_mapping=mapping)
        e1.record(stream)
        h.synchronize()
        torch.cuda.synchronize()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop() if rank == 0 else None
        launches = h.kernel_launches - launches0
        ctr = h.counters()
        phases = h.phase_cycles()
        dbgc = h.debug_counters()
        prof = h.profile_read(steps) if do_profile else None
        resampled_frames = None
        res = dict(P=P, N=N, M=M, shape=h.launch_shape(), phases=phases, dbg=dbgc, ms_total=ms, steps=steps, launches=launches, counters=ctr, prof=prof, clocks=clocks,
                   mean_components=ctr["comps_out"] / max(1, ctr["particle_frames"]),
                   pairs_per_particle_frame=ctr["pairs"] / max(1, ctr["particle_frames"]))

        # ---------------- host-buffer pass through the C ABI: `e2e`
        if e2e_steps > 0:
            initial_state()
            ne = min(e2e_steps, steps)
            h2d = d2h = 0
            nres = 0

            def host_frame(fr):
                nonlocal h2d, d2h, nres
                if world == 1:
                    if not mapping:
                        h.update(fr.reading, synth.DT, fr.gauss[lo:hi])
                    best, r = h.slam_update(fr.z, fr.u, only_mapping=mapping)
                    bl = best
                else:
                    h.upload_frame_inputs(fr.gauss[lo:hi], fr.z, slot=0)
                    best, r = nav.frame(fr.reading, synth.DT, M, fr.u, slot=0, only_mapping=mapping)
                    bl = best - lo if lo <= best < hi else -1
                h2d += 8 * (6 * Pl + 3 * M + 6)
                d2h += 64
                if bl >= 0:   # what the C# wrapper needs back each frame: the best particle's map
                    w, _, _ = h.get_map(bl)
                    d2h += 8 * 13 * len(w) + 4
                if r:
                    nres += 1
                    d2h += 4 * P

            for f in range(warmup):
                host_frame(frames[f])
            h.synchronize()
            barrier()
            h2d = d2h = 0
            nres = 0
            t0 = time.perf_counter()
            for f in range(warmup, warmup + ne):
                host_frame(frames[f])
            h.synchronize()
            dt = time.perf_counter() - t0
            barrier()
            dt = max_over_ranks(dt)
            res.update(e2e_seconds=dt, e2e_steps=ne, h2d=sum_over_ranks(h2d) / ne, d2h=sum_over_ranks(d2h) / ne,
                       e2e_resamples=nres)
        del resampled_frames
        h.close()
        return res

    peak, peak_kind = load_peaks()
    e2e_steps = args.e2e_steps or min(args.steps, 10)
    main_res = run_workload(args.workload, args.steps, args.warmup, e2e_steps, True)

    P, N, M = main_res["P"], main_res["N"], main_res["M"]
    sec = main_res["ms_total"] / 1e3
    value = P * args.steps / sec
    ctr = main_res["counters"]
    total_pf = sum_over_ranks(ctr["particle_frames"])
    abytes_local = algorithmic_bytes(ctr)

    roofline = None
    if main_res["prof"] is not None and len(main_res["prof"]):
        prof = main_res["prof"]
        kms = float(np.mean(prof[:, 2]))
        per_launch = abytes_local / len(prof)
        achieved = per_launch / (kms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "k_particle_update", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_kind + " MEASURED_PEAKS.json hbm_gbs",
                    "traffic": None, "algorithmic_bytes_per_launch": per_launch, "kernel_ms": kms,
                    "stage_ms": {k: float(np.mean(prof[:, i])) for i, k in enumerate(capi.Handle.STAGES)}}
        tot = float(sum(main_res["phases"].values())) or 1.0
        roofline["phase_share"] = {k: round(v / tot, 4) for k, v in main_res["phases"].items() if v}
        roofline["events_per_particle"] = {k: round(v / max(1, ctr["particle_frames"]), 1)
                                           for k, v in main_res["dbg"].items() if v}
        roofline["phase_kcycles_per_particle"] = {k: round(v / max(1, ctr["particle_frames"]) / 1e3, 1)
                                                  for k, v in main_res["phases"].items() if v}
        tpath = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.workload)
        if os.path.exists(tpath):
            try:
                with open(tpath) as fh:
                    roofline["traffic"] = json.load(fh).get("dram_bytes_per_launch")
            except Exception:
                pass

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": main_res["ms_total"] / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "particles": P, "components": N, "measurements": M,
                   "particles_per_gpu": P // world, "mean_components_per_particle": main_res["mean_components"],
                   "gated_pairs_per_particle_frame": main_res["pairs_per_particle_frame"],
                   "l2_policy": "inputs larger than L2 (per-GPU map state %.1f GB read+written per frame)"
                                % (abytes_local / max(1, args.steps) / 1e9),
                   "sharding": "particles by rank, weight allgather per frame" if world > 1 else "single GPU",
                   "launch_shape": main_res["shape"]},
        "component_updates_per_sec": value * N * M,
        "frames_per_sec": args.steps / sec,
        "gpu_launches": int(sum_over_ranks(main_res["launches"])),
        "clocks": main_res["clocks"] if rank == 0 else None,
    }
    if roofline:
        line["roofline"] = roofline
    if "e2e_seconds" in main_res:
        line["e2e"] = {"value": P * main_res["e2e_steps"] / main_res["e2e_seconds"], "unit": UNIT,
                       "h2d_bytes_per_step": main_res["h2d"], "d2h_bytes_per_step": main_res["d2h"],
                       "steps": main_res["e2e_steps"], "resampling_frames": main_res["e2e_resamples"]}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ncpu = os.cpu_count() or 1
        sample = {"c4": 24 * ncpu, "c4s": 24 * ncpu, "c2": 2000, "c3": ncpu, "tiny": 64}.get(args.workload, 16)
        nfr = {"c4": 3, "c4s": 3, "c2": 4, "c3": 1, "tiny": 4}.get(args.workload, 2)
        line["cpu_baseline"] = cpu_baseline(args.workload, sample, nfr, synth.SEED)

    if not args.no_secondary and args.workload == "c4" and world == 1:
        sec_res = run_workload("c2", max(10, args.steps), 3, 5, False)
        s2 = sec_res["ms_total"] / 1e3
        line["secondary"] = {"workload": "c2", "particles": sec_res["P"], "components": sec_res["N"],
                             "measurements": sec_res["M"], "value": sec_res["P"] * sec_res["steps"] / s2,
                             "unit": UNIT, "ms_per_step": sec_res["ms_total"] / sec_res["steps"],
                             "e2e_value": sec_res["P"] * sec_res["e2e_steps"] / sec_res["e2e_seconds"],
                             "e2e_resampling_frames": sec_res["e2e_resamples"],
                             "mean_components_per_particle": sec_res["mean_components"]}
    del total_pf
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
