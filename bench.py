#!/usr/bin/env python
"""bench.py -- RB-PHD SLAM per-frame update throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c2|c3|...] [--impl reference]

One "step" = one frame = PHDNavigator.Update + SlamUpdate (PHD:295,323) over all particles of the
workload, including weight normalisation / ESS test / resampling when it triggers.
  value  particle-frames/s, whole job, device-resident loop (inputs already in HBM), CUDA events
  e2e    the same metric over the SAME frames through the reference-facing C ABI with HOST buffers each
         frame (rbphd_update + rbphd_slam_update + best-map read-back), wall clock
Before the W warm-up frames the run is advanced by max(0, 15 - W) untimed "settle" frames so that the timed
frames see saturated maps (MaxQuantity active); both passes start from the same initial state.
N > 1 (torchrun): particles sharded by rank (strong scaling: the workload's particle count is fixed); the
collectives (weight allgather, record migration) are issued by librbphd.so itself (rbphd_comm_init_rank).
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
MAPPING_ONLY = {"c3", "c3l"}
LEAVE_ONE_OUT = {"c3l"}
RECORD_BYTES = 80          # algorithmic FP64 record: weight + mean[3] + symmetric cov[6]
PER_PARTICLE_BYTES = 128   # pose read+written (112) + weight read+written (16)
SETTLE_TO = 15             # untimed frames (settle + warm-up) before the timed region
SAMPLE_PER_CORE = 24       # particles per host thread of the CPU arms (c4-shaped workloads)


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


def cpu_sample_size(workload, cores):
    """Particles of the CPU arms: a fixed multiple of the host threads (per-particle work does not depend on
    the particle count, so the throughput of the sample is the throughput of the workload)."""
    from monorfs_b200 import synth
    wl = synth.WORKLOADS[workload]
    per_core = {"c4": SAMPLE_PER_CORE, "c4s": SAMPLE_PER_CORE, "c4m": SAMPLE_PER_CORE, "c2": 4 * SAMPLE_PER_CORE,
                "c2x": 4 * SAMPLE_PER_CORE, "c3": 1, "c3l": 1, "tiny": 4}.get(workload, SAMPLE_PER_CORE)
    return max(1, min(wl["P"], per_core * cores))


def run_oracle(workload, sample_particles, warmup, steps, seed):
    """The oracle (a port: CPU restatement of the C# reference) on all host cores, bounded particle sample."""
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
    for f in range(warmup + steps):
        fr = sc.next_frame()
        t0 = time.perf_counter()
        if not mapping:
            nav.update(fr.reading, synth.DT, fr.gauss)
        nav.slam_update(fr.z, fr.u)
        dt = time.perf_counter() - t0
        if f >= warmup:
            times.append(dt)
    nav.close()
    tsum = float(np.sum(times))
    return {"value": S * len(times) / tsum, "unit": UNIT, "cores": min(cores, S), "kind": "port",
            "sample": "%d of %d particles (%d per host thread) x %d frames after %d warm-up frames of workload %s "
                      "(%d comps x %d meas), %.1f s CPU wall; per-particle work does not depend on the particle count"
                      % (S, wl["P"], max(1, S // max(1, min(cores, S))), len(times), warmup, workload, wl["N"], wl["M"], tsum),
            "seconds": tsum, "ms_per_step": 1e3 * tsum / len(times), "sampled_particles": S}


def run_reference(args, rank, world):
    if rank != 0:
        return
    from monorfs_b200 import synth
    wl = synth.WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    S = cpu_sample_size(args.workload, cores)
    # bounded run: at most 2 warm-up frames and 4 timed frames of the sample (each c4 frame is ~15 s of CPU)
    warm, steps = min(args.warmup, 2), max(1, min(args.steps, 4))
    res = run_oracle(args.workload, S, warm, steps, synth.SEED)
    value = res["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "particles": wl["P"], "components": wl["N"], "measurements": wl["M"],
                   "sampled_particles": S, "timed_frames": steps, "warmup_frames": warm},
        "component_updates_per_sec": value * wl["N"] * wl["M"],
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": res["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_sim_workload(args, device):
    """BASELINE configs 1 and 5 around the GPU navigator (single GPU, host buffers every frame by construction):
    c1 = `monorfs -i=simulation -f=map.world -c=movements.in -p=20 -x` on a synthesised scene / command file,
    written out as data.zip in the reference's format and read back; c5 = a 1000-frame recorded run (data.zip)
    replayed through a 1000-particle navigator (the recorded-run input path of the Loopy PHD processor; the
    smoother itself stays in the C# host)."""
    import tempfile
    from monorfs_b200 import recordio, simulation, synth
    tmp = tempfile.mkdtemp(prefix="monorfs_")
    if args.workload == "c1":
        P, n_lm, nfr = 20, 200, max(args.steps, 30)
        pose0, measurer, landmarks = simulation.synthetic_scene(n_lm)
        commands = simulation.synthetic_commands(nfr)
        with open(os.path.join(tmp, "map.world"), "w") as fh:
            fh.write(recordio.scene_to_text(pose0, measurer, landmarks, lossless=True))
        with open(os.path.join(tmp, "movements.in"), "w") as fh:
            fh.write(recordio.commands_to_text(commands))
        pose0, measurer, landmarks = recordio.parse_scene(open(os.path.join(tmp, "map.world")).read())
        commands = recordio.parse_commands(open(os.path.join(tmp, "movements.in")).read())
        run = simulation.HeadlessRun(pose0, measurer, landmarks, commands, P, device=device)
        rec = run.run()
        secs, nres = run.gpu_seconds, run.resamples
        run.close()
        recordio.save(rec, os.path.join(tmp, "data.zip"))
        back = recordio.load(os.path.join(tmp, "data.zip"))
        ok = len(back.trajectory) == nfr and len(back.maps) == nfr and len(back.measurements) == nfr
        comps = float(np.mean([len(m[1][0]) for m in rec.maps]))
        meas = float(np.mean([len(z) for _, z in rec.measurements]))
        extra = {"data_zip_round_trip": bool(ok), "data_zip_bytes": os.path.getsize(os.path.join(tmp, "data.zip")),
                 "landmarks": n_lm, "mean_best_map_components": comps, "mean_measurements": meas}
        # the same frames through the oracle (the CPU reference arm of this config)
        cpu = None
        if not args.no_cpu_baseline:
            from oracle import orc
            prm = synth.params(n_lm)
            prm["measurer"] = [float(v) for v in measurer]
            prm["nthreads"] = os.cpu_count() or 1
            nav = orc.Navigator(orc.make_config(prm), P, pose0)
            rng = np.random.default_rng(synth.SEED + 1)
            t0 = time.perf_counter()
            for (t, reading), (_, z) in zip(rec.odometry, rec.measurements):
                nav.update(reading, synth.DT, rng.normal(size=(P, 6)))
                nav.slam_update(z, float(rng.random()))
            cs = time.perf_counter() - t0
            nav.close()
            cpu = {"value": P * nfr / cs, "unit": UNIT, "cores": min(P, os.cpu_count() or 1), "kind": "port",
                   "sample": "all %d particles x %d frames of the recorded run" % (P, nfr)}
    else:
        P, n_lm, nfr = 1000, 200, 1000
        pose0, measurer, landmarks = simulation.synthetic_scene(n_lm)
        gen = simulation.HeadlessRun(pose0, measurer, landmarks, simulation.synthetic_commands(nfr), 8, device=device)
        rec = gen.run()
        gen.close()
        recordio.save(rec, os.path.join(tmp, "data.zip"))
        back = recordio.load(os.path.join(tmp, "data.zip"))
        est, fmap, secs, nres = simulation.replay(back, P, device=device)
        extra = {"recorded_frames": len(back.odometry), "data_zip_bytes": os.path.getsize(os.path.join(tmp, "data.zip")),
                 "final_best_map_components": int(len(fmap[0])), "landmarks": n_lm,
                 "note": "PHD navigator replay of the recorded run; the Loopy smoother's outer loop is not part of this repository"}
        cpu = None
    value = P * nfr / secs
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": nfr, "warmup": 0,
            "ms_per_step": 1e3 * secs / nfr, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": dict({"workload": args.workload, "particles": P, "frames": nfr, "resampling_frames": nres,
                            "timing": "navigator calls through the C ABI with host buffers (rbphd_update + rbphd_slam_update "
                                      "+ pose / best-map read-back), wall clock; the simulated vehicle is host code"}, **extra),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 8 * (6 * P + 6), "d2h_bytes_per_step": 8 * 7 * P,
                    "steps": nfr, "resampling_frames": nres},
            "gpu_launches": None}
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))


def kernel_counters():
    """ncu-derived per-particle-frame counters of k_particle_update (profiles/kernel_counters.json, written by
    tools/ncu_summary.py from a --set full capture); only trusted when they were taken on this tree."""
    from monorfs_b200 import build
    path = os.path.join(ROOT, "profiles", "kernel_counters.json")
    if not os.path.exists(path):
        return None, "no ncu capture recorded"
    try:
        with open(path) as fh:
            d = json.load(fh)
    except Exception:
        return None, "unreadable"
    if d.get("source_hash") != build.source_hash():
        return None, "stale: captured on another tree (%s)" % d.get("source_hash")
    return d, "ncu --set full, %s" % d.get("capture", "?")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c4", choices=["c4", "c2", "c3", "c3l", "tiny", "c4s", "c4m", "c2x", "c1", "c5"])
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="frames of the host-buffer pass (default: steps)")
    ap.add_argument("--settle", type=int, default=-1, help="untimed frames before the warm-up (default 15 - warmup)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    args = ap.parse_args()

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    settle = args.settle if args.settle >= 0 else max(0, SETTLE_TO - args.warmup)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    if args.workload in ("c1", "c5"):
        if rank == 0:
            run_sim_workload(args, local_rank)
        return
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

    def run_workload(name, steps, warmup, pre, e2e_steps, do_profile, keep_particles=0):
        """pre: untimed settle frames before the warm-up.  keep_particles: return the final state of that many
        particles of rank 0 (for the parity check)."""
        wl = synth.WORKLOADS[name]
        P, N, M = wl["P"], wl["N"], wl["M"]
        mapping = name in MAPPING_ONLY   # config 3: independent mapping-only filters (PHD:297-300, 334)
        lo, hi = sharded.block_range(rank, world, P)
        Pl = hi - lo
        sc = synth.make_workload(name, seed=synth.SEED)
        skip = pre + warmup
        nframes = skip + steps
        frames = [sc.next_frame() for _ in range(nframes)]
        h = capi.Handle(sc.params, max_particles=Pl, max_components=2 * N, max_measurements=M, max_pairs=16 * M,
                        device=local_rank, resident_frames=nframes)

        def initial_state():
            h.reset(Pl, sc.poses[lo], sc.map_w, sc.map_m, sc.map_P)
            h.set_poses(sc.poses[lo:hi])

        # ---------------- device-resident pass: `value`
        initial_state()
        nav = sharded.ShardedNavigator(h, P, rank, world, local_rank)
        for f, fr in enumerate(frames):
            h.upload_frame_inputs(fr.gauss[lo:hi], fr.z, slot=f)
        h.synchronize()
        stream = torch.cuda.ExternalStream(h.stream, device=local_rank)
        loo = name in LEAVE_ONE_OUT

        def loo_prepare(f):
            # leave-one-out batch: every filter gets this frame's trajectory pose, filter (f mod P) skips the frame
            h.set_poses(np.tile(frames[f].true_pose.reshape(1, 7), (Pl, 1)))
            g = f % P
            h.set_holdout(g - lo if lo <= g < hi else -1)

        for f in range(skip):
            if loo:
                loo_prepare(f)
            nav.frame(frames[f].reading, synth.DT, M, frames[f].u, slot=f, only_mapping=mapping)
        h.synchronize()
        h.counters(reset=True)
        launches0 = h.kernel_launches
        comm0 = h.comm_stats() if world > 1 else None
        if do_profile:
            h.profile_enable(steps)
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        for f in range(skip, nframes):
            if loo:
                loo_prepare(f)
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
        res = dict(P=P, N=N, M=M, Pl=Pl, shape=h.launch_shape(), phases=phases, dbg=dbgc, ms_total=ms, steps=steps,
                   launches=launches, counters=ctr, prof=prof, clocks=clocks, frames_before_timing=skip,
                   mean_components=ctr["comps_out"] / max(1, ctr["particle_frames"]),
                   pairs_per_particle_frame=ctr["pairs"] / max(1, ctr["particle_frames"]))
        if world > 1:
            c1 = h.comm_stats()
            res["comm"] = {k: c1[k] - comm0[k] for k in c1}
        if keep_particles and rank == 0:
            kp = min(keep_particles, Pl)
            res["final"] = dict(counts=h.get_map_counts()[:kp], alphas=h.get_alphas()[:kp], poses=h.get_poses()[:kp],
                                maps=[h.get_map(i) for i in range(kp)], frames=frames, scene=sc, lo=lo)

        # ---------------- host-buffer pass through the C ABI over the same frames: `e2e`
        if e2e_steps > 0:
            if world > 1:
                h.comm_destroy()
            initial_state()
            nav2 = sharded.ShardedNavigator(h, P, rank, world, local_rank)
            ne = min(e2e_steps, steps)
            h2d = d2h = 0
            nres = 0

            def host_frame(fr, f=0):
                nonlocal h2d, d2h, nres
                if loo:
                    h.set_poses(np.tile(fr.true_pose.reshape(1, 7), (Pl, 1)))
                    g = f % P
                    h.set_holdout(g - lo if lo <= g < hi else -1)
                    h2d += 8 * 7 * Pl
                if world == 1:
                    if not mapping:
                        h.update(fr.reading, synth.DT, fr.gauss[lo:hi])
                    best, r = h.slam_update(fr.z, fr.u, only_mapping=mapping)
                    bl = best
                else:
                    h.upload_frame_inputs(fr.gauss[lo:hi], fr.z, slot=0)
                    out = nav2.frame(fr.reading, synth.DT, M, fr.u, slot=0, only_mapping=mapping)
                    best, r = out if out is not None else (0, False)
                    bl = best - lo if lo <= best < hi else -1
                h2d += 8 * (6 * Pl + 3 * M + 6)
                d2h += 64
                if bl >= 0:   # what the C# wrapper needs back each frame: the best particle's map
                    w, _, _ = h.get_map(bl)
                    d2h += 8 * 13 * len(w) + 4
                if r:
                    nres += 1
                    d2h += 4 * P

            for f in range(skip):
                host_frame(frames[f], f)
            h.synchronize()
            barrier()
            h2d = d2h = 0
            nres = 0
            t0 = time.perf_counter()
            for f in range(skip, skip + ne):
                host_frame(frames[f], f)
            h.synchronize()
            dt = time.perf_counter() - t0
            barrier()
            dt = max_over_ranks(dt)
            res.update(e2e_seconds=dt, e2e_steps=ne, h2d=sum_over_ranks(h2d) / ne, d2h=sum_over_ranks(d2h) / ne,
                       e2e_resamples=nres)
        h.close()
        return res

    def parity_check(final, name, nframes):
        """The timed run's final maps of a few particles against the oracle run over the same frames."""
        from oracle import orc
        sc, frames, K = final["scene"], final["frames"], len(final["counts"])
        prm = dict(sc.params)
        prm["nthreads"] = K
        nav = orc.Navigator(orc.make_config(prm), K, sc.poses[0])
        for i in range(K):
            nav.set_pose(i, sc.poses[i])
            nav.set_map(i, sc.map_w, sc.map_m, sc.map_P)
        t0 = time.perf_counter()
        resampled = 0
        for fr in frames[:nframes]:
            nav.update(fr.reading, synth.DT, fr.gauss[:K])
            _, r, _ = nav.slam_update(fr.z, fr.u)
            resampled += int(r)
        secs = time.perf_counter() - t0
        counts_exact, rel_w, rel_m, rel_P, cov_ok = True, 0.0, 0.0, 0.0, True
        eps = np.finfo(float).eps
        for i in range(K):
            ow, om, oP = nav.get_map(i)
            gw, gm, gP = final["maps"][i]
            if len(ow) != len(gw) or int(final["counts"][i]) != len(ow):
                counts_exact = False
                continue
            if not len(ow):
                continue
            rel_w = max(rel_w, float(np.max(np.abs(gw - ow) / np.maximum(np.abs(ow), 1e-300))))
            rel_m = max(rel_m, float(np.max(np.abs(gm - om) / np.maximum(np.abs(om), 1e-12))))
            dP = np.abs(gP - oP)
            rel_P = max(rel_P, float(np.max(dP / np.maximum(np.abs(oP), 1e-15))))
            # covariances: the reference's raw-moment merge (GAUSS:329-344) cancels |m|^2 / |P| leading digits, so a
            # one-ulp difference in a weight moves a merged covariance by ~eps |m|^2
            mm = np.sum(om ** 2, axis=1)[:, None, None]
            cov_ok = cov_ok and bool(np.all(dP <= 1e-9 * np.abs(oP) + 64 * eps * mm + 1e-15))
        max_rel = max(rel_w, rel_m)
        oa = nav.get_alphas()
        alphas_equal = bool(np.allclose(final["alphas"], oa, rtol=1e-9, atol=0))
        pose_abs = float(np.max(np.abs(final["poses"] - nav.get_poses())))
        nav.close()
        return {"workload": name, "particles": K, "frames": nframes, "counts_exact": bool(counts_exact),
                "max_rel": max_rel, "max_rel_weights": rel_w, "max_rel_means": rel_m, "max_rel_covariances": rel_P,
                "covariances_within_1e-9_plus_64eps_m2": cov_ok, "alphas_match_1e-9": alphas_equal,
                "pose_max_abs_diff": pose_abs,
                "oracle_resampling_frames": resampled, "oracle_seconds": round(secs, 1),
                "note": "final maps of the first particles after all untimed + timed frames vs the oracle on the same "
                        "inputs; max_rel = weights and means; covariances carry the raw-moment conditioning term of "
                        "tests/test_gpu_parity.py; valid because the run never resampled"}

    def sharded_parity():
        """N > 1: a small resampling scene sharded over the ranks must equal the single-GPU run bit for bit."""
        P, N, M, nfr = 96, 150, 48, 5
        sc = synth.make_scene(P, N, M, seed=6, min_effective_particle=0.3)
        fr = [sc.next_frame() for _ in range(nfr)]
        lo, hi = sharded.block_range(rank, world, P)
        h = capi.Handle(sc.params, max_particles=hi - lo, max_components=2 * N, max_measurements=M, max_pairs=16 * M,
                        device=local_rank)
        h.reset(hi - lo, sc.poses[lo], sc.map_w, sc.map_m, sc.map_P)
        h.set_poses(sc.poses[lo:hi])
        nv = sharded.ShardedNavigator(h, P, rank, world, local_rank)
        dec = []
        for f in fr:
            h.upload_frame_inputs(f.gauss[lo:hi], f.z, slot=0)
            dec.append(nv.frame(f.reading, synth.DT, M, f.u, slot=0))
        h.synchronize()
        mine = dict(lo=lo, w=h.get_weights(), poses=h.get_poses(), counts=h.get_map_counts(),
                    maps=[h.get_map(i) for i in range(hi - lo)], dec=dec, comm=h.comm_stats())
        h.close()
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        out = None
        if rank == 0:
            ref = capi.Handle(sc.params, max_particles=P, max_components=2 * N, max_measurements=M, max_pairs=16 * M,
                              device=local_rank)
            ref.reset(P, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)
            ref.set_poses(sc.poses)
            rdec = []
            for f in fr:
                ref.update(f.reading, synth.DT, f.gauss)
                rdec.append(ref.slam_update(f.z, f.u))
            rw, rp, rc = ref.get_weights(), ref.get_poses(), ref.get_map_counts()
            ok = True
            for g in gathered:
                a, n = g["lo"], len(g["w"])
                ok = ok and bool(np.allclose(g["w"], rw[a:a + n], rtol=1e-12, atol=0))
                ok = ok and np.array_equal(g["poses"], rp[a:a + n]) and np.array_equal(g["counts"], rc[a:a + n])
                ok = ok and [tuple(d) for d in g["dec"]] == [tuple(d) for d in rdec]
                for i in range(n):
                    for x, y in zip(g["maps"][i], ref.get_map(a + i)):
                        ok = ok and np.array_equal(x, y)
            ref.close()
            out = {"particles": P, "frames": nfr, "world": world, "identical": bool(ok),
                   "resampling_frames": int(sum(int(r) for _, r in rdec)),
                   "migrated_bytes": int(sum(g["comm"]["sent_bytes"] for g in gathered)),
                   "note": "decisions, poses, counts and maps bit-equal to one GPU holding all particles; weights to 1e-12"}
        return out

    peak, peak_kind = load_peaks()
    e2e_steps = args.e2e_steps or args.steps
    keep = 4 if (world == 1 and not args.no_parity_check and args.workload not in MAPPING_ONLY) else 0
    main_res = run_workload(args.workload, args.steps, args.warmup, settle, e2e_steps, True, keep_particles=keep)

    P, N, M = main_res["P"], main_res["N"], main_res["M"]
    sec = main_res["ms_total"] / 1e3
    value = P * args.steps / sec
    ctr = main_res["counters"]
    abytes_local = algorithmic_bytes(ctr)

    roofline = None
    if main_res["prof"] is not None and len(main_res["prof"]):
        prof = main_res["prof"]
        kms = float(np.mean(prof[:, 2]))
        per_launch = abytes_local / len(prof)
        achieved = per_launch / (kms * 1e-3) / 1e9
        pf_per_launch = ctr["particle_frames"] / len(prof)
        roofline = {"bound": "hbm", "kernel": "k_particle_update", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_kind + " MEASURED_PEAKS.json hbm_gbs",
                    "traffic": None, "algorithmic_bytes_per_launch": per_launch, "kernel_ms": kms,
                    "particle_frames_per_launch": pf_per_launch,
                    "stage_ms": {k: float(np.mean(prof[:, i])) for i, k in enumerate(capi.Handle.STAGES)}}
        tot = float(sum(main_res["phases"].values())) or 1.0
        roofline["phase_share"] = {k: round(v / tot, 4) for k, v in main_res["phases"].items() if v}
        roofline["events_per_particle"] = {k: round(v / max(1, ctr["particle_frames"]), 1)
                                           for k, v in main_res["dbg"].items() if v}
        roofline["phase_kcycles_per_particle"] = {k: round(v / max(1, ctr["particle_frames"]) / 1e3, 1)
                                                  for k, v in main_res["phases"].items() if v}
        kc, kc_src = kernel_counters()
        roofline["traffic_source"] = kc_src
        fp64 = {"peak_source": "rbphd_bench_fp64 (this run, this GPU)"}
        if rank == 0:
            try:
                pk = capi.bench_fp64(local_rank)
                fp64.update(peak_dfma_tflops=pk["dfma_tflops"], peak_unfused_tflops=pk["dmul_dadd_tflops"],
                            peak_unfused_tinst_per_s=pk["fp64_tinst_per_s_unfused"] * 1e12)
            except Exception as exc:   # instrumentation only
                fp64["error"] = str(exc)
        if kc is not None:
            # per-launch figures of THIS rank: the capture's per-particle-frame counts x this launch's particle-frames
            roofline["traffic"] = kc["dram_bytes_per_particle_frame"] * pf_per_launch
            inst = kc.get("fp64_thread_inst_per_particle_frame")
            if inst:
                rate = inst * pf_per_launch / (kms * 1e-3)
                fp64.update(thread_inst_per_particle_frame=inst, achieved_tinst_per_s=rate,
                            pipe_active_pct_ncu=kc.get("fp64_pipe_active_pct"))
                if fp64.get("peak_unfused_tinst_per_s"):
                    fp64["frac"] = rate / fp64["peak_unfused_tinst_per_s"]
        roofline["fp64"] = fp64

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": main_res["ms_total"] / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "particles": P, "components": N, "measurements": M,
                   "particles_per_gpu": P // world, "mean_components_per_particle": main_res["mean_components"],
                   "gated_pairs_per_particle_frame": main_res["pairs_per_particle_frame"],
                   "settle_frames": settle, "frames_before_timing": main_res["frames_before_timing"],
                   "l2_policy": "inputs larger than L2 (per-GPU map state %.1f GB read+written per frame)"
                                % (abytes_local / max(1, args.steps) / 1e9),
                   "sharding": ("particles by rank; weight allgather + record migration issued by librbphd.so (NCCL)"
                                if world > 1 else "single GPU"),
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
                       "steps": main_res["e2e_steps"], "resampling_frames": main_res["e2e_resamples"],
                       "frames": "the same %d frames as `value`, after the same %d untimed frames"
                                 % (main_res["e2e_steps"], main_res["frames_before_timing"])}
    if "comm" in main_res:
        line["comm"] = {k: int(sum_over_ranks(v)) if k != "resampling_frames" else int(v) for k, v in main_res["comm"].items()}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ncpu = os.cpu_count() or 1
        nfr = {"c4": 2, "c4s": 2, "c4m": 2, "c2": 4, "c2x": 4, "c3": 1, "c3l": 1, "tiny": 4}.get(args.workload, 2)
        line["cpu_baseline"] = run_oracle(args.workload, cpu_sample_size(args.workload, ncpu), 1, nfr, synth.SEED)

    if "final" in main_res and rank == 0:
        nfr_total = main_res["frames_before_timing"] + args.steps
        if main_res.get("e2e_resamples", 0) == 0:
            line["parity_check"] = parity_check(main_res["final"], args.workload, nfr_total)
        else:
            line["parity_check"] = {"skipped": "the run resampled: particles are coupled, a per-particle oracle replay "
                                               "is not meaningful (see tests/test_gpu_parity.py for resampling runs)"}

    if not args.no_secondary and args.workload == "c4" and world == 1:
        sec_res = run_workload("c2", max(10, args.steps), 3, 0, 5, False)
        s2 = sec_res["ms_total"] / 1e3
        line["secondary"] = {"workload": "c2", "particles": sec_res["P"], "components": sec_res["N"],
                             "measurements": sec_res["M"], "value": sec_res["P"] * sec_res["steps"] / s2,
                             "unit": UNIT, "ms_per_step": sec_res["ms_total"] / sec_res["steps"],
                             "e2e_value": sec_res["P"] * sec_res["e2e_steps"] / sec_res["e2e_seconds"],
                             "e2e_resampling_frames": sec_res["e2e_resamples"],
                             "mean_components_per_particle": sec_res["mean_components"]}
    if world > 1:
        # c4 never resamples (its WeightAlpha underflows to 0): the migration path is measured
        # on config 2's per-particle shape at config 4's particle count, which resamples every frame
        if not args.no_secondary:
            mig = run_workload("c2x", 10, 3, 0, 0, True)
            s2 = mig["ms_total"] / 1e3
            prof = mig["prof"]
            comm = {k: int(sum_over_ranks(v)) if k != "resampling_frames" else int(v) for k, v in mig["comm"].items()}
            line["migration"] = {"workload": "c2x", "particles": mig["P"], "components": mig["N"],
                                 "measurements": mig["M"], "value": mig["P"] * mig["steps"] / s2, "unit": UNIT,
                                 "ms_per_step": mig["ms_total"] / mig["steps"],
                                 "resampling_frames": comm["resampling_frames"],
                                 "migration_ms": float(np.mean(prof[:, 4])) if prof is not None and len(prof) else None,
                                 "tail_ms": float(np.mean(prof[:, 3])) if prof is not None and len(prof) else None,
                                 "kernel_ms": float(np.mean(prof[:, 2])) if prof is not None and len(prof) else None,
                                 "migration_bytes_per_frame": comm["sent_bytes"] / max(1, mig["steps"]),
                                 "records_per_frame": comm["records"] / max(1, mig["steps"])}
        line["sharded_parity"] = sharded_parity()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
