"""The N>1 path on CPU: two processes (gloo) shard the particles of the ORACLE navigator exactly as
monorfs_b200.sharded shards the GPU engine -- local map update, weight allgather, identical
normalise / ESS / wheel on every rank, record migration by the same plan -- and must reproduce the
single-process result bit for bit."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P, N, M, FRAMES, SEED = 10, 30, 12, 6, 17


def _single():
    from monorfs_b200 import synth
    from oracle import orc
    sc = synth.make_scene(P, N, M, seed=SEED, min_effective_particle=0.6)
    nav = orc.Navigator(orc.make_config(sc.params), P, sc.poses[0])
    for i in range(P):
        nav.set_pose(i, sc.poses[i])
        nav.set_map(i, sc.map_w, sc.map_m, sc.map_P)
    res = 0
    for _ in range(FRAMES):
        fr = sc.next_frame()
        nav.update(fr.reading, synth.DT, fr.gauss)
        _, r, _ = nav.slam_update(fr.z, fr.u)
        res += int(r)
    return nav.get_weights(), [nav.get_map(i) for i in range(P)], nav.get_poses(), res


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from monorfs_b200 import sharded, synth
    from oracle import orc
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    sc = synth.make_scene(P, N, M, seed=SEED, min_effective_particle=0.6)
    cfg = orc.make_config(sc.params)
    lo, hi = sharded.block_range(rank, world, P)
    Pl = hi - lo
    nav = orc.Navigator(cfg, Pl, sc.poses[0])
    nav.set_weights(np.full(Pl, 1.0 / P))
    for i in range(Pl):
        nav.set_pose(i, sc.poses[lo + i])
        nav.set_map(i, sc.map_w, sc.map_m, sc.map_P)
    best = 0
    for _ in range(FRAMES):
        fr = sc.next_frame()
        nav.update(fr.reading, synth.DT, fr.gauss[lo:hi])
        nav.map_update_range(fr.z, 0, Pl)                       # Parallel.For body, weights *= alpha
        local = torch.from_numpy(nav.get_weights().copy())
        parts = [torch.zeros(sharded.block_range(r, world, P)[1] - sharded.block_range(r, world, P)[0],
                             dtype=torch.float64) for r in range(world)]
        dist.all_gather(parts, local)                           # the ONE collective of a frame
        gw = torch.cat(parts).numpy()
        gw, best, anc, res = orc.normalize_resample(cfg, gw, fr.u, best=best)
        nav.set_weights(gw[lo:hi])
        if res:
            plan = sharded.migration_plan(anc, rank, world)
            maps = [nav.get_map(i) for i in range(Pl)]
            poses = nav.get_poses()
            # this rank's outgoing records, keyed by destination; exchanged with one object allgather
            outgoing = {}
            for dest, idx in sorted(plan["send"].items()):
                outgoing[dest] = []
                for i in idx:
                    w, m, Pm = maps[i]
                    outgoing[dest].append(np.concatenate([[len(w)], poses[i], w, m.ravel(), Pm.ravel()]))
            everyone = [None] * world
            dist.all_gather_object(everyone, outgoing)
            recv_bufs = {}
            for src, items in sorted(plan["recv"].items()):
                sent = everyone[src][rank]
                assert len(sent) == len(items)
                for j in range(len(items)):
                    recv_bufs[(src, j)] = sent[j]
            new_maps, new_poses = [None] * Pl, np.zeros((Pl, 7))
            for s, a in enumerate(plan["local_sources"]):
                if a >= 0:
                    new_maps[s], new_poses[s] = maps[a], poses[a]
            for src, items in plan["recv"].items():
                for j, (a, slots) in enumerate(items):
                    rec = recv_bufs[(src, j)]
                    n = int(rec[0])
                    pose = rec[1:8]
                    w = rec[8:8 + n]
                    m = rec[8 + n:8 + 4 * n].reshape(n, 3)
                    Pm = rec[8 + 4 * n:8 + 13 * n].reshape(n, 3, 3)
                    for s in slots:
                        new_maps[s], new_poses[s] = (w, m, Pm), pose
            for i in range(Pl):
                nav.set_pose(i, new_poses[i])
                nav.set_map(i, *new_maps[i])
    out = dict(rank=rank, lo=lo, w=nav.get_weights(), maps=[nav.get_map(i) for i in range(Pl)],
               poses=nav.get_poses())
    q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_reproduces_single_process():
    ew, emaps, eposes, nres = _single()
    assert nres >= 1, "the run must resample at least once to exercise migration"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for o in outs:
        lo = o["lo"]
        n = len(o["w"])
        assert np.array_equal(o["w"], ew[lo:lo + n])
        assert np.array_equal(o["poses"], eposes[lo:lo + n])
        for i in range(n):
            for a, b in zip(o["maps"][i], emaps[lo + i]):
                assert np.array_equal(a, b)
