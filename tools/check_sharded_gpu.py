#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun on N >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded_gpu.py

Every rank runs the sharded navigator (particles by rank, weight allgather, identical wheel on every
rank, record migration); rank 0 also runs the same frames on one handle holding all particles.  The
sharded state must equal the single-GPU state bit for bit (weights, poses, component counts, maps)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monorfs_b200 import capi, sharded, synth  # noqa: E402


def main():
    import faulthandler
    faulthandler.enable()
    faulthandler.dump_traceback_later(240, exit=True)   # a hang dumps every thread's stack instead of eating the GPU lease
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    # (the third case has an uneven block partition: 97 particles)
    for (P, N, M, frames, seed, meff) in [(64, 60, 24, 8, 5, 0.5), (96, 150, 48, 5, 6, 0.3), (97, 60, 24, 6, 9, 0.5)]:
        sc = synth.make_scene(P, N, M, seed=seed, min_effective_particle=meff)
        fr = [sc.next_frame() for _ in range(frames)]
        lo, hi = sharded.block_range(rank, world, P)
        h = capi.Handle(sc.params, max_particles=hi - lo, max_components=2 * N, max_measurements=M, max_pairs=16 * M,
                        device=local)
        h.reset(hi - lo, sc.poses[lo], sc.map_w, sc.map_m, sc.map_P)
        h.set_poses(sc.poses[lo:hi])
        nav = sharded.ShardedNavigator(h, P, rank, world, local)
        decisions = []
        for f in fr:
            h.upload_frame_inputs(f.gauss[lo:hi], f.z, slot=0)
            decisions.append(nav.frame(f.reading, synth.DT, M, f.u, slot=0))
        h.synchronize()
        mine = dict(lo=lo, w=h.get_weights(), poses=h.get_poses(), counts=h.get_map_counts(),
                    maps=[h.get_map(i) for i in range(hi - lo)], dec=decisions)
        h.close()
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            ref = capi.Handle(sc.params, max_particles=P, max_components=2 * N, max_measurements=M, max_pairs=16 * M,
                              device=local)
            ref.reset(P, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)
            ref.set_poses(sc.poses)
            rdec = []
            for f in fr:
                ref.update(f.reading, synth.DT, f.gauss)
                rdec.append(ref.slam_update(f.z, f.u))
            rw, rp, rc = ref.get_weights(), ref.get_poses(), ref.get_map_counts()
            nres = sum(int(r) for _, r in rdec)
            for g in gathered:
                a, n = g["lo"], len(g["w"])
                checks = {
                    "weights(1e-12)": np.allclose(g["w"], rw[a:a + n], rtol=1e-12, atol=0),
                    "poses": np.array_equal(g["poses"], rp[a:a + n]),
                    "counts": np.array_equal(g["counts"], rc[a:a + n]),
                    "decisions": [tuple(d) for d in g["dec"]] == [tuple(d) for d in rdec],
                }
                same_maps = True
                for i in range(n):
                    for x, y in zip(g["maps"][i], ref.get_map(a + i)):
                        same_maps = same_maps and np.array_equal(x, y)
                checks["maps"] = same_maps
                bad = [k for k, v in checks.items() if not v]
                if bad:
                    print("  rank slice starting at", a, "differs in", bad,
                          "max weight rel diff %.3e" % float(np.max(np.abs(g["w"] - rw[a:a + n]) / np.maximum(rw[a:a + n], 1e-300))))
                ok = ok and not bad
            ref.close()
            print("sharded vs single GPU: P=%d N=%d M=%d frames=%d world=%d resampling frames=%d -> %s"
                  % (P, N, M, frames, world, nres, "IDENTICAL" if ok else "MISMATCH"))
            ok = ok and nres >= 1
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
