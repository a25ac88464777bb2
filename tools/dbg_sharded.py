import os, sys, faulthandler
faulthandler.enable()
faulthandler.dump_traceback_later(60, exit=True)
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
from monorfs_b200 import capi, sharded, synth
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def log(*a):
    print("[r%d]" % rank, *a, file=sys.stderr, flush=True)
P, N, M, frames, seed, meff = 64, 60, 24, 4, 5, 0.5
sc = synth.make_scene(P, N, M, seed=seed, min_effective_particle=meff)
fr = [sc.next_frame() for _ in range(frames)]
lo, hi = sharded.block_range(rank, world, P)
h = capi.Handle(sc.params, max_particles=hi - lo, max_components=2 * N, max_measurements=M, max_pairs=16 * M, device=local)
log("handle")
h.reset(hi - lo, sc.poses[lo], sc.map_w, sc.map_m, sc.map_P)
h.set_poses(sc.poses[lo:hi])
log("reset")
nav = sharded.ShardedNavigator(h, P, rank, world, local)
log("comm init done")
for i, f in enumerate(fr):
    h.upload_frame_inputs(f.gauss[lo:hi], f.z, slot=0)
    log("uploaded", i)
    d = nav.frame(f.reading, synth.DT, M, f.u, slot=0)
    log("frame", i, d)
h.synchronize()
log("sync")
h.close()
log("closed")
dist.destroy_process_group()
