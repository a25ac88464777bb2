"""Particle sharding across the GPUs of one box (one process per GPU).

Particles are independent inside Update and the Parallel.For body of SlamUpdate (PHD:326-339); the only
coupling is the weight normalisation / ESS test / resampling at PHD:343-358.  Rank g owns the block
[g*P/G, (g+1)*P/G).  The data path is entirely inside librbphd.so (rbphd_comm_init_rank, include/rbphd.h):
  1. every rank runs the fused per-particle kernel on its block (no communication);
  2. ONE ncclAllGather of the un-normalised weights (8 B per particle);
  3. every rank runs the identical serial normalise / argmax / ESS / wheel code on the identical global
     vector, so all ranks obtain the same ancestors without a second exchange;
  4. only when resampling fired: an allgather of the component counts, then grouped ncclSend / ncclRecv of
     the ancestors' (pose, map) records of 8 + 13 n doubles, each distinct remote ancestor once per
     destination rank.
This module only bootstraps the communicator (the NCCL unique id travels over torch.distributed, or any
other side channel a host has) and holds a pure-Python model of the exchange plan that the tests compare
the device plan with (tests/test_sharded_gloo.py on CPU, tests/test_gpu_comm.py on the GPU).
"""
import numpy as np


def block_range(rank, world, total):
    """Block partition: rank g owns [g*P/G, (g+1)*P/G) (P divisible by G is not required)."""
    lo = (total * rank) // world
    hi = (total * (rank + 1)) // world
    return lo, hi


def owner_of(index, world, total):
    """Rank that owns global particle `index` under block_range."""
    # smallest g with (total*(g+1))//world > index
    g = (index * world) // total
    while (total * (g + 1)) // world <= index:
        g += 1
    while (total * g) // world > index:
        g -= 1
    return g


def migration_plan(ancestors, rank, world):
    """What this rank must send / receive / copy locally after a resampling decision (reference model of
    k_migration_plan).

    ancestors: global array, new particle i takes ancestor ancestors[i] (identical on all ranks).
    Returns dict with
      local_sources: for each local new slot, the LOCAL index of its ancestor or -1 if remote
      send: {dest_rank: [local indices to pack, ascending, distinct]}
      recv: {src_rank: [(global ancestor id, [local slots that take it]), ...] in ascending ancestor order}
    The k-th record a rank sends to `dest` is the k-th entry of recv[src] on `dest`.
    """
    ancestors = np.asarray(ancestors, dtype=np.int64)
    total = len(ancestors)
    lo, hi = block_range(rank, world, total)
    owners = np.array([owner_of(int(a), world, total) for a in ancestors], dtype=np.int64)
    local_sources = np.full(hi - lo, -1, dtype=np.int32)
    recv = {}
    for slot, i in enumerate(range(lo, hi)):
        a, src = int(ancestors[i]), int(owners[i])
        if src == rank:
            local_sources[slot] = a - lo
        else:
            recv.setdefault(src, {}).setdefault(a, []).append(slot)
    recv = {src: sorted((a, slots) for a, slots in d.items()) for src, d in recv.items()}
    send = {}
    for dest in range(world):
        if dest == rank:
            continue
        dlo, dhi = block_range(dest, world, total)
        need = sorted({int(a) for a, o in zip(ancestors[dlo:dhi], owners[dlo:dhi]) if o == rank})
        if need:
            send[dest] = [a - lo for a in need]
    return dict(local_sources=local_sources, send=send, recv=recv)


def record_doubles(n):
    """Doubles of one migration record: [n][pose 7][13 fields x n]."""
    return 8 + 13 * int(n)


def bootstrap_unique_id(rank, world):
    """NCCL unique id from rank 0 to everybody over an already initialised torch.distributed group."""
    import torch
    import torch.distributed as dist
    from . import capi
    if dist.get_backend() == "nccl":
        t = torch.zeros(128, dtype=torch.uint8, device="cuda")
    else:
        t = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


class ShardedNavigator:
    """Update / SlamUpdate over particles sharded across ranks; wraps one capi.Handle per process.

    The handle must already hold this rank's block (reset with hi - lo particles).  With world > 1 the
    constructor joins the library's NCCL communicator; frame() then is a single library call."""

    def __init__(self, handle, total_particles, rank=0, world=1, device=0, unique_id=None):
        self.h = handle
        self.rank, self.world, self.total = rank, world, total_particles
        self.lo, self.hi = block_range(rank, world, total_particles)
        self.device = device
        if world > 1:
            if unique_id is None:
                unique_id = bootstrap_unique_id(rank, world)
            handle.comm_init(unique_id, rank, world, total_particles)

    def frame(self, reading, dt, m, u, slot=0, only_mapping=False):
        """One Update + SlamUpdate; inputs must already be in input slot `slot`.  Returns (best, resampled)
        with GLOBAL particle indices when sharded (known to the host as soon as the call returns), None on a
        single GPU (fully asynchronous; ask handle.frame_result() when needed)."""
        self.h.frame_async(reading, dt, m, u, only_mapping=only_mapping, slot=slot)
        if self.world == 1 or only_mapping:
            return None
        return self.h.frame_result()

    @property
    def resamples(self):
        return self.h.comm_stats()["resampling_frames"] if self.world > 1 else 0
