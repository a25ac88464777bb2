"""Particle sharding across the GPUs of one box (one process per GPU, torch.distributed for plumbing).

Particles are independent inside Update and the Parallel.For body of SlamUpdate (PHD:326-339); the only
coupling is the weight normalisation / ESS test / resampling at PHD:343-358.  Rank g owns the block
[g*P/G, (g+1)*P/G).  Per frame:
  1. every rank runs the fused per-particle kernel on its block (no communication);
  2. ONE allgather of the un-normalised weights (8 B per particle);
  3. every rank runs the identical serial normalise / argmax / ESS / wheel code on the identical global
     vector, so all ranks obtain the same ancestors without a second exchange;
  4. only when resampling fired: ancestors' (pose, map) records move between ranks (send/recv), each
     distinct remote ancestor once per destination rank.

The exchange plan (which records go where) is pure index arithmetic and is unit-tested on CPU with the
gloo backend (tests/test_sharded_gloo.py); on GPUs the same plan drives NCCL send/recv on device buffers.
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
    """What this rank must send / receive / copy locally after a resampling decision.

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


class DevArray:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can alias it."""

    def __init__(self, ptr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


class ShardedNavigator:
    """Update / SlamUpdate over particles sharded across ranks; wraps one capi.Handle per process."""

    def __init__(self, handle, total_particles, rank=0, world=1, device=0):
        self.h = handle
        self.rank, self.world, self.total = rank, world, total_particles
        self.lo, self.hi = block_range(rank, world, total_particles)
        self.device = device
        self._torch = None
        self._stream = None
        self._gw = None
        if world > 1:
            import torch
            self._torch = torch
            self._stream = torch.cuda.ExternalStream(handle.stream, device=device)
            self._gw = torch.empty(total_particles, dtype=torch.float64, device="cuda:%d" % device)
            counts = [block_range(r, world, total_particles) for r in range(world)]
            self._even = all(c[1] - c[0] == counts[0][1] - counts[0][0] for c in counts)
            self._parts = [self._gw[c[0]:c[1]] for c in counts]
        self.resamples = 0

    def frame(self, reading, dt, m, u, slot=0, only_mapping=False):
        """One Update + SlamUpdate; inputs must already be in input slot `slot`."""
        if self.world == 1:
            self.h.frame_async(reading, dt, m, u, only_mapping=only_mapping, slot=slot)
            return None
        torch = self._torch
        import torch.distributed as dist
        # local phase: poses + fused per-particle kernel (weights *= alpha), nothing copied back
        self.h.update_async(reading, dt, slot)
        self.h.slam_update_local(m, only_mapping=only_mapping, slot=slot)
        if only_mapping:
            return None
        ptr, n = self.h.device_weights()
        lw = torch.as_tensor(DevArray(ptr, (n,)), device="cuda:%d" % self.device)
        with torch.cuda.stream(self._stream):
            if self._even:
                dist.all_gather_into_tensor(self._gw, lw)
            else:
                dist.all_gather(self._parts, lw)
        best, res, anc = self.h.resample_global(self._gw.data_ptr(), self.total, self.lo, u)
        if res:
            self.resamples += 1
            self._migrate(anc)
        return best, res

    def _migrate(self, ancestors):
        torch = self._torch
        import torch.distributed as dist
        plan = migration_plan(ancestors, self.rank, self.world)
        rd = self.h.record_doubles()
        ops, recv_bufs = [], {}
        dests = sorted(plan["send"].items())
        send_all = [i for _, idx in dests for i in idx]
        with torch.cuda.stream(self._stream):
            if send_all:
                ptr, nbytes = self.h.pack_particles(send_all)      # one buffer, records in destination order
                sendbuf = torch.as_tensor(DevArray(ptr, (nbytes // 8,)), device="cuda:%d" % self.device)
                off = 0
                for dest, idx in dests:
                    ops.append(dist.P2POp(dist.isend, sendbuf[off:off + rd * len(idx)], dest))
                    off += rd * len(idx)
            for src, items in sorted(plan["recv"].items()):
                buf = torch.empty(len(items) * rd, dtype=torch.float64, device="cuda:%d" % self.device)
                recv_bufs[src] = buf
                ops.append(dist.P2POp(dist.irecv, buf, src))
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
        self._torch.cuda.current_stream(self.device).synchronize()
        self._stream.synchronize()
        for src, items in sorted(plan["recv"].items()):
            buf = recv_bufs[src]
            records = [k for k, (_, slots) in enumerate(items) for _ in slots]
            targets = [slot for _, slots in items for slot in slots]
            self.h.unpack_particles(buf.data_ptr(), records, targets)
        self.h.commit_resample_local(plan["local_sources"])
