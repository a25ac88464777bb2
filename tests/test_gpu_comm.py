"""Multi-GPU exchange plan on the device (k_migration_plan, used by the in-library NCCL path) against the
pure-Python model in monorfs_b200.sharded -- one GPU is enough: the plan is the same arithmetic on every rank."""
import numpy as np
import pytest

from monorfs_b200 import sharded

pytestmark = pytest.mark.gpu


def expected(anc, counts, rank, world):
    total = len(anc)
    lo, hi = sharded.block_range(rank, world, total)
    plan = sharded.migration_plan(anc, rank, world)
    send_idx, send_doubles = [], np.zeros(world, np.int64)
    for dest in sorted(plan["send"]):
        for i in plan["send"][dest]:
            send_idx.append(i)
            send_doubles[dest] += sharded.record_doubles(counts[lo + i])
    send_off = np.concatenate([[0], np.cumsum([sharded.record_doubles(counts[lo + i]) for i in send_idx])])[:-1] \
        if send_idx else np.zeros(0, np.int64)
    rec_off = np.full(hi - lo, -1, np.int64)
    recv_doubles = np.zeros(world, np.int64)
    off = 0
    nrec = 0
    for src in sorted(plan["recv"]):
        for a, slots in plan["recv"][src]:
            for s in slots:
                rec_off[s] = off
            off += sharded.record_doubles(counts[a])
            recv_doubles[src] += sharded.record_doubles(counts[a])
            nrec += 1
    return plan["local_sources"], rec_off, np.array(send_idx, np.int32), send_off, send_doubles, recv_doubles, nrec


@pytest.mark.parametrize("total,world,kind", [
    (64, 2, "wheel"), (97, 2, "wheel"), (100, 3, "wheel"), (2000, 8, "wheel"), (2000, 8, "one"),
    (257, 8, "identity"), (1000, 4, "two"), (20000, 8, "wheel"),
])
def test_device_plan_matches_model(total, world, kind):
    from monorfs_b200 import capi
    rng = np.random.default_rng(total * 31 + world)
    if kind == "wheel":       # what the systematic wheel produces: sorted, a few heavy ancestors
        w = rng.random(total) ** 6
        w /= w.sum()
        anc = np.minimum(np.searchsorted(np.cumsum(w), (np.arange(total) + rng.random()) / total), total - 1)
    elif kind == "one":
        anc = np.full(total, total // 3)
    elif kind == "two":
        anc = np.sort(rng.choice([3, total - 2], size=total))
    else:
        anc = np.arange(total)
    anc = anc.astype(np.int32)
    counts = rng.integers(0, 50, size=total).astype(np.int32)
    for rank in range(world):
        got = capi.debug_migration_plan(anc, counts, world, rank)
        ls, ro, si, so, sd, rd, nrec = expected(anc, counts, rank, world)
        assert got["sorted"]
        assert np.array_equal(got["local_src"], ls)
        assert np.array_equal(got["rec_off"], ro)
        assert np.array_equal(got["send_idx"], si)
        assert np.array_equal(got["send_off"], so)
        assert np.array_equal(got["send_doubles"], sd)
        assert np.array_equal(got["recv_doubles"], rd)
        assert got["n_recv"] == nrec
    # what one rank sends to another is what the other expects from it
    plans = [capi.debug_migration_plan(anc, counts, world, r) for r in range(world)]
    for a in range(world):
        for b in range(world):
            assert plans[a]["send_doubles"][b] == plans[b]["recv_doubles"][a]


def test_unsorted_ancestors_are_flagged():
    from monorfs_b200 import capi
    anc = np.array([0, 3, 2, 5, 5, 7, 7, 7], np.int32)
    got = capi.debug_migration_plan(anc, np.ones(8, np.int32), 2, 0)
    assert not got["sorted"]
