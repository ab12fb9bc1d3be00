"""CPU, world_size 2 over gloo: read sharding and the per-read result gather (the only
exchange on the path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from warpstr_b200.shard import gather_records, partition_reads


def test_partition_is_a_balanced_partition():
    rng = np.random.default_rng(0)
    costs = rng.integers(2000, 60000, size=5000) * 240.0
    for world in (1, 2, 4, 8):
        shards = partition_reads(costs, world)
        allidx = np.sort(np.concatenate(shards))
        assert np.array_equal(allidx, np.arange(len(costs)))
        loads = np.array([costs[s].sum() for s in shards])
        assert loads.max() / loads.mean() < 1.01
    # few, very uneven reads
    shards = partition_reads([100.0, 1.0, 1.0, 1.0, 50.0, 49.0], 2)
    assert sorted(np.concatenate(shards).tolist()) == [0, 1, 2, 3, 4, 5]
    assert abs(sum([100, 1, 1, 1, 50, 49][i] for i in shards[0]) - 101) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        costs = (np.arange(n_total) % 7 + 1) * 1000.0
        mine = partition_reads(costs, world)[rank]
        ints = np.stack((mine * 3, mine * 3 + 1, mine % 2), axis=1).astype(np.int32)
        floats = np.stack((mine * 0.5, mine * 0.25), axis=1)
        out = gather_records(mine, ints, floats, n_total)
        ok = (np.array_equal(out['len1'], np.arange(n_total) * 3) and
              np.array_equal(out['len2'], np.arange(n_total) * 3 + 1) and
              np.array_equal(out['status'], np.arange(n_total) % 2) and
              np.array_equal(out['cost1'], np.arange(n_total) * 0.5) and
              np.array_equal(out['cost2'], np.arange(n_total) * 0.25))
        q.put((rank, bool(ok), len(mine)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_records_two_ranks_gloo():
    world, n_total = 2, 37
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert all(ok for _, ok, _ in res)
    assert sum(n for _, _, n in res) == n_total


def test_chunk_schedule_of_the_pipelined_call():
    """CallerEngine.call_arrays: chunk sizes ramp up from a short first chunk; when the previous
    call's copies were slow the chunks stay small and shrink again at the end."""
    from warpstr_b200.caller import CallerEngine

    class Probe:
        def __init__(self, ms):
            self.ms = ms

        def query(self):
            return True

        def elapsed_time(self, other):
            return other.ms

    class Stub:
        _h2d_probe = None

    for n in (0, 1, 10, 3125, 7000, 100000, 123457):
        for probe in (None, (Probe(0), Probe(48.0), 2650306272), (Probe(0), Probe(100.0), 2650306272)):
            e = Stub()
            e._h2d_probe = probe
            b = CallerEngine._chunk_bounds(e, n, 25000)
            assert b[0] == 0 and b[-1] == n
            sizes = np.diff(b)
            assert (sizes > 0).all() and sizes.sum() == n
            if n >= 100000:
                assert sizes[0] == 3125                       # the only copy nothing overlaps is short
                if probe is not None and probe[1].ms > 66:    # < 40 GB/s: small, mirrored chunks
                    assert sizes.max() <= 12500 and sizes[-1] <= 6250
                else:
                    assert sizes.max() == 25000


def _worker_by_id(rank, world, port, n_total, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from warpstr_b200.shard import gather_by_id
        costs = (np.arange(n_total) % 11 + 1) * 1000.0
        shards = partition_reads(costs, world)
        if world == 2:                       # unequal shards: move a few reads over
            shards = [np.concatenate((shards[0], shards[1][:5])), shards[1][5:]]
        mine = shards[rank]
        t = torch.from_numpy(mine)
        out = gather_by_id(t, (t * 3).int(), (t * 3 + 1).int(), (t % 2).int(), t * 0.5, t * 0.25,
                           [len(s) for s in shards], n_total + 2)
        want = np.arange(n_total)
        ok = (np.array_equal(out['len1'][:n_total].numpy(), want * 3) and
              np.array_equal(out['len2'][:n_total].numpy(), want * 3 + 1) and
              np.array_equal(out['status'][:n_total].numpy(), want % 2) and
              np.array_equal(out['cost1'][:n_total].numpy(), want * 0.5) and
              np.array_equal(out['cost2'][:n_total].numpy(), want * 0.25) and
              (out['len2'][n_total:].numpy() == -1).all() and np.isnan(out['cost2'][n_total:].numpy()).all())
        q.put((rank, bool(ok), len(mine)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_by_id_two_ranks_gloo():
    """The panel's gather: unequal shards, ids scattered over the batch, reads nobody owns stay -1/NaN."""
    world, n_total = 2, 101
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_by_id, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert all(ok for _, ok, _ in res)
    assert sum(n for _, _, n in res) == n_total
