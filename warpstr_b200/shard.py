"""Sharding of a read batch over the GPUs of one box and the gather of per-read results.

Reads are independent given the (tiny, replicated) automata, so the path has no data-path
collective: every rank calls its own slice and only the fixed-width per-read records
{read id, len(seq), len(resc_seq), status, cost, resc_cost} are exchanged -- one all_gather
per batch (NCCL over NVLink when the tensors live on GPUs, gloo in the CPU tests).  The
reference's only parallelism is a process pool over reads (caller/wrapper.py:107-109).
"""
from typing import Dict, List, Optional, Sequence

import numpy as np


def partition_reads(costs: Sequence[float], world: int) -> List[np.ndarray]:
    """Longest-processing-time-first split of reads into ``world`` shards of near-equal total
    cost (cost = T * S, the DP cells of the read).  Deterministic; every rank computes the
    same answer.  Returns, per rank, the ascending indices of its reads."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-costs, kind='stable')
    load = np.zeros(world, dtype=np.float64)
    shards: List[List[int]] = [[] for _ in range(world)]
    if len(order) > 64 * world:
        # LPT on the long tail only matters for the first few; after that round-robin over the
        # sorted list (snake order) is within a read of optimal and O(n)
        head = order[:32 * world]
        tail = order[32 * world:]
    else:
        head, tail = order, order[:0]
    for i in head:
        r = int(np.argmin(load))
        shards[r].append(int(i))
        load[r] += costs[i]
    if len(tail):
        ranks = np.argsort(load, kind='stable')
        period = np.concatenate((ranks, ranks[::-1]))
        assign = period[np.arange(len(tail)) % (2 * world)]
        for r in range(world):
            shards[r].extend(tail[assign == r].tolist())
    return [np.sort(np.asarray(s, dtype=np.int64)) for s in shards]


def gather_records(ids: np.ndarray, ints: np.ndarray, floats: np.ndarray, n_total: int,
                   group=None, device=None) -> Optional[Dict[str, np.ndarray]]:
    """all_gather of this rank's records; every rank returns the batch-ordered arrays.

    ids    int64[n_local]      global read index of each local record
    ints   int32[n_local, 3]   len1, len2, status
    floats float64[n_local, 2] cost1, cost2
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = device if device is not None else torch.device('cpu')
    n_local = int(len(ids))
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([n_local], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine, group=group) if dev.type == 'cuda' else \
        dist.all_gather(list(counts.split(1)), mine, group=group)
    cap = int(counts.max().item())
    rec = torch.zeros((cap, 6), dtype=torch.float64, device=dev)
    if n_local:
        local = np.concatenate((ids.reshape(-1, 1).astype(np.float64), ints.astype(np.float64),
                                floats.astype(np.float64)), axis=1)
        rec[:n_local] = torch.from_numpy(local).to(dev)
    out = [torch.zeros_like(rec) for _ in range(world)]
    dist.all_gather(out, rec, group=group)
    len1 = np.full(n_total, -1, dtype=np.int32)
    len2 = np.full(n_total, -1, dtype=np.int32)
    status = np.full(n_total, -1, dtype=np.int32)
    cost1 = np.full(n_total, np.nan)
    cost2 = np.full(n_total, np.nan)
    for r in range(world):
        n = int(counts[r].item())
        if not n:
            continue
        a = out[r][:n].cpu().numpy()
        idx = a[:, 0].astype(np.int64)
        len1[idx] = a[:, 1].astype(np.int32)
        len2[idx] = a[:, 2].astype(np.int32)
        status[idx] = a[:, 3].astype(np.int32)
        cost1[idx] = a[:, 4]
        cost2[idx] = a[:, 5]
    return dict(len1=len1, len2=len2, status=status, cost1=cost1, cost2=cost2)


def gather_device(len1, len2, status, cost1, cost2, group=None):
    """all_gather of equal-sized per-read device results (every rank called the same number of
    reads): two collectives, 12 + 16 bytes per read.  Returns rank-major concatenations."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    ints = torch.stack((len1, len2, status), dim=1).contiguous()
    flts = torch.stack((cost1, cost2), dim=1).contiguous()
    g_i = torch.empty((world * ints.shape[0], 3), dtype=ints.dtype, device=ints.device)
    g_f = torch.empty((world * flts.shape[0], 2), dtype=flts.dtype, device=flts.device)
    dist.all_gather_into_tensor(g_i, ints, group=group)
    dist.all_gather_into_tensor(g_f, flts, group=group)
    return dict(len1=g_i[:, 0], len2=g_i[:, 1], status=g_i[:, 2], cost1=g_f[:, 0], cost2=g_f[:, 1])


_SEL_CACHE: dict = {}


def _all_gather_rows(t, world, group=None):
    """[world, *t.shape] with every rank's ``t`` (same shape everywhere)."""
    import torch
    import torch.distributed as dist
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    if t.device.type == 'cuda':
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    else:
        dist.all_gather(list(out.unbind(0)), t.contiguous(), group=group)
    return out


def gather_by_id(ids, len1, len2, status, cost1, cost2, counts: Sequence[int], n_total: int, group=None):
    """Id-keyed gather of the per-read results of unequal shards, entirely on the tensors' device and
    without a host synchronisation: ``counts`` (every rank's shard size) is known to all ranks from the
    deterministic partition, so nothing has to be asked.  Two collectives (16 + 16 B per read, padded to
    the largest shard); every rank returns batch-ordered tensors (len1, len2, status int32; cost1, cost2
    float64; -1 / NaN where no rank reported a read)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    assert len(counts) == world
    dev = ids.device
    n_local = int(ids.shape[0])
    cap = int(max(counts)) if len(counts) else 0
    ints = torch.full((cap, 4), -1, dtype=torch.int32, device=dev)
    flts = torch.full((cap, 2), float('nan'), dtype=torch.float64, device=dev)
    if n_local:
        ints[:n_local] = torch.stack((ids.to(torch.int32), len1, len2, status), dim=1)
        flts[:n_local] = torch.stack((cost1, cost2), dim=1)
    if world > 1:
        g_i = _all_gather_rows(ints, world, group)
        g_f = _all_gather_rows(flts, world, group)
    else:
        g_i, g_f = ints[None], flts[None]
    # rows that hold a record, as an index list built on the host (a boolean mask would make torch ask
    # the device for the number of set bits, i.e. synchronise)
    key = (tuple(int(c) for c in counts), str(dev))
    sel = _SEL_CACHE.get(key)
    if sel is None:
        sel = torch.as_tensor(np.concatenate([r * cap + np.arange(int(c), dtype=np.int64)
                                              for r, c in enumerate(counts)] or [np.zeros(0, np.int64)]), device=dev)
        _SEL_CACHE.clear()
        _SEL_CACHE[key] = sel
    rows_i = g_i.reshape(-1, 4).index_select(0, sel)
    rows_f = g_f.reshape(-1, 2).index_select(0, sel)
    idx = rows_i[:, 0].long()
    out_i = torch.full((n_total, 3), -1, dtype=torch.int32, device=dev)
    out_f = torch.full((n_total, 2), float('nan'), dtype=torch.float64, device=dev)
    out_i.index_copy_(0, idx, rows_i[:, 1:])
    out_f.index_copy_(0, idx, rows_f)
    return dict(len1=out_i[:, 0], len2=out_i[:, 1], status=out_i[:, 2], cost1=out_f[:, 0], cost2=out_f[:, 1])


def call_sharded_packed(engine, ids, d_sig, off, lengths, aut, rev, counts: Sequence[int], n_total: int,
                        group=None, d_ids=None):
    """This rank's shard, already resident on its GPU (``d_sig`` etc. as for ``call_packed``; ``ids`` = the
    global index of each local read), through the caller and the id-keyed gather.  Asynchronous: nothing
    here waits for the device."""
    import os
    import time
    import torch
    t0 = time.perf_counter()
    o = engine.call_packed(d_sig, off, lengths, aut, rev, want_seq=False)
    t1 = time.perf_counter()
    if d_ids is None:
        d_ids = torch.as_tensor(np.asarray(ids, dtype=np.int64), device=d_sig.device)
    g = gather_by_id(d_ids, o['len1'], o['len2'], o['status'], o['cost1'], o['cost2'], counts, n_total, group)
    if os.environ.get('WSTR_DEBUG_TIMING'):
        import sys
        print(f'[wstr call_sharded_packed] call_packed {1e3 * (t1 - t0):.1f} ms, gather_by_id '
              f'{1e3 * (time.perf_counter() - t1):.1f} ms (host)', file=sys.stderr, flush=True)
    return g


def call_sharded(engine, signals: Sequence[np.ndarray], aut_ids: Sequence[int], reverse: Sequence[bool],
                 n_states: Sequence[int], group=None) -> Dict[str, np.ndarray]:
    """Every rank holds the same read list (or at least its own shard's signals), calls its
    shard on its GPU and receives everybody's per-read lengths and costs."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    costs = [len(s) * n_states[a] for s, a in zip(signals, aut_ids)]
    mine = partition_reads(costs, world)[rank]
    sub = [signals[i] for i in mine]
    packed = engine.upload(sub, [aut_ids[i] for i in mine], [reverse[i] for i in mine])
    o = engine.call_packed(*packed, want_seq=False)
    ints = torch.stack((o['len1'], o['len2'], o['status']), dim=1).cpu().numpy()
    floats = torch.stack((o['cost1'], o['cost2']), dim=1).cpu().numpy()
    return gather_records(mine, ints, floats, len(signals), group=group, device=engine.device)
